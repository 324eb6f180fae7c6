"""Reference-named harness mirror of the calls that reach the hot path.

Same names, argument meaning and accept/reject behaviour as the reference's call sites, so the parity
tests read like the reference's own; everything computes on the GPU through the C-ABI.  The C++ twin of
this file, used by the Rust-side integration, is `accumulation_b200/host/ark_mirror.hpp`.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np


@dataclass
class CommitterKey:
    """trivial_pc::CommitterKey{generators, hiding_generator} / ipa_pc::CommitterKey{comm_key, h, s}.

    The hiding generator is registered as the base after the last generator so that
    `commit(elems, randomizer)` is a single device MSM."""
    bases: "object"          # accumulation_b200.Bases: generators || hiding_generator
    num_generators: int

    @staticmethod
    def new(ctx, curve: int, generators_xy, hiding_generator_xy=None, precompute: bool = False) -> "CommitterKey":
        """precompute: also build the window table (once per key; trim / index time in the reference)"""
        gens = np.ascontiguousarray(generators_xy, dtype=np.uint64).reshape(-1, 8)
        allb = gens if hiding_generator_xy is None else np.concatenate(
            [gens, np.ascontiguousarray(hiding_generator_xy, dtype=np.uint64).reshape(1, 8)])
        bases = ctx.register_bases(curve, allb)
        if precompute:
            bases.precompute()
        return CommitterKey(bases, gens.shape[0])

    def supported_num_elems(self) -> int:   # src/hp_as/mod.rs:129,681
        return self.num_generators

    @property
    def curve(self) -> int:
        return self.bases.curve


class PedersenCommitment:
    """ark_poly_commit::trivial_pc::PedersenCommitment (call sites: src/hp_as/mod.rs:196,197,214,377,910-918;
    src/r1cs_nark_as/r1cs_nark/mod.rs:216-261,375-407; src/r1cs_nark_as/mod.rs:394-410,1081-1097)."""

    @staticmethod
    def commit(ck: CommitterKey, elems, randomizer=None):
        """-> (xy[8], infinity).  ark-ec truncates to min(len(generators), len(elems))."""
        el = np.ascontiguousarray(elems, dtype=np.uint64).reshape(-1, 4)[: ck.num_generators]
        ctx = ck.bases.ctx
        if randomizer is None:
            return ctx.msm(ck.bases, el, montgomery=True)
        return ctx.commit(ck.bases, el, hiding_index=ck.num_generators, randomizer_mont=randomizer)


class SuccinctCheckPolynomial:
    """ark_poly_commit::ipa_pc::SuccinctCheckPolynomial(pub Vec<F>) (src/ipa_pc_as/mod.rs:245,400,418)."""

    def __init__(self, ctx, field: int, challenges):
        self.ctx, self.field = ctx, field
        self.challenges = np.ascontiguousarray(challenges, dtype=np.uint64).reshape(-1, 4)

    def compute_coeffs(self):
        return self.ctx.compute_coeffs(self.field, self.challenges)


_MODULI = (0x40000000000000000000000000000000224698FC094CF91B992D30ED00000001,   # Fp: Pallas base / Vesta scalar
           0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001)   # Fq: Pallas scalar / Vesta base


def _fe_to_int(field: int, mont) -> int:
    """host-side scalar helper (one element per IPA round, like ark-ff on the Rust host): Montgomery limbs -> int"""
    m = _MODULI[field]
    v = sum(int(x) << (64 * i) for i, x in enumerate(np.asarray(mont, dtype=np.uint64).reshape(4)))
    return v * pow(1 << 256, -1, m) % m


def _int_to_fe(field: int, v: int) -> np.ndarray:
    m = _MODULI[field]
    w = (v % m) * (1 << 256) % m
    return np.array([(w >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


class InnerProductArgPC:
    """The parts of ark_poly_commit::ipa_pc::InnerProductArgPC that run the MSM."""

    @staticmethod
    def open(ck: CommitterKey, combined_coeffs, point, h_prime_xy, round_challenge, log_d: Optional[int] = None, xi0=None):
        """Core of open_individual_opening_challenges (src/ipa_pc_as/mod.rs:454-462) after the
        host has combined the polynomials and derived h' = xi_0 * h (pass either the point h_prime_xy, or xi0 when h is the
        hiding generator of `ck`: faster, no separate scalar multiplication).  `round_challenge(prev_xi, l, r) -> xi` is the
        host sponge (Montgomery limbs in and out; prev_xi is None in the first round).
        Returns (l_vec, r_vec, final_comm_key_xy, c, challenges)."""
        from . import scalar_field
        ctx = ck.bases.ctx
        field = scalar_field(ck.curve)
        cf = np.ascontiguousarray(combined_coeffs, dtype=np.uint64).reshape(-1, 4)
        k = log_d if log_d is not None else max(cf.shape[0] - 1, 0).bit_length()
        if xi0 is not None:      # h' = xi_0 * (hiding generator registered after the generators): no point needed
            sess = ctx.ipa_open_begin(ck.bases, cf, k, point, None)
            ctx.ipa_open_use_hiding_generator(sess, ck.num_generators, xi0)
        else:
            sess = ctx.ipa_open_begin(ck.bases, cf, k, point, h_prime_xy)
        l_vec, r_vec, challenges, xi = [], [], [], None
        lr = ctx.ipa_open_round(sess) if k else None
        while lr is not None:                      # one library call per round: fold (xi^-1 inside) + next round
            l, r = lr
            xi = np.ascontiguousarray(round_challenge(xi, l, r), dtype=np.uint64).reshape(4)
            l_vec.append(l); r_vec.append(r); challenges.append(xi)
            lr = ctx.ipa_open_fold_round(sess, xi)
        final_key, c = ctx.ipa_open_finish(sess)
        return l_vec, r_vec, final_key, c, challenges

    @staticmethod
    def cm_commit(comm_key: CommitterKey, scalars, hiding_generator_index: Optional[int] = None, randomizer=None):
        if randomizer is None:
            return PedersenCommitment.commit(comm_key, scalars)
        el = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)[: comm_key.num_generators]
        idx = comm_key.num_generators if hiding_generator_index is None else hiding_generator_index
        return comm_key.bases.ctx.commit(comm_key.bases, el, hiding_index=idx, randomizer_mont=randomizer)

    @staticmethod
    def succinct_check_equation(ctx, curve: int, comm, point, value, l_vec, r_vec, round_challenges, h_prime_xy, final_comm_key_xy, c) -> bool:
        """Group equation of IpaPC::succinct_check (src/ipa_pc_as/mod.rs:198-205) with the transcript
        values supplied by the host:  C + v h' + sum_i (xi_i^-1 l_i + xi_i r_i) == c final_comm_key + (h(z) c) h'.
        One 2k + 3 term MSM over unregistered points (accmsm_msm_oneshot) whose result must be the identity; the O(k)
        scalar preparation (inverses, h(z)) is host arithmetic like in the reference.  comm = (xy, inf)."""
        from . import scalar_field
        f = scalar_field(curve)
        m = _MODULI[f]
        xis = [_fe_to_int(f, x) for x in np.asarray(round_challenges, dtype=np.uint64).reshape(-1, 4)]
        k = len(xis)
        z, v, cc = _fe_to_int(f, point), _fe_to_int(f, value), _fe_to_int(f, c)
        hz = 1
        for i, xi in enumerate(xis, start=1):
            hz = hz * (1 + xi * pow(z, 1 << (k - i), m)) % m
        bases = [np.asarray(comm[0], dtype=np.uint64)] + [np.asarray(p[0], dtype=np.uint64) for p in l_vec] + \
                [np.asarray(p[0], dtype=np.uint64) for p in r_vec] + [np.asarray(h_prime_xy, dtype=np.uint64), np.asarray(final_comm_key_xy, dtype=np.uint64)]
        inf = [int(comm[1])] + [int(p[1]) for p in l_vec] + [int(p[1]) for p in r_vec] + [0, 0]
        scal = [1] + [pow(xi, -1, m) for xi in xis] + xis + [(v - hz * cc) % m, (-cc) % m]
        sc = np.array([[(s_ >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)] for s_ in scal], dtype=np.uint64)
        _, res_inf = ctx.msm_oneshot(curve, np.array(bases), sc, montgomery=False, infinity=np.array(inf, dtype=np.uint8))
        return bool(res_inf)

    @staticmethod
    def succinct_check_terms(curve: int, comm, point, value, l_vec, r_vec, round_challenges, h_prime_xy, final_comm_key_xy, c):
        """(bases (2k + 3, 8), infinity flags, canonical scalars (2k + 3, 4)) of the equation above"""
        from . import scalar_field
        f = scalar_field(curve)
        m = _MODULI[f]
        xis = [_fe_to_int(f, x) for x in np.asarray(round_challenges, dtype=np.uint64).reshape(-1, 4)]
        k = len(xis)
        z, v, cc = _fe_to_int(f, point), _fe_to_int(f, value), _fe_to_int(f, c)
        hz = 1
        for i, xi in enumerate(xis, start=1):
            hz = hz * (1 + xi * pow(z, 1 << (k - i), m)) % m
        bases = [np.asarray(comm[0], dtype=np.uint64)] + [np.asarray(p[0], dtype=np.uint64) for p in l_vec] + \
                [np.asarray(p[0], dtype=np.uint64) for p in r_vec] + [np.asarray(h_prime_xy, dtype=np.uint64), np.asarray(final_comm_key_xy, dtype=np.uint64)]
        inf = [int(comm[1])] + [int(p[1]) for p in l_vec] + [int(p[1]) for p in r_vec] + [0, 0]
        scal = [1] + [pow(xi, -1, m) for xi in xis] + xis + [(v - hz * cc) % m, (-cc) % m]
        sc = np.array([[(s_ >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)] for s_ in scal], dtype=np.uint64)
        return np.array(bases), np.array(inf, dtype=np.uint8), sc

    @staticmethod
    def succinct_check_equations(ctx, curve: int, instances) -> list:
        """The equations of ALL inputs and accumulators of one prove / verify (the loops at src/ipa_pc_as/mod.rs:262-270 and
        :625-640 call succinct_check once per instance) as ONE batched call (accmsm_msm_oneshot_batch): `instances` is a list
        of argument tuples of succinct_check_equation (without ctx and curve), all of the same degree."""
        terms = [InnerProductArgPC.succinct_check_terms(curve, *inst) for inst in instances]
        if not terms:
            return []
        _, inf = ctx.msm_oneshot_batch(curve, np.stack([t[0] for t in terms]), np.stack([t[2] for t in terms]), montgomery=False,
                                       infinity=np.stack([t[1] for t in terms]))
        return [bool(x) for x in inf]

    @staticmethod
    def check_final_key(vk: CommitterKey, check_poly_challenges, final_comm_key_xy, final_comm_key_inf: int = 0) -> bool:
        """Tail of check_individual_opening_challenges (reached from decide, src/ipa_pc_as/mod.rs:836-845):
        final_key = cm_commit(vk.comm_key, h.compute_coeffs()); accept iff final_key == proof.final_comm_key."""
        ok, _, _ = vk.bases.ctx.ipa_check_final_key(vk.bases, check_poly_challenges, final_comm_key_xy, final_comm_key_inf)
        return ok


def matrix_vec_mul(ctx, field: int, matrices, inp, wit):
    """r1cs_nark::matrix_vec_mul for A, B, C in one launch (src/r1cs_nark_as/r1cs_nark/mod.rs:443-447)."""
    return ctx.csr_matvec(field, matrices, inp, wit)


class ASForHadamardProducts:
    """Vector / commitment steps of src/hp_as/mod.rs that sit on the hot path."""

    @staticmethod
    def compute_hp(ctx, field, a_vec, b_vec):                       # :278-285
        return ctx.hadamard(field, a_vec, b_vec)

    @staticmethod
    def compute_t_vecs(ctx, field, a_vecs, b_vecs, mu_challenges, hp_vec_len, hiding_vecs=None):   # :288-349
        ha, hb = hiding_vecs if hiding_vecs is not None else (None, None)
        return ctx.tvecs(field, a_vecs, b_vecs, mu_challenges, hp_vec_len, ha, hb)

    @staticmethod
    def combine_vectors(ctx, field, vectors, challenges, hiding_vecs=None):   # :492-512
        return ctx.lincomb(field, vectors, challenges, hiding_vecs)

    @staticmethod
    def scale_vector(ctx, field, vector, coeff):                    # :482-489
        return ctx.scale(field, vector, coeff)

    @staticmethod
    def compute_product_poly_comm(ck: CommitterKey, t_vecs) -> tuple:   # :354-388
        """(low, high): commitments to every t-vector except the middle one, without randomiser."""
        n2 = len(t_vecs)
        mid = (n2 - 1) // 2
        low = [PedersenCommitment.commit(ck, t_vecs[i]) for i in range(mid)]
        high = [PedersenCommitment.commit(ck, t_vecs[i]) for i in range(mid + 1, n2)]
        return low, high

    @staticmethod
    def compute_t_vecs_and_product_poly_comm(ck: CommitterKey, a_vecs, b_vecs, mu_challenges, hp_vec_len, hiding_vecs=None):
        """compute_t_vecs (:288-349) feeding compute_product_poly_comm (:354-388) without the t-vectors leaving HBM.
        -> (low, high) lists of (xy, inf)"""
        ha, hb = hiding_vecs if hiding_vecs is not None else (None, None)
        (low, linf), (high, hinf), _ = ck.bases.ctx.hp_product_poly_comm(ck.bases, a_vecs, b_vecs, mu_challenges, hp_vec_len, ha, hb)
        return [(low[i], int(linf[i])) for i in range(low.shape[0])], [(high[i], int(hinf[i])) for i in range(high.shape[0])]

    @staticmethod
    def decide(ck: CommitterKey, instance, witness) -> bool:        # :894-925
        """instance = (comm_1, comm_2, comm_3) as (xy, inf) pairs; witness = (a_vec, b_vec, randomness|None),
        randomness = (rand_1, rand_2, rand_3).  One fused device call: Hadamard product + the three commitments in a
        shared pass of the MSM pipeline + the comparison."""
        a_vec, b_vec, rand = witness
        a = np.ascontiguousarray(a_vec, dtype=np.uint64).reshape(-1, 4)[: ck.num_generators]
        b = np.ascontiguousarray(b_vec, dtype=np.uint64).reshape(-1, 4)[: ck.num_generators]
        exp_xy = np.array([np.asarray(c[0], dtype=np.uint64) for c in instance])
        exp_inf = np.array([int(c[1]) for c in instance], dtype=np.uint8)
        r = None if rand is None else np.array([np.asarray(x, dtype=np.uint64) for x in rand])
        ok, _, _ = ck.bases.ctx.hp_decide(ck.bases, a, b, exp_xy, exp_inf, hiding_index=ck.num_generators, randomness=r)
        return ok


class R1CSNark:
    """The mat-vec + commitment steps of src/r1cs_nark_as/r1cs_nark/mod.rs (prove :183-218, verify :356-389) and of
    ASForR1CSNark::decide (src/r1cs_nark_as/mod.rs:1052-1097) with the index matrices registered once."""

    def __init__(self, ck: CommitterKey, matrices):
        from . import scalar_field
        self.ck = ck
        self.n_mats = len(matrices)
        self.n_rows = int(np.asarray(matrices[0][0]).size - 1)
        self.csr = ck.bases.ctx.register_csr(scalar_field(ck.curve), matrices)

    def matvec_commit(self, inp, wit, blinders=None):
        """-> ([M (input || witness)], [(comm_xy, inf)])"""
        vecs, xy, inf = self.ck.bases.ctx.csr_matvec_commit(self.ck.bases, self.csr, self.n_mats, self.n_rows, inp, wit,
                                                             hiding_index=self.ck.num_generators, blinders=blinders)
        return vecs, [(xy[i], int(inf[i])) for i in range(self.n_mats)]

    def release(self):
        self.ck.bases.ctx.release_csr(self.csr)
