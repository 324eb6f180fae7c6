"""GPU parity tests of the IPA opening session (row a7 of SURVEY.md 8a: IpaPC::open as called from
AtomicASForInnerProductArgPC::prove, src/ipa_pc_as/mod.rs:454-462): every (l, r) pair, the final commitment
key and the final coefficient are bit-exact with the per-round oracle restatement; the proof satisfies
succinct_check's group equation and the decider's final-key check (src/ipa_pc_as/mod.rs:836-845)."""
import hashlib

import numpy as np
import pytest

import accumulation_b200 as ab
from accumulation_b200.mirror import _int_to_fe
from oracle import cref
from tests.util import fe_ints, fe_mont, ints, load_golden, point_result, points_mont, same_point

pytestmark = pytest.mark.gpu


def sponge_stand_in(field):
    """host transcript stand-in (the reference uses a Poseidon DomainSeparatedSponge on the host):
    xi = Blake2s(prev_xi || l || r) truncated to 128 bits like the reference's challenge size (src/ipa_pc_as/mod.rs:42)"""
    def squeeze(prev, l, r):
        h = hashlib.blake2s()
        if prev is not None:
            h.update(np.asarray(prev, dtype=np.uint64).tobytes())
        for pt, inf in (l, r):
            h.update(np.asarray(pt, dtype=np.uint64).tobytes()); h.update(bytes([inf]))
        return _int_to_fe(field, int.from_bytes(h.digest()[:16], "little") | 1)
    return squeeze


def oracle_open(curve, key, coeffs, z, hp, squeeze):
    sf = cref.scalar_field(curve)
    zv = cref.powers(sf, z, coeffs.shape[0])
    l_vec, r_vec, chs, xi = [], [], [], None
    while coeffs.shape[0] > 1:
        l, r = cref.ipa_open_round_lr(curve, key, coeffs, zv, hp)
        xi = squeeze(xi, l, r)
        key, coeffs, zv = cref.ipa_open_fold(curve, key, coeffs, zv, xi, cref.fe_inv(sf, xi.reshape(1, 4)).reshape(4))
        l_vec.append(l); r_vec.append(r); chs.append(xi)
    return l_vec, r_vec, key[0], coeffs[0], chs


def test_ipa_open_golden_vectors(ctx):
    for case in load_golden("ipa_open"):
        curve, k = case["curve"], case["k"]
        sf = cref.scalar_field(curve)
        key = points_mont(curve, case["key"])
        hp = points_mont(curve, [case["h_prime"]]).reshape(8)
        coeffs = fe_mont(sf, ints(case["coeffs"]))
        z = fe_mont(sf, [int(case["z"], 16)]).reshape(4)
        chs = iter(fe_mont(sf, ints(case["challenges"])))
        ck = ab.CommitterKey.new(ctx, curve, key)
        l_vec, r_vec, fk, c, _ = ab.InnerProductArgPC.open(ck, coeffs, z, hp, lambda prev, l, r: next(chs), log_d=k)
        for got, exp in zip(l_vec, case["l_vec"]):
            assert same_point(got, point_result(curve, exp))
        for got, exp in zip(r_vec, case["r_vec"]):
            assert same_point(got, point_result(curve, exp))
        assert same_point((fk, 0), point_result(curve, case["final_key"]))
        assert fe_ints(sf, c.reshape(1, 4)) == [int(case["c"], 16)]
        ck.bases.release()


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("k,precompute", [(0, False), (1, False), (1, True), (4, True), (5, False), (10, False), (10, True), (13, True)])
def test_ipa_open_vs_oracle_and_verifies(ctx, curve, k, precompute):
    """config 1 of BASELINE.json is k = 10 (degree 2^10 - 1); polynomials shorter than the key are zero-padded."""
    sf = cref.scalar_field(curve)
    n = 1 << k
    pts = cref.gen_points(curve, 40 + k, n + 1)
    key, hp = pts[:n], pts[n]
    n_coeffs = n if k != 5 else n - 3
    coeffs = cref.gen_scalars(sf, 41 + k, n_coeffs, True)
    padded = np.concatenate([coeffs, np.zeros((n - n_coeffs, 4), np.uint64)])
    z = cref.gen_scalars(sf, 42, 1, True).reshape(4)
    squeeze = sponge_stand_in(sf)
    ck = ab.CommitterKey.new(ctx, curve, key)
    if precompute:
        ck.bases.precompute()
    l_vec, r_vec, fk, c, chs = ab.InnerProductArgPC.open(ck, coeffs, z, hp, squeeze, log_d=k)
    el, er, efk, ec, echs = oracle_open(curve, key, padded, z, hp, squeeze)
    assert len(l_vec) == k
    for a, b in zip(l_vec + r_vec, el + er):
        assert same_point(a, b)
    assert all(np.array_equal(a, b) for a, b in zip(chs, echs))
    assert np.array_equal(fk, efk) and np.array_equal(c, ec)
    if k:
        comm = cref.commit(curve, key, padded)
        v = cref.poly_evaluate(sf, padded, z)
        lx, rx, xs = np.array([p[0] for p in l_vec]), np.array([p[0] for p in r_vec]), np.array(chs)
        assert cref.ipa_succinct_check(curve, comm, z, v, lx, rx, xs, hp, fk, c)
        # the same equation on the device (one 2k + 3 term one-shot MSM == identity), accept and reject
        assert ab.InnerProductArgPC.succinct_check_equation(ctx, curve, comm, z, v, l_vec, r_vec, xs, hp, fk, c)
        v_bad = v.copy(); v_bad[0] ^= np.uint64(1)
        assert not ab.InnerProductArgPC.succinct_check_equation(ctx, curve, comm, z, v_bad, l_vec, r_vec, xs, hp, fk, c)
        assert not ab.InnerProductArgPC.succinct_check_equation(ctx, curve, comm, z, v, r_vec, l_vec, xs, hp, fk, c)
        # all instances of a prove / verify in ONE batched call: accept, reject, accept
        good, bad_inst = (comm, z, v, l_vec, r_vec, xs, hp, fk, c), (comm, z, v_bad, l_vec, r_vec, xs, hp, fk, c)
        assert ab.InnerProductArgPC.succinct_check_equations(ctx, curve, [good, bad_inst, good]) == [True, False, True]
        # the decider's half of check(): final_key == cm_commit(key, h.compute_coeffs())  (GPU, fused K3 -> K2)
        assert ab.InnerProductArgPC.check_final_key(ck, xs, fk, 0)
        bad = fk.copy(); bad[3] ^= np.uint64(2)
        assert not ab.InnerProductArgPC.check_final_key(ck, xs, bad, 0)
    ck.bases.release()


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("k", [1, 6, 11])
def test_ipa_open_with_indexed_hiding_generator(ctx, curve, k):
    """h' = xi_0 * h with h the hiding generator of the key: passing (index, xi_0) instead of the point gives the same
    (l, r), final key and c bit for bit (the inner-product term rides in the round MSM as one more pair)."""
    sf = cref.scalar_field(curve)
    n = 1 << k
    pts = cref.gen_points(curve, 60 + k, n + 1)
    key, h = pts[:n], pts[n]
    xi0 = cref.gen_scalars(sf, 61, 1, True).reshape(4)
    hp, hp_inf = cref.point_mul(curve, h, 0, cref.from_mont(sf, xi0.reshape(1, 4)).reshape(4))
    assert hp_inf == 0
    coeffs = cref.gen_scalars(sf, 62 + k, n, True)
    z = cref.gen_scalars(sf, 63, 1, True).reshape(4)
    squeeze = sponge_stand_in(sf)
    ck = ab.CommitterKey.new(ctx, curve, key, h, precompute=(k != 6))
    a = ab.InnerProductArgPC.open(ck, coeffs, z, hp, squeeze, log_d=k)
    b = ab.InnerProductArgPC.open(ck, coeffs, z, None, squeeze, log_d=k, xi0=xi0)
    for x, y in zip(a[0] + a[1], b[0] + b[1]):
        assert same_point(x, y)
    assert np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])
    el, er, efk, ec, _ = oracle_open(curve, key, coeffs, z, hp, squeeze)
    assert all(same_point(x, y) for x, y in zip(b[0] + b[1], el + er)) and np.array_equal(b[2], efk) and np.array_equal(b[3], ec)
    ck.bases.release()


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("k,rounds,min_log,indexed", [(6, 1, 2, False), (8, 2, 3, True), (10, 3, 4, False), (11, 5, 6, True),
                                                      (11, 4, 11, True), (13, 5, 10, False), (12, 10, 12, True)])
def test_ipa_open_with_materialised_folded_key(ctx, curve, k, rounds, min_log, indexed):
    """accmsm_set_ipa_fold: every `rounds` rounds the session switches to the materialised folded key (one shared-scalar
    batched MSM over the window table, k_fold_* in ipa.cuh).  (l, r) of every round, the final key and c stay bit-exact
    with the oracle, which folds the key round by round like the reference (key_l[i] + xi key_r[i])."""
    sf = cref.scalar_field(curve)
    n = 1 << k
    pts = cref.gen_points(curve, 80 + k, n + 1)
    key, h = pts[:n], pts[n]
    xi0 = cref.gen_scalars(sf, 81, 1, True).reshape(4)
    hp, hp_inf = cref.point_mul(curve, h, 0, cref.from_mont(sf, xi0.reshape(1, 4)).reshape(4))
    assert hp_inf == 0
    coeffs = cref.gen_scalars(sf, 82 + k, n, True)
    z = cref.gen_scalars(sf, 83, 1, True).reshape(4)
    squeeze = sponge_stand_in(sf)
    ck = ab.CommitterKey.new(ctx, curve, key, h, precompute=True)
    launches0 = ctx.kernel_launches()
    ctx.set_ipa_fold(0, 0)
    plain = ab.InnerProductArgPC.open(ck, coeffs, z, None if indexed else hp, squeeze, log_d=k, xi0=xi0 if indexed else None)
    launches_plain = ctx.kernel_launches() - launches0
    ctx.set_ipa_fold(rounds, min_log)
    try:
        got = ab.InnerProductArgPC.open(ck, coeffs, z, None if indexed else hp, squeeze, log_d=k, xi0=xi0 if indexed else None)
    finally:
        ctx.set_ipa_fold()
    assert ctx.kernel_launches() - launches0 - launches_plain != launches_plain      # the fold kernels did run
    el, er, efk, ec, _ = oracle_open(curve, key, coeffs, z, hp, squeeze)
    for res in (plain, got):
        assert all(same_point(x, y) for x, y in zip(res[0] + res[1], el + er))
        assert np.array_equal(res[2], efk) and np.array_equal(res[3], ec)
    ck.bases.release()


def test_set_ipa_fold_rejects_bad_arguments(ctx):
    lib, h = ctx._lib, ctx._h
    assert lib.accmsm_set_ipa_fold(h, -1, 17) == -2 and lib.accmsm_set_ipa_fold(h, 11, 17) == -2
    assert lib.accmsm_set_ipa_fold(h, 5, 32) == -2 and lib.accmsm_set_ipa_fold(None, 5, 17) == -2
    assert lib.accmsm_set_ipa_fold(h, 5, 17) == 0
