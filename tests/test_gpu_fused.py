"""GPU parity tests of the fused steps at the shapes of BASELINE configs 2 and 3:
hp-as prove/decide with 2 inputs of length 2^16, r1cs-nark on a 2^16-constraint system (the scaling-nark circuit:
1 non-zero per row, last row empty, 6 instance variables; plus a denser synthetic), and the ipa-pc-as prover's
combined-polynomial opening (config 1, k = 10)."""
import numpy as np
import pytest

import accumulation_b200 as ab
from oracle import cref
from tests.util import same_point
from tests.test_gpu_ipa_open import oracle_open, sponge_stand_in

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("curve,L,zk", [(0, 1 << 16, True), (0, 1 << 16, False), (1, 3000, True), (0, 1, True)])
def test_hp_as_decide_fused(ctx, curve, L, zk):
    """config 2 decide: 1 Hadamard + 3 commitments (src/hp_as/mod.rs:894-925), accept and reject."""
    sf = cref.scalar_field(curve)
    pts = cref.gen_points(curve, 210 + curve, L + 1)
    ck = ab.CommitterKey.new(ctx, curve, pts[:L], pts[L])
    if L >= 1 << 16:
        ck.bases.precompute()
    a, b = cref.gen_scalars(sf, 211, L, True), cref.gen_scalars(sf, 212, L, True)
    r = [cref.gen_scalars(sf, 213 + i, 1, True).reshape(4) for i in range(3)] if zk else [None] * 3
    prod = cref.hadamard(sf, a, b)
    inst = [cref.commit(curve, pts[:L], v, pts[L] if zk else None, rr) for v, rr in zip((a, b, prod), r)]
    wit = (a, b, tuple(r) if zk else None)
    assert ab.ASForHadamardProducts.decide(ck, inst, wit)
    ok, xy, inf = ctx.hp_decide(ck.bases, a, b, np.array([c[0] for c in inst]), [c[1] for c in inst], hiding_index=L,
                                randomness=None if not zk else np.array(r))
    assert ok and all(same_point((xy[j], inf[j]), inst[j]) for j in range(3))
    b_bad = b.copy(); b_bad[L // 2, 2] ^= np.uint64(1)
    assert not ab.ASForHadamardProducts.decide(ck, inst, (a, b_bad, wit[2]))
    assert not ab.ASForHadamardProducts.decide(ck, [inst[1], inst[0], inst[2]], wit)
    if zk:
        assert not ab.ASForHadamardProducts.decide(ck, inst, (a, b, (r[0], r[1], r[0])))
        assert not ab.ASForHadamardProducts.decide(ck, inst, (a, b, None))
    ck.bases.release()


@pytest.mark.parametrize("curve,n_in,L,zk", [(0, 2, 1 << 16, False), (0, 3, 5000, True), (1, 2, 777, True), (0, 1, 100, False), (0, 6, 300, False)])
def test_hp_as_product_poly_comm_fused(ctx, curve, n_in, L, zk):
    """config 2 prove: t-vectors (3 for n = 2) and the 2n - 2 commitments (src/hp_as/mod.rs:288-388)."""
    sf = cref.scalar_field(curve)
    pts = cref.gen_points(curve, 220 + curve, L)
    ck = ab.CommitterKey.new(ctx, curve, pts)
    a = [cref.gen_scalars(sf, 230 + i, L - (5 * i if L > 100 else 0), True) for i in range(n_in)]
    b = [cref.gen_scalars(sf, 240 + i, L, True) for i in range(n_in)]
    mu = cref.gen_scalars(sf, 250, n_in + 1, True)
    hid = (cref.gen_scalars(sf, 251, L, True), cref.gen_scalars(sf, 252, L, True)) if zk else None
    low, high = ab.ASForHadamardProducts.compute_t_vecs_and_product_poly_comm(ck, a, b, mu, L, hid)
    t = cref.tvecs(sf, a, b, mu, L, *(hid if hid else (None, None)))
    assert len(low) == n_in - 1 and len(high) == n_in - 1
    for j in range(n_in - 1):
        assert same_point(low[j], cref.commit(curve, pts, t[j])), ("low", j)
        assert same_point(high[j], cref.commit(curve, pts, t[n_in + j])), ("high", j)
    # the unfused reference-shaped path gives the same commitments
    tv = ab.ASForHadamardProducts.compute_t_vecs(ctx, sf, a, b, mu, L, hid)
    low2, high2 = ab.ASForHadamardProducts.compute_product_poly_comm(ck, tv)
    assert all(same_point(x, y) for x, y in zip(low + high, low2 + high2))
    ck.bases.release()


def scaling_nark_matrices(field, M, dense=False, seed=1):
    """examples/scaling-nark.rs:22-56: M - 1 constraints a * b = c over the same (a, b), one empty last row;
    z = (1, c_pub.., witness..) with 6 instance variables.  dense: 8 random non-zeros per row instead."""
    n_in, n_var = 6, M
    one = cref.to_mont(field, cref.from_int(1).reshape(1, 4))
    rng = np.random.default_rng(seed)
    mats = []
    for mi in range(3):
        if dense:
            nnz = np.full(M, 8); nnz[-1] = 0
            row_ptr = np.zeros(M + 1, np.uint32); row_ptr[1:] = np.cumsum(nnz)
            cols = rng.integers(0, n_var, int(row_ptr[-1])).astype(np.uint32)
            coeffs = cref.gen_scalars(field, seed + mi, int(row_ptr[-1]), True)
            coeffs[::5] = one
        else:
            row_ptr = np.arange(M + 1, dtype=np.uint32); row_ptr[-1] = M - 1
            cols = np.full(M - 1, n_in + mi, dtype=np.uint32) if mi < 2 else (n_in + 2 + np.arange(M - 1) % (n_var - n_in - 2)).astype(np.uint32)
            coeffs = np.repeat(one, M - 1, axis=0)
        mats.append((row_ptr, cols, coeffs))
    return mats, n_in, n_var - n_in


@pytest.mark.parametrize("dense,zk", [(False, False), (False, True), (True, True)])
def test_r1cs_nark_matvec_commit_fused(ctx, dense, zk):
    """config 3: A z, B z, C z and their three commitments at M = 2^16 (prover :183-218 / decider :1052-1097)."""
    curve, M = 0, 1 << 16
    sf = cref.scalar_field(curve)
    mats, n_in, n_wit = scaling_nark_matrices(sf, M, dense)
    pts = cref.gen_points(curve, 260, M + 1)
    ck = ab.CommitterKey.new(ctx, curve, pts[:M], pts[M])
    ck.bases.precompute()
    nark = ab.R1CSNark(ck, mats)
    inp, wit = cref.gen_scalars(sf, 261, n_in, True), cref.gen_scalars(sf, 262, n_wit, True)
    bl = cref.gen_scalars(sf, 263, 3, True) if zk else None
    vecs, comms = nark.matvec_commit(inp, wit, bl)
    for m in range(3):
        exp_v = cref.csr_matvec(sf, *mats[m], inp, wit)
        assert np.array_equal(vecs[m], exp_v)
        assert same_point(comms[m], cref.commit(curve, pts[:M], exp_v, pts[M] if zk else None, None if not zk else bl[m]))
    # same through the unfused reference-shaped calls
    outs = ab.matrix_vec_mul(ctx, sf, mats, inp, wit)
    for m in range(3):
        assert np.array_equal(outs[m], vecs[m])
        assert same_point(ab.PedersenCommitment.commit(ck, outs[m], None if not zk else bl[m]), comms[m])
    nark.release(); ck.bases.release()


@pytest.mark.parametrize("curve,k,m", [(0, 10, 2), (1, 6, 3), (0, 13, 3)])
def test_ipa_pc_as_prove_open_combined(ctx, curve, k, m):
    """config 1 (k = 10, 1 input + 1 accumulator -> m = 2): the prover's combined check polynomial is built,
    evaluated and opened on the device; equals the host-side combine + evaluate + open."""
    sf = cref.scalar_field(curve)
    n = 1 << k
    pts = cref.gen_points(curve, 270 + k, n + 1)
    key, hp = pts[:n], pts[n]
    ck = ab.CommitterKey.new(ctx, curve, key)
    chm = cref.gen_scalars(sf, 271, m * k, True).reshape(m, k, 4)
    al = cref.gen_scalars(sf, 272, m, True)
    rp = cref.gen_scalars(sf, 273, 2, True)          # the random linear polynomial of the zk variant
    z = cref.gen_scalars(sf, 274, 1, True).reshape(4)
    squeeze = sponge_stand_in(sf)
    combined = cref.combine_check_polys(sf, chm, al, rp)
    sess, ev = ctx.ipa_open_begin_combined(ck.bases, chm, al, z, hp, rp)
    assert np.array_equal(ev, cref.poly_evaluate(sf, combined, z).reshape(4))
    l_vec, r_vec, chs, xi = [], [], [], None
    from accumulation_b200.mirror import _fe_to_int, _int_to_fe, _MODULI
    for _ in range(k):
        l, r = ctx.ipa_open_round(sess)
        xi = squeeze(xi, l, r)
        ctx.ipa_open_fold(sess, xi, _int_to_fe(sf, pow(_fe_to_int(sf, xi), -1, _MODULI[sf])))
        l_vec.append(l); r_vec.append(r); chs.append(xi)
    fk, c = ctx.ipa_open_finish(sess)
    el, er, efk, ec, echs = oracle_open(curve, key, combined, z, hp, squeeze)
    assert all(same_point(x, y) for x, y in zip(l_vec + r_vec, el + er))
    assert np.array_equal(fk, efk) and np.array_equal(c, ec)
    assert ab.InnerProductArgPC.check_final_key(ck, np.array(chs), fk, 0)     # ... and the decider accepts it
    ck.bases.release()
