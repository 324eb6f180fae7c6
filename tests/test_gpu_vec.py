"""GPU parity tests of the field-vector kernels (K3 materialised, K4, K5) through the C-ABI: bit-exact
against the golden fixtures and the C oracle; shapes from the reference's hp-as / r1cs-nark fixtures
(src/hp_as/mod.rs:278-349,482-512; src/r1cs_nark_as/r1cs_nark/mod.rs:443-462; src/ipa_pc_as/mod.rs:391-439)."""
import numpy as np
import pytest

import accumulation_b200 as ab
from oracle import cref
from tests.util import fe_mont, ints, load_golden

pytestmark = pytest.mark.gpu


def test_vec_golden_vectors(ctx):
    for case in load_golden("vec"):
        f = case["field"]
        if "matvec" in case:
            mv = case["matvec"]
            row_ptr, cols, coeffs = [0], [], []
            for row in mv["rows"]:
                for c, col in row:
                    coeffs.append(int(c, 16)); cols.append(col)
                row_ptr.append(len(cols))
            mat = (np.array(row_ptr, np.uint32), np.array(cols, np.uint32), fe_mont(f, coeffs))
            outs = ab.matrix_vec_mul(ctx, f, [mat, mat, mat], fe_mont(f, ints(mv["input"])), fe_mont(f, ints(mv["witness"])))
            for o in outs:
                assert (o == fe_mont(f, ints(mv["out"]))).all()
            continue
        n, L = case["n"], case["len"]
        a = [fe_mont(f, ints(v)) for v in case["a"]]
        b = [fe_mont(f, ints(v)) for v in case["b"]]
        mu = fe_mont(f, ints(case["mu"]))
        H = ab.ASForHadamardProducts
        assert (H.compute_hp(ctx, f, a[0], b[0]) == fe_mont(f, ints(case["hp"]))).all()
        t = H.compute_t_vecs(ctx, f, a, b, mu, L)
        for k in range(2 * n - 1):
            assert (t[k] == fe_mont(f, ints(case["tvecs"][k]))).all()
        if "tvecs_zk" in case:
            t = H.compute_t_vecs(ctx, f, a, b, mu, L, (fe_mont(f, ints(case["ha"])), fe_mont(f, ints(case["hb"]))))
            for k in range(2 * n - 1):
                assert (t[k] == fe_mont(f, ints(case["tvecs_zk"][k]))).all()
        ragged = [v[: L - i] for i, v in enumerate(a)]
        assert (H.combine_vectors(ctx, f, ragged, mu[:n]) == fe_mont(f, ints(case["combine"]))).all()
        assert (H.combine_vectors(ctx, f, ragged, mu[:n], fe_mont(f, ints(case["ha"]))[:5]) == fe_mont(f, ints(case["combine_hiding"]))).all()
        assert (H.scale_vector(ctx, f, a[0], mu[1]) == fe_mont(f, ints(case["scale"]))).all()


@pytest.mark.parametrize("field", [0, 1])
@pytest.mark.parametrize("n", [1, 255, 256, 257, 1 << 16])
def test_hadamard_scale_vs_oracle(ctx, field, n):
    a = cref.gen_scalars(field, 1 + n, n, True)
    b = cref.gen_scalars(field, 2 + n, n, True)
    assert (ctx.hadamard(field, a, b) == cref.hadamard(field, a, b)).all()
    assert (ctx.scale(field, a, b[0]) == cref.scale(field, a, b[0])).all()


@pytest.mark.parametrize("field", [0, 1])
def test_lincomb_ragged_vs_oracle(ctx, field):
    lens = [1000, 977, 1024, 3, 0]
    vecs = [cref.gen_scalars(field, 10 + i, n, True) for i, n in enumerate(lens)]
    ch = cref.gen_scalars(field, 20, len(lens), True)
    hid = cref.gen_scalars(field, 21, 1100, True)
    assert (ctx.lincomb(field, vecs, ch) == cref.combine_vectors(field, vecs, ch)).all()
    got = ctx.lincomb(field, vecs, ch, hid)
    assert got.shape[0] == 1100 and (got == cref.combine_vectors(field, vecs, ch, hid)).all()
    assert ctx.lincomb(field, [], np.zeros((0, 4), np.uint64)).shape[0] == 0


@pytest.mark.parametrize("field", [0, 1])
@pytest.mark.parametrize("n_in,zk", [(1, False), (2, False), (2, True), (3, True), (5, False)])
def test_tvecs_vs_oracle(ctx, field, n_in, zk):
    L = 1000
    a = [cref.gen_scalars(field, 30 + i, L - 7 * i, True) for i in range(n_in)]
    b = [cref.gen_scalars(field, 40 + i, L - 3 * i, True) for i in range(n_in)]
    mu = cref.gen_scalars(field, 50, n_in + 1, True)
    ha = cref.gen_scalars(field, 51, L, True) if zk else None
    hb = cref.gen_scalars(field, 52, L - 1, True) if zk else None
    got = ctx.tvecs(field, a, b, mu, L, ha, hb)
    exp = cref.tvecs(field, a, b, mu, L, ha, hb)
    assert (got == np.asarray(exp).reshape(got.shape)).all()


@pytest.mark.parametrize("field", [0, 1])
@pytest.mark.parametrize("k", [0, 1, 5, 16])
def test_compute_coeffs_combine_evaluate(ctx, field, k):
    ch = cref.gen_scalars(field, 60 + k, k, True)
    coeffs = ctx.compute_coeffs(field, ch)
    assert (coeffs == cref.compute_coeffs(field, ch)).all()
    z = cref.gen_scalars(field, 61, 1, True).reshape(4)
    # evaluate(P_dense, z) == verifier's O(log D) shortcut (src/ipa_pc_as/mod.rs:407-421,439)
    ev = ctx.poly_evaluate(field, coeffs, z)
    assert (ev == cref.succinct_evaluate(field, ch, z).reshape(4)).all()
    assert (ev == cref.poly_evaluate(field, coeffs, z).reshape(4)).all()
    m = 3
    chm = cref.gen_scalars(field, 62 + k, m * k, True).reshape(m, k, 4)
    al = cref.gen_scalars(field, 63, m, True)
    rp = cref.gen_scalars(field, 64, min(2, 1 << k), True)
    assert (ctx.combine_check_polys(field, chm, al, rp) == cref.combine_check_polys(field, chm, al, rp)).all()
    assert (ctx.combine_check_polys(field, chm, al) == cref.combine_check_polys(field, chm, al)).all()


@pytest.mark.parametrize("n", [1, 17, 4097])
def test_poly_evaluate_ragged_lengths(ctx, n):
    cf = cref.gen_scalars(1, 70 + n, n, True)
    z = cref.gen_scalars(1, 71, 1, True).reshape(4)
    assert (ctx.poly_evaluate(1, cf, z) == cref.poly_evaluate(1, cf, z).reshape(4)).all()


def _random_csr(field, n_rows, n_cols, max_nnz, seed, ones_every=3):
    rng = np.random.default_rng(seed)
    nnz = rng.integers(0, max_nnz + 1, n_rows)
    nnz[-1] = 0                                     # last row empty like examples/scaling-nark.rs
    row_ptr = np.zeros(n_rows + 1, np.uint32)
    row_ptr[1:] = np.cumsum(nnz)
    tot = int(row_ptr[-1])
    cols = rng.integers(0, n_cols, tot).astype(np.uint32)
    coeffs = cref.gen_scalars(field, seed, tot, True)
    one = cref.to_mont(field, cref.from_int(1).reshape(1, 4))
    coeffs[::ones_every] = one                      # the coeff.is_one() fast path (:459)
    return row_ptr, cols, coeffs


@pytest.mark.parametrize("field", [0, 1])
def test_csr_matvec_vs_oracle(ctx, field):
    n_rows, n_in, n_wit = 5000, 6, 4994
    mats = [_random_csr(field, n_rows, n_in + n_wit, 8, 80 + i) for i in range(3)]
    inp = cref.gen_scalars(field, 90, n_in, True)
    wit = cref.gen_scalars(field, 91, n_wit, True)
    outs = ab.matrix_vec_mul(ctx, field, mats, inp, wit)
    for (rp, cl, cf), o in zip(mats, outs):
        assert (o == cref.csr_matvec(field, rp, cl, cf, inp, wit)).all()


@pytest.mark.parametrize("curve", [0, 1])
def test_hp_as_decide_accept_reject(ctx, curve):
    """hp_as::decide (src/hp_as/mod.rs:894-925): Hadamard + 3 commitments; accept a valid accumulator and
    reject when a witness element, a randomiser or an instance commitment is corrupted."""
    sf = cref.scalar_field(curve)
    L = 1 << 12
    pts = cref.gen_points(curve, 200 + curve, L + 1)
    ck = ab.CommitterKey.new(ctx, curve, pts[:L], pts[L])
    a = cref.gen_scalars(sf, 201, L, True)
    b = cref.gen_scalars(sf, 202, L, True)
    for zk in (False, True):
        r = [cref.gen_scalars(sf, 203 + i, 1, True).reshape(4) for i in range(3)] if zk else [None] * 3
        prod = cref.hadamard(sf, a, b)
        inst = [cref.commit(curve, pts[:L], v, pts[L] if zk else None, rr) for v, rr in zip((a, b, prod), r)]
        wit = (a, b, tuple(r) if zk else None)
        assert ab.ASForHadamardProducts.decide(ck, inst, wit)
        a_bad = a.copy(); a_bad[17, 0] ^= np.uint64(1)
        assert not ab.ASForHadamardProducts.decide(ck, inst, (a_bad, b, wit[2]))
        bad_inst = [inst[0], inst[1], (inst[0][0], inst[0][1])]
        assert not ab.ASForHadamardProducts.decide(ck, bad_inst, wit)
        if zk:
            r_bad = (r[0], r[2], r[1])
            assert not ab.ASForHadamardProducts.decide(ck, inst, (a, b, r_bad))
    ck.bases.release()


@pytest.mark.parametrize("field", [0, 1])
def test_combine_check_polys_edge_shapes(ctx, field):
    """no h-polynomials at all (only the random linear polynomial), no random polynomial, k = 0"""
    k = 6
    rp = cref.gen_scalars(field, 300, 2, True)
    none = np.zeros((0, k, 4), np.uint64)
    got = ctx.combine_check_polys(field, none, np.zeros((0, 4), np.uint64), rp)
    exp = np.zeros((1 << k, 4), np.uint64); exp[:2] = rp
    assert (got == exp).all()
    ch = cref.gen_scalars(field, 301, 3 * k, True).reshape(3, k, 4)
    al = cref.gen_scalars(field, 302, 3, True)
    assert (ctx.combine_check_polys(field, ch, al) == cref.combine_check_polys(field, ch, al)).all()
    ch0 = np.zeros((2, 0, 4), np.uint64)
    al2 = cref.gen_scalars(field, 303, 2, True)
    assert (ctx.combine_check_polys(field, ch0, al2) == cref.combine_check_polys(field, ch0, al2)).all()   # k = 0: alpha_1 + alpha_2
