"""CPU suite: pins the C oracle (oracle/oracle.c) against the curve KATs of SURVEY.md App. B, the committed
golden fixtures (independent Python big-int oracle) and the algebraic identities of SURVEY.md 8c(4)."""
import numpy as np
import pytest

from oracle import cref, pyref
from tests.util import fe_canon, fe_ints, fe_mont, ints, load_golden, point_result, points_mont, same_point

P, Q = pyref.P_PALLAS_BASE, pyref.Q_PALLAS_SCALAR


def test_field_constants_app_b():
    assert cref.to_int(cref.to_mont(cref.FP, cref.from_int(1))) == 0x3fffffffffffffffffffffffffffffff992c350be41914ad34786d38fffffffd
    assert cref.to_int(cref.to_mont(cref.FQ, cref.from_int(1))) == 0x3fffffffffffffffffffffffffffffff992c350be34205675b2b3e9cfffffffd
    assert cref.to_int(cref.to_mont(cref.FP, cref.to_mont(cref.FP, cref.from_int(1)))) == 0x096d41af7b9cb7147797a99bc3c95d18d7d30dbd8b0de0e78c78ecb30000000f
    assert cref.to_int(cref.to_mont(cref.FQ, cref.to_mont(cref.FQ, cref.from_int(1)))) == 0x096d41af7ccfdaa97fae231004ccf59067bb433d891a16e3fc9678ff0000000f


@pytest.mark.parametrize("field,m", [(0, P), (1, Q)])
def test_field_ops_vs_python(field, m):
    rng = pyref.SplitMix64(5 + field)
    xs = [rng.field(m) for _ in range(200)] + [0, 1, m - 1, 2, (1 << 255) % m]
    ys = [rng.field(m) for _ in range(200)] + [m - 1, m - 1, m - 1, 0, 1]
    a, b = fe_mont(field, xs), fe_mont(field, ys)
    assert fe_ints(field, cref.fe_mul(field, a, b)) == [x * y % m for x, y in zip(xs, ys)]
    assert fe_ints(field, cref.fe_add(field, a, b)) == [(x + y) % m for x, y in zip(xs, ys)]
    assert fe_ints(field, cref.fe_sub(field, a, b)) == [(x - y) % m for x, y in zip(xs, ys)]
    assert fe_ints(field, cref.fe_inv(field, a[:20])) == [pow(x, -1, m) for x in xs[:20]]


@pytest.mark.parametrize("curve", [0, 1])
def test_curve_kats_app_b(curve):
    bf = cref.base_field(curve)
    G = pyref.generator(curve)
    Gm = fe_mont(bf, list(G)).reshape(8)
    assert cref.on_curve(curve, Gm)
    two_g, inf = cref.point_mul(curve, Gm, 0, cref.from_int(2))
    got = tuple(fe_ints(bf, two_g.reshape(2, 4)))
    assert inf == 0 and got == pyref.add(G, G, curve)
    if curve == 0:
        assert got == (0x1c0000000000000000000000000000000efee2ee4411acfc1303c567b0000003,
                       0x2b00000000000000000000000000000017076ec9563fb75e8aea5cdf3bfffffc)
    # group order: q * G = O on Pallas, p * G = O on Vesta; (order - 1) * G = -G
    order = pyref.scalar_modulus(curve)
    _, inf = cref.point_mul(curve, Gm, 0, cref.from_int(order))
    assert inf == 1
    mg, inf = cref.point_mul(curve, Gm, 0, cref.from_int(order - 1))
    assert inf == 0 and tuple(fe_ints(bf, mg.reshape(2, 4))) == pyref.neg(G, curve)


def test_msm_golden_vectors():
    for case in load_golden("msm"):
        curve = case["curve"]
        bases = points_mont(curve, case["bases"])
        scal = fe_canon(ints(case["scalars"]))
        got = cref.msm_ark(curve, bases, scal)
        assert same_point(got, point_result(curve, case["result"])), case["name"]
        # the commit() entry takes Montgomery scalars (into_repr() inside)
        got2 = cref.commit(curve, bases, fe_mont(cref.scalar_field(curve), ints(case["scalars"])))
        assert same_point(got2, point_result(curve, case["result"])), case["name"]


def test_ipa_golden_vectors():
    for case in load_golden("ipa"):
        curve, k = case["curve"], case["k"]
        sf = cref.scalar_field(curve)
        ch = fe_mont(sf, ints(case["challenges"]))
        assert fe_ints(sf, cref.compute_coeffs(sf, ch)) == ints(case["coeffs"])
        key = points_mont(curve, case["key"])
        exp = point_result(curve, case["final_key"])
        ok, xy, inf = cref.ipa_check_final_key(curve, key, ch, exp[0], exp[1])
        assert ok and same_point((xy, inf), exp)
        assert same_point(cref.ipa_fold_key(curve, key, ch), exp)            # App. A.2 identity
        z = fe_mont(sf, [int(case["z"], 16)]).reshape(4)
        assert fe_ints(sf, cref.succinct_evaluate(sf, ch, z)) == [int(case["h_of_z"], 16)]
        assert fe_ints(sf, cref.poly_evaluate(sf, cref.compute_coeffs(sf, ch), z)) == [int(case["h_of_z"], 16)]
        bad = exp[0].copy(); bad[0] ^= np.uint64(1)
        ok, _, _ = cref.ipa_check_final_key(curve, key, ch, bad, 0)
        assert not ok


def test_vec_golden_vectors():
    for case in load_golden("vec"):
        f = case["field"]
        if "matvec" in case:
            mv = case["matvec"]
            row_ptr, cols, coeffs = [0], [], []
            for row in mv["rows"]:
                for c, col in row:
                    coeffs.append(int(c, 16)); cols.append(col)
                row_ptr.append(len(cols))
            out = cref.csr_matvec(f, row_ptr, cols, fe_mont(f, coeffs), fe_mont(f, ints(mv["input"])), fe_mont(f, ints(mv["witness"])))
            assert fe_ints(f, out) == ints(mv["out"])
            continue
        n, L = case["n"], case["len"]
        a = [fe_mont(f, ints(v)) for v in case["a"]]
        b = [fe_mont(f, ints(v)) for v in case["b"]]
        mu = fe_mont(f, ints(case["mu"]))
        assert fe_ints(f, cref.hadamard(f, a[0], b[0])) == ints(case["hp"])
        t = cref.tvecs(f, a, b, mu, L)
        assert [fe_ints(f, t[k]) for k in range(2 * n - 1)] == [ints(v) for v in case["tvecs"]]
        if "tvecs_zk" in case:
            t = cref.tvecs(f, a, b, mu, L, fe_mont(f, ints(case["ha"])), fe_mont(f, ints(case["hb"])))
            assert [fe_ints(f, t[k]) for k in range(2 * n - 1)] == [ints(v) for v in case["tvecs_zk"]]
        ragged = [v[: L - i] for i, v in enumerate(a)]
        assert fe_ints(f, cref.combine_vectors(f, ragged, mu[:n])) == ints(case["combine"])
        assert fe_ints(f, cref.combine_vectors(f, ragged, mu[:n], fe_mont(f, ints(case["ha"]))[:5])) == ints(case["combine_hiding"])
        assert fe_ints(f, cref.scale(f, a[0], mu[1])) == ints(case["scale"])


@pytest.mark.parametrize("curve", [0, 1])
def test_msm_algebraic_identities(curve):
    sf = cref.scalar_field(curve)
    n = 300
    pts = cref.gen_points(curve, 17 + curve, n)
    assert all(cref.on_curve(curve, pts[i]) for i in range(0, n, 37))
    a = cref.gen_scalars(sf, 1, n, True)
    b = cref.gen_scalars(sf, 2, n, True)
    ca, cb, cab = cref.commit(curve, pts, a), cref.commit(curve, pts, b), cref.commit(curve, pts, cref.fe_add(sf, a, b))
    assert same_point(cref.point_add(curve, ca[0], ca[1], cb[0], cb[1]), cab)          # linearity
    zero = cref.commit(curve, pts, np.zeros((n, 4), np.uint64))
    assert zero[1] == 1                                                                  # MSM(bases, 0) = identity
    # randomizer * hiding generator
    r = cref.gen_scalars(sf, 3, 1, True).reshape(4)
    with_r = cref.commit(curve, pts[:-1], a[:-1], pts[-1], r)
    manual = cref.commit(curve, pts, np.concatenate([a[:-1], r.reshape(1, 4)]))
    assert same_point(with_r, manual)
    # combine_succinct_check_polynomials == sum alpha_j * coeffs_j (+ random poly)
    k, m = 4, 3
    ch = cref.gen_scalars(sf, 4, m * k, True).reshape(m, k, 4)
    al = cref.gen_scalars(sf, 5, m, True)
    rp = cref.gen_scalars(sf, 6, 2, True)
    comb = cref.combine_check_polys(sf, ch, al, rp)
    q = pyref.scalar_modulus(curve)
    exp = [0] * (1 << k)
    for j in range(m):
        cj = pyref.compute_coeffs(fe_ints(sf, ch[j]), q)
        aj = fe_ints(sf, al[j:j + 1])[0]
        exp = [(e + aj * c) % q for e, c in zip(exp, cj)]
    for i, v in enumerate(fe_ints(sf, rp)):
        exp[i] = (exp[i] + v) % q
    assert fe_ints(sf, comb) == exp


def _ipa_open_with_oracle(curve, key, coeffs, z, h_prime, challenges):
    """drives the per-round C oracle exactly like the GPU session is driven"""
    sf = cref.scalar_field(curve)
    zv = cref.powers(sf, z, coeffs.shape[0])
    l_vec, r_vec = [], []
    for xi in challenges:
        l, r = cref.ipa_open_round_lr(curve, key, coeffs, zv, h_prime)
        l_vec.append(l); r_vec.append(r)
        key, coeffs, zv = cref.ipa_open_fold(curve, key, coeffs, zv, xi, cref.fe_inv(sf, xi.reshape(1, 4)).reshape(4))
    return l_vec, r_vec, key[0], coeffs[0]


def test_ipa_open_golden_vectors():
    for case in load_golden("ipa_open"):
        curve, k = case["curve"], case["k"]
        sf = cref.scalar_field(curve)
        key = points_mont(curve, case["key"])
        hp = points_mont(curve, [case["h_prime"]]).reshape(8)
        coeffs = fe_mont(sf, ints(case["coeffs"]))
        z = fe_mont(sf, [int(case["z"], 16)]).reshape(4)
        ch = fe_mont(sf, ints(case["challenges"]))
        l_vec, r_vec, fk, c = _ipa_open_with_oracle(curve, key, coeffs, z, hp, ch)
        for got, exp in zip(l_vec, case["l_vec"]):
            assert same_point(got, point_result(curve, exp))
        for got, exp in zip(r_vec, case["r_vec"]):
            assert same_point(got, point_result(curve, exp))
        assert same_point((fk, 0), point_result(curve, case["final_key"]))
        assert fe_ints(sf, c.reshape(1, 4)) == [int(case["c"], 16)]
        comm = point_result(curve, case["comm"])
        v = fe_mont(sf, [int(case["v"], 16)]).reshape(4)
        lx = np.array([p[0] for p in l_vec]); rx = np.array([p[0] for p in r_vec])
        assert cref.ipa_succinct_check(curve, comm, z, v, lx, rx, ch, hp, fk, c)
        bad = c.copy(); bad[0] ^= np.uint64(1)
        assert not cref.ipa_succinct_check(curve, comm, z, v, lx, rx, ch, hp, fk, bad)


def test_ipa_open_larger_self_consistency():
    """k = 7: proof from the restated open verifies under succinct_check and final_key == MSM(key, h coeffs)."""
    curve, k = 0, 7
    sf = cref.scalar_field(curve)
    n = 1 << k
    key = cref.gen_points(curve, 31, n + 1)
    hp, key = key[n], key[:n]
    coeffs = cref.gen_scalars(sf, 32, n, True)
    z = cref.gen_scalars(sf, 33, 1, True).reshape(4)
    ch = cref.gen_scalars(sf, 34, k, True)
    l_vec, r_vec, fk, c = _ipa_open_with_oracle(curve, key, coeffs, z, hp, ch)
    comm = cref.commit(curve, key, coeffs)
    v = cref.poly_evaluate(sf, coeffs, z)
    assert cref.ipa_succinct_check(curve, comm, z, v, np.array([p[0] for p in l_vec]), np.array([p[0] for p in r_vec]), ch, hp, fk, c)
    assert same_point((fk, 0), cref.commit(curve, key, cref.compute_coeffs(sf, ch)))
