"""Generates tests/golden/*.json from the independent Python big-int oracle (oracle/pyref.py).

The reference repository holds no golden vectors for this path and cannot be run here
(Rust, no cargo), so these fixtures pin the *mathematical* outputs: every value is computed with plain
Python integers and affine formulas, independently of both the C oracle and the CUDA path.
Run:  python tests/golden/make_golden.py     (deterministic; commit the JSON it writes)
"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import pyref as R  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def hx(v):
    return hex(v)


def pt(p):
    return None if p is None else [hx(p[0]), hx(p[1])]


def msm_cases():
    out = []
    for curve in (R.PALLAS, R.VESTA):
        q = R.scalar_modulus(curve)
        rng = R.SplitMix64(0xACC0 + curve)
        pts = [R.random_point(rng, curve) for _ in range(96)]
        G = R.generator(curve)
        cases = [
            ("generator_times_1", [G], [1]), ("generator_times_2", [G], [2]), ("generator_times_q_minus_1", [G], [q - 1]),
            ("generator_times_q", [G], [q % (1 << 255)]) if False else ("two_generators", [G, G], [5, 7]),
            ("empty", [], []), ("all_zero", pts[:8], [0] * 8), ("all_one", pts[:8], [1] * 8),
            ("all_q_minus_1", pts[:8], [q - 1] * 8), ("one_hot", pts[:8], [0] * 7 + [rng.field(q)]),
            ("constant_scalar", pts[:33], [rng.field(q)] * 33),
            ("p_and_minus_p", [pts[0], R.neg(pts[0], curve), pts[1]], [9, 9, 3]),
            ("duplicates", [pts[2]] * 5, [rng.field(q) for _ in range(5)]),
        ]
        for n in (1, 2, 3, 31, 32, 33, 64):
            cases.append((f"random_n{n}", pts[:n], [rng.field(q) for _ in range(n)]))
        cases.append(("truncated_128bit", pts[:40], [rng.field(q) & ((1 << 128) - 1) for _ in range(40)]))
        for name, bases, scalars in cases:
            res = R.msm_naive(bases, scalars, curve)
            out.append({"curve": curve, "name": name, "bases": [pt(b) for b in bases], "scalars": [hx(s) for s in scalars],
                        "result": pt(res)})
        # a mid-size case through the independent bucket method
        n = 96
        sc = [rng.field(q) for _ in range(n)]
        res = R.msm_bucket(pts, sc, curve, c=6)
        assert res == R.msm_naive(pts, sc, curve)
        out.append({"curve": curve, "name": "random_n96", "bases": [pt(b) for b in pts], "scalars": [hx(s) for s in sc],
                    "result": pt(res)})
    return out


def ipa_cases():
    out = []
    for curve in (R.PALLAS, R.VESTA):
        q = R.scalar_modulus(curve)
        rng = R.SplitMix64(0xACC4 + curve)
        for k in (1, 2, 4, 5):
            key = [R.random_point(rng, curve) for _ in range(1 << k)]
            ch = [rng.field(q) for _ in range(k)]
            coeffs = R.compute_coeffs(ch, q)
            final = R.msm_naive(key, coeffs, curve)
            assert final == R.fold_key(key, ch, curve)   # App. A.2 identity: two independent computations
            z = rng.field(q)
            assert R.horner(coeffs, z, q) == R.succinct_evaluate(ch, z, q)
            out.append({"curve": curve, "k": k, "key": [pt(p) for p in key], "challenges": [hx(c) for c in ch],
                        "coeffs": [hx(c) for c in coeffs], "final_key": pt(final), "z": hx(z),
                        "h_of_z": hx(R.horner(coeffs, z, q))})
    return out


def ipa_open_cases():
    """IpaPC::open with fixed round challenges (k = 1, 3, 4): l_vec, r_vec, final_comm_key, c; the proof must
    satisfy succinct_check's group equation and final_comm_key must equal MSM(key, compute_coeffs(xi))."""
    out = []
    for curve in (R.PALLAS, R.VESTA):
        q = R.scalar_modulus(curve)
        rng = R.SplitMix64(0xACC7 + curve)
        for k in (1, 3, 4):
            n = 1 << k
            key = [R.random_point(rng, curve) for _ in range(n)]
            h_prime = R.random_point(rng, curve)
            coeffs = [rng.field(q) for _ in range(n)]
            z = rng.field(q)
            ch = [rng.field(q) for _ in range(k)]
            l_vec, r_vec, final_key, c = R.ipa_open(key, coeffs, z, h_prime, ch, curve)
            comm = R.msm_naive(key, coeffs, curve)
            v = R.horner(coeffs, z, q)
            assert R.ipa_succinct_check(comm, z, v, l_vec, r_vec, ch, h_prime, final_key, c, curve)
            assert final_key == R.msm_naive(key, R.compute_coeffs(ch, q), curve)
            out.append({"curve": curve, "k": k, "key": [pt(p) for p in key], "h_prime": pt(h_prime), "coeffs": [hx(x) for x in coeffs],
                        "z": hx(z), "challenges": [hx(x) for x in ch], "l_vec": [pt(p) for p in l_vec], "r_vec": [pt(p) for p in r_vec],
                        "final_key": pt(final_key), "c": hx(c), "comm": pt(comm), "v": hx(v)})
    return out


def vec_cases():
    out = []
    for field, m in ((0, R.P_PALLAS_BASE), (1, R.Q_PALLAS_SCALAR)):
        rng = R.SplitMix64(0xACC2 + field)
        L = 11   # the reference fixture length (src/hp_as/mod.rs:959-1045)
        for n in (1, 2, 3):
            a = [[rng.field(m) for _ in range(L)] for _ in range(n)]
            b = [[rng.field(m) for _ in range(L)] for _ in range(n)]
            mu = [1] + [rng.field(m) & ((1 << 128) - 1) for _ in range(n)]
            ha = [rng.field(m) for _ in range(L)]
            hb = [rng.field(m) for _ in range(L)]
            case = {"field": field, "n": n, "len": L, "a": [[hx(x) for x in v] for v in a], "b": [[hx(x) for x in v] for v in b],
                    "mu": [hx(x) for x in mu], "ha": [hx(x) for x in ha], "hb": [hx(x) for x in hb]}
            case["hp"] = [hx(x) for x in R.compute_hp(a[0], b[0], m)]
            case["tvecs"] = [[hx(x) for x in v] for v in R.compute_t_vecs(a, b, mu, L, m)]
            if n >= 2:
                case["tvecs_zk"] = [[hx(x) for x in v] for v in R.compute_t_vecs(a, b, mu, L, m, hiding=(ha, hb))]
            ragged = [v[: L - i] for i, v in enumerate(a)]
            case["combine"] = [hx(x) for x in R.combine_vectors(ragged, mu[:n], m)]
            case["combine_hiding"] = [hx(x) for x in R.combine_vectors(ragged, mu[:n], m, hiding=ha[:5])]
            case["scale"] = [hx(x) for x in R.scale_vector(a[0], mu[1], m)]
            out.append(case)
        # sparse mat-vec: ragged rows, an empty row, unit coefficients
        n_in, n_w = 3, 7
        inp = [rng.field(m) for _ in range(n_in)]
        wit = [rng.field(m) for _ in range(n_w)]
        rows = [[(1, 0), (rng.field(m), 4)], [], [(rng.field(m), i) for i in range(10)], [(1, 9)], [(rng.field(m), 2), (1, 2)]]
        out.append({"field": field, "matvec": {"rows": [[[hx(c), col] for c, col in r] for r in rows], "input": [hx(x) for x in inp],
                                                  "witness": [hx(x) for x in wit],
                                                  "out": [hx(x) for x in R.matrix_vec_mul(rows, inp, wit, m)]}})
    return out


if __name__ == "__main__":
    only = sys.argv[1:]
    for name, fn in (("msm", msm_cases), ("ipa", ipa_cases), ("vec", vec_cases), ("ipa_open", ipa_open_cases)):
        if only and name not in only:
            continue
        path = os.path.join(HERE, f"{name}.json")
        with open(path, "w") as f:
            json.dump(fn(), f, indent=0, separators=(",", ":"))
        print(path, os.path.getsize(path), "bytes")
