"""Quick GPU parity run used during development: MSM sizes / distributions vs the C oracle."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import accumulation_b200 as ab
from oracle import cref

ctx = ab.Context(0)
fails = 0
def check(name, got, exp):
    global fails
    ok = got[1] == exp[1] and np.array_equal(got[0], exp[0])
    print(("PASS " if ok else "FAIL ") + name, flush=True)
    if not ok:
        fails += 1
        print("   got", got, "\n   exp", exp)

for curve in (0, 1):
    sf = cref.scalar_field(curve)
    nmax = 1 << 14
    pts = cref.gen_points(curve, 100 + curve, nmax)
    B = ctx.register_bases(curve, pts)
    for n in (1, 2, 3, 31, 32, 33, 100, 1 << 10, 5000, 1 << 14):
        sc = cref.gen_scalars(sf, 7 * n + curve, n, montgomery=True)
        exp = cref.commit(curve, pts[:n], sc)
        check(f"curve{curve} msm n={n} mont", ctx.msm(B, sc), exp)
    n = 1 << 12
    sc = cref.gen_scalars(sf, 5, n, montgomery=False)
    check(f"curve{curve} canonical scalars", ctx.msm(B, sc, montgomery=False), cref.msm_ark(curve, pts[:n], sc))
    # degenerate distributions: constant, zero, one, q-1, one-hot
    from oracle import pyref
    q = pyref.scalar_modulus(curve)
    one = cref.gen_scalars(sf, 9, 1, True)
    for nm, vec in [("const", np.repeat(one, n, axis=0)), ("zero", np.zeros((n, 4), np.uint64)),
                    ("ones", cref.to_mont(sf, np.tile(cref.from_int(1), (n, 1)))),
                    ("q-1", cref.to_mont(sf, np.tile(cref.from_int(q - 1), (n, 1)))),
                    ("onehot", np.concatenate([np.zeros((n - 1, 4), np.uint64), one]))]:
        check(f"curve{curve} {nm}", ctx.msm(B, vec), cref.commit(curve, pts[:n], vec))
    # offset sub-range and duplicated / opposite points
    sc = cref.gen_scalars(sf, 77, 1000, True)
    check(f"curve{curve} offset", ctx.msm(B, sc, offset=123), cref.commit(curve, pts[123:1123], sc))
    # IPA final key
    for k in (1, 4, 10):
        ch = cref.gen_scalars(sf, 1000 + k, k, True)
        got = ctx.ipa_final_key(B, ch)
        ok, exp_xy, exp_inf = cref.ipa_check_final_key(curve, pts[: 1 << k], ch, got[0], got[1])
        check(f"curve{curve} ipa_final_key k={k}", got, (exp_xy, exp_inf))
    B.release()
print("timings", ctx.last_timings())
if len(sys.argv) > 1:
    n = 1 << int(sys.argv[1])
    pts = cref.gen_points(0, 1, n); B = ctx.register_bases(0, pts)
    sc = cref.gen_scalars(cref.FQ, 2, n, True)
    for it in range(3):
        t = time.time(); got = ctx.msm(B, sc); dt = time.time() - t
        print(f"n=2^{sys.argv[1]} wall {dt*1e3:.2f} ms", ctx.last_timings(), flush=True)
    t = time.time(); exp = cref.commit(0, pts, sc); print("oracle", time.time() - t, "s")
    check("big msm", got, exp)
print("FAILS", fails)
sys.exit(1 if fails else 0)
