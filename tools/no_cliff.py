"""MSM time for the degenerate scalar distributions of the reference's fixtures vs uniform ('no cliff')."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import accumulation_b200 as ab

k = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << k
ctx = ab.Context(0)
key = ctx.register_synthetic_bases(0, 1, n)
if "--plain" not in sys.argv:
    key.precompute()
rng = np.random.default_rng(1)
uni = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64); uni[:, 3] &= np.uint64((1 << 62) - 1)
dists = {"uniform": uni, "constant": np.repeat(uni[:1], n, axis=0), "a_a_a_0": np.concatenate([np.repeat(uni[:1], n - 1, axis=0), np.zeros((1, 4), np.uint64)]),
         "zero": np.zeros((n, 4), np.uint64), "one": np.tile(np.array([1, 0, 0, 0], np.uint64), (n, 1)),
         "trunc128": np.concatenate([uni[:, :2], np.zeros((n, 2), np.uint64)], axis=1), "two_values": uni[rng.integers(0, 2, n)]}
base = None
for name, sc in dists.items():
    ts = []
    for _ in range(4):
        t0 = time.perf_counter(); ctx.msm(key, sc, montgomery=False); ts.append((time.perf_counter() - t0) * 1e3)
    st = ctx.last_timings()
    t = min(ts)
    base = base or t
    print(f"{name:10s} {t:8.3f} ms  x{t / base:5.2f}  " + " ".join(f"{a}={b:.3f}" for a, b in st.items() if b > 0.0005))
