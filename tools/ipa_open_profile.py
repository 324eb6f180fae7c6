"""Per-round wall-clock and MSM stage breakdown of the IPA opening session (development aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import accumulation_b200 as ab
from accumulation_b200.mirror import _int_to_fe

k = int(sys.argv[1]) if len(sys.argv) > 1 else 18
ctx = ab.Context(0)
n = 1 << k
key = ctx.register_synthetic_bases(0, 1, n + 1)
if "--table" in sys.argv:
    key.precompute()
if "--no-fold" in sys.argv:
    ctx.set_ipa_fold(0, 0)
for a in sys.argv:
    if a.startswith("--fold="):
        ctx.set_ipa_fold(*[int(v) for v in a[7:].split(",")])
rng = np.random.default_rng(1)
coeffs = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64); coeffs[:, 3] &= np.uint64((1 << 62) - 1)
z = coeffs[0].copy()
hp = ctx.download_bases(key, n, 1).reshape(8)
xi = _int_to_fe(1, 0x1234567890abcdef1234567890abcdef)
xinv = _int_to_fe(1, pow(0x1234567890abcdef1234567890abcdef, -1, 0x40000000000000000000000000000000224698FC0994A8DD8C46EB2100000001))
for rep in range(2):
    t0 = time.perf_counter()
    s = ctx.ipa_open_begin(key, coeffs, k, z, hp)
    tb = time.perf_counter() - t0
    rows = []
    for r in range(k):
        t1 = time.perf_counter(); ctx.ipa_open_round(s); t2 = time.perf_counter()
        st = ctx.last_timings()
        ctx.ipa_open_fold(s, xi, xinv)
        # force the fold to finish so that it is attributed to this round
        import ctypes
        t3 = time.perf_counter()
        rows.append((r, (t2 - t1) * 1e3, st, (t3 - t2) * 1e3))
    t4 = time.perf_counter(); ctx.ipa_open_finish(s); t5 = time.perf_counter()
    if rep == 1:
        print(f"begin {tb*1e3:.2f} ms; finish(+last fold) {(t5-t4)*1e3:.2f} ms; total {(t5-t0)*1e3:.2f} ms")
        for r, ms, st, fms in rows:
            print(f"round {r:2d} n={n>>r:8d} round_call(prev fold + ip + 2 msm) {ms:7.3f} ms  fold_call {fms:6.3f} ms ", {a: round(b, 3) for a, b in st.items() if b})
