"""Development tool: latency of the window-table build (accmsm_precompute_bases) for short keys."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import accumulation_b200 as ab
ctx = ab.Context(0)
for lg in (2, 6, 10, 13, 14, 15):
    key = ctx.register_synthetic_bases(0, 7, 1 << lg)
    ts = []
    for _ in range(4):
        t0 = time.perf_counter(); key.precompute(); ts.append((time.perf_counter() - t0) * 1e3)
    print(f"n=2^{lg}: precompute {min(ts):.3f} ms (coop={'off' if os.environ.get('ACCMSM_NO_COOP_PRECOMPUTE') else 'on'})", flush=True)
    key.release()
