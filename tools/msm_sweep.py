"""BASELINE config 5: standalone Pallas / Vesta variable-base MSM sweep 2^12 .. 2^24 on one GPU vs the host CPU
restatement of ark-ec VariableBaseMSM.  One JSON line per (curve, log n): device-resident time (CUDA events inside the
library), end-to-end time from pinned host scalars, registration (upload + window table) time, CPU time, bit-exact
check against the oracle (CPU leg only up to --cpu-max).  Development / reporting aid: bench.py is the contract."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import accumulation_b200 as ab

ap = argparse.ArgumentParser()
ap.add_argument("--min", type=int, default=12); ap.add_argument("--max", type=int, default=24)
ap.add_argument("--cpu-max", type=int, default=24); ap.add_argument("--curves", default="0,1")
args = ap.parse_args()
ctx = ab.Context(0)
for curve in [int(c) for c in args.curves.split(",")]:
    for k in range(args.min, args.max + 1):
        n = 1 << k
        t0 = time.perf_counter(); key = ctx.register_synthetic_bases(curve, 0xACC5, n); t_gen = time.perf_counter() - t0
        t0 = time.perf_counter(); key.precompute(); t_pre = time.perf_counter() - t0
        rng = np.random.default_rng(k)
        sc = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64); sc[:, 3] &= np.uint64((1 << 62) - 1)
        h = torch.empty((n, 4), dtype=torch.int64).pin_memory(); h.numpy().view(np.uint64)[:] = sc
        d = h.cuda()
        reps = 5 if k <= 22 else 3
        dev, e2e = [], []
        for _ in range(reps):
            t0 = time.perf_counter(); got = ctx.msm_dev(key, d.data_ptr(), n, montgomery=False); dev.append((time.perf_counter() - t0) * 1e3)
            st = ctx.last_timings()
        for _ in range(reps):
            t0 = time.perf_counter(); got2 = ctx.msm_ptr(key, h.data_ptr(), n, montgomery=False); e2e.append((time.perf_counter() - t0) * 1e3)
        rec = {"curve": "pallas" if curve == 0 else "vesta", "log_n": k, "gpu_ms": round(min(dev), 4), "gpu_mpts": round(n / min(dev) / 1e3, 2),
               "e2e_ms": round(min(e2e), 4), "e2e_mpts": round(n / min(e2e) / 1e3, 2), "register_ms": round((t_pre) * 1e3, 2),
               "stages_ms": {a: round(b, 4) for a, b in st.items() if b > 0.0005}}
        # size-independent check at every size: MSM(bases[0, n)) == MSM(bases[0, n/2)) + MSM(bases[n/2, n)), the halves through the
        # offset / n arguments of the same entry point and the final addition on the CPU (a two-term MSM with scalars 1, 1)
        from oracle import cref
        h1 = ctx.msm_dev(key, d.data_ptr(), n // 2, montgomery=False, offset=0)
        h2 = ctx.msm_dev(key, d.data_ptr() + (n // 2) * 32, n - n // 2, montgomery=False, offset=n // 2)
        parts = [p for p in (h1, h2) if not p[1]]
        one = np.array([[1, 0, 0, 0]] * len(parts), dtype=np.uint64)
        ssum = cref.msm_ark(curve, np.array([p[0] for p in parts]), one) if parts else None
        rec["split_sum_equal"] = bool(ssum is not None and ssum[1] == got[1] and np.array_equal(ssum[0], got[0]) and np.array_equal(got2[0], got[0]))
        if k <= args.cpu_max:
            from oracle import cref
            pts = ctx.download_bases(key)
            t0 = time.perf_counter(); exp = cref.msm_ark(curve, pts, sc); t_cpu = time.perf_counter() - t0
            rec.update({"cpu_ms": round(t_cpu * 1e3, 2), "cpu_mpts": round(n / t_cpu / 1e6, 3), "cpu_threads": cref.num_threads(),
                        "speedup_e2e": round(t_cpu * 1e3 / min(e2e), 1),
                        "bit_exact": bool(got[1] == exp[1] and np.array_equal(got[0], exp[0]) and np.array_equal(got2[0], exp[0]))})
        print(json.dumps(rec), flush=True)
        key.release(); del d, h
