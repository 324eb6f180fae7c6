"""Development tool: where the end-to-end (host-scalar) MSM loses time against the device-resident one.

For every setting of the segment knobs (ACCMSM_SEG_PCTS, cumulative boundaries in percent) it times accmsm_msm from a
page-locked host buffer against the device-resident MSM and prints the library's segment timeline (ACCMSM_TRACE) of one call.
One process per setting (the knobs are read at init).

    python tools/e2e_segments.py            # sweep, one subprocess per setting
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(log_n: int, steps: int):
    import numpy as np
    import torch
    import accumulation_b200 as ab
    sys.path.insert(0, ROOT)
    from bench import rand_scalars, SEED

    ctx = ab.Context(0)
    n = 1 << log_n
    key = ctx.register_synthetic_bases(ab.PALLAS, SEED, n + 1, first_index=0)
    key.precompute(0)
    sc = rand_scalars(n, SEED + 1)
    h = torch.empty((n, 4), dtype=torch.int64).pin_memory()
    h_np = h.numpy().view(np.uint64)
    h_np[:] = sc
    d = h.to("cuda:0")
    for _ in range(3):
        ref = ctx.msm_dev(key, d.data_ptr(), n, montgomery=False) if hasattr(ctx, "msm_dev") else None
        got = ctx.msm(key, h_np, montgomery=False, n=n)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        got = ctx.msm(key, h_np, montgomery=False, n=n)
    t1 = time.perf_counter()
    out = {"e2e_ms": round((t1 - t0) / steps * 1e3, 4), "stages": {k: round(v, 4) for k, v in ctx.last_timings().items()}}
    if ref is not None:
        out["equal_to_device_resident"] = bool(np.array_equal(got[0], ref[0]) and got[1] == ref[1])
        t0 = time.perf_counter()
        for _ in range(steps):
            ctx.msm_dev(key, d.data_ptr(), n, montgomery=False)
        out["resident_ms"] = round((time.perf_counter() - t0) / steps * 1e3, 4)
    print(json.dumps(out), flush=True)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(int(sys.argv[2]), int(sys.argv[3]))
        return
    settings = [s for s in (sys.argv[1:] or ["12", "25", "12,62", "10,45", "25,75", "6,30,70"])]
    for pcts in settings:
        env = dict(os.environ, ACCMSM_SEG_PCTS=pcts)
        r = subprocess.run([sys.executable, __file__, "--child", "20", "30"], env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        print(f"pcts={pcts}: {line[-1] if line else r.stderr[-400:]}", flush=True)
        env = dict(os.environ, ACCMSM_SEG_PCTS=pcts, ACCMSM_TRACE="1")
        r = subprocess.run([sys.executable, __file__, "--child", "20", "1"], env=env, capture_output=True, text=True)
        tr = [l for l in r.stderr.splitlines() if "accmsm trace" in l]
        k = max(1, len(tr) // 4)
        print("\n".join(tr[-k:]), flush=True)


if __name__ == "__main__":
    main()
