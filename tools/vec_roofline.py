"""HBM roofline of the field-vector kernels (K3 materialised, K4, K5; SURVEY.md 8d gives the algorithmic bytes per unit):
kernel-only time from the library's CUDA events (stage 'vec_kernel'), achieved GB/s against the measured copy bandwidth."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import accumulation_b200 as ab

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]); KIND = "measured"
except Exception:
    PEAK, KIND = 6650.0, "fallback"
ctx = ab.Context(0)
rng = np.random.default_rng(1)
def scal(n):
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64); a[:, 3] &= np.uint64((1 << 62) - 1); return a
def run(name, alg_bytes, fn, reps=5):
    best = 1e9
    for _ in range(reps):
        fn(); best = min(best, ctx.last_timings()["vec_kernel"])
    gbs = alg_bytes / (best * 1e-3) / 1e9
    print(json.dumps({"kernel": name, "algorithmic_bytes": alg_bytes, "kernel_ms": round(best, 4), "achieved_GBps": round(gbs, 1),
                      "peak_GBps": PEAK, "peak_kind": KIND, "frac": round(gbs / PEAK, 3)}), flush=True)
F = 1
for k in (16, 20, 22):
    n = 1 << k
    a, b, c, d = scal(n), scal(n), scal(n), scal(n)
    run(f"k_hadamard len=2^{k}", 96 * n, lambda: ctx.hadamard(F, a, b))
    run(f"k_scale len=2^{k}", 64 * n, lambda: ctx.scale(F, a, b[0]))
    run(f"k_lincomb m=4 len=2^{k}", 32 * 5 * n, lambda: ctx.lincomb(F, [a, b, c, d], scal(4)))
    if k <= 20:
        run(f"k_tvecs n=2 len=2^{k} (3 outputs)", 32 * (4 + 3) * n, lambda: ctx.tvecs(F, [a, b], [c, d], scal(3), n))
    run(f"k_compute_coeffs k={k}", 32 * n, lambda: ctx.compute_coeffs(F, scal(k)))
    run(f"k_combine_check_polys m=2 k={k}", 32 * n, lambda: ctx.combine_check_polys(F, scal(2 * k).reshape(2, k, 4), scal(2), scal(2)))
    run(f"k_poly_eval len=2^{k}", 32 * n, lambda: ctx.poly_evaluate(F, a, b[0]))
    for nnz in (1, 8):
        row_ptr = (np.arange(n + 1, dtype=np.uint64) * nnz).astype(np.uint32)
        cols = rng.integers(0, n, n * nnz).astype(np.uint32)
        coeffs = scal(n * nnz)
        mats = [(row_ptr, cols, coeffs)] * 3
        alg = 3 * (n * nnz * (32 + 4 + 32) + n * (4 + 32))
        if k <= 20:
            run(f"k_csr_matvec x3 rows=2^{k} nnz/row={nnz}", alg, lambda: ctx.csr_matvec(F, mats, a[:6], a[6:]), reps=3)
