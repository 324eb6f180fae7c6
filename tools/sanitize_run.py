"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): every kernel family once at small sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import accumulation_b200 as ab

ctx = ab.Context(0)
rng = np.random.default_rng(3)
def scal(n):
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64); a[:, 3] &= np.uint64((1 << 62) - 1); return a
for curve in (0, 1):
    n = 3000
    key = ctx.register_synthetic_bases(curve, 5, n + 1)
    sf = ab.scalar_field(curve)
    for pre in (False, True):
        if pre:
            key.precompute(10)
        ctx.msm(key, scal(n), montgomery=False)
        ctx.msm(key, np.repeat(scal(1), n, axis=0))
        ctx.msm(key, scal(33), offset=7)
        ctx.commit(key, scal(n), hiding_index=n, randomizer_mont=scal(1)[0])
        ctx.msm_batch(key, scal(3 * 500).reshape(3, 500, 4))
        ctx.ipa_final_key(key, scal(11))
    pts = ctx.download_bases(key, 0, 64)
    ctx.msm_oneshot(curve, pts, scal(64), montgomery=False)
    ctx.msm_oneshot_batch(curve, ctx.download_bases(key, 0, 11 * 43).reshape(11, 43, 8), scal(11 * 43).reshape(11, 43, 4), montgomery=False)   # two passes of <= 8 jobs
    # wire formats: compressed ark-serialize image -> decompression kernel (Tonelli-Shanks) -> the same key
    blob = ctx.serialize_bases(key, 0, 200)
    k2 = ctx.register_bases_compressed(curve, blob); assert (ctx.download_bases(k2) == ctx.download_bases(key, 0, 200)).all(); k2.release()
    # long pass: the radix sort staged through shared memory (sort.cuh; >= 2^17 (bucket, point) pairs) on a table key
    big = ctx.register_synthetic_bases(curve, 6, 20000); big.precompute(10)
    r1 = ctx.msm(big, scal(20000), montgomery=False)
    ctx.msm(big, np.repeat(scal(1), 20000, axis=0), montgomery=False)       # constant vector: the skew gate hands over to the counting sort
    import torch
    d_part = torch.zeros(16, dtype=torch.int64, device="cuda")
    ctx.msm_partial(big, scal(20000), d_part.data_ptr(), montgomery=False)
    ctx.combine_partials_dev(curve, d_part.data_ptr(), 1)
    big.release()
    if curve == 0 and os.environ.get("SANITIZE_LARGE"):
        # host-scalar MSM of >= 16 MiB: two point segments, the second uploaded behind the accumulation of the first (k_accumulate `into` mode)
        huge = ctx.register_synthetic_bases(0, 9, 1 << 19); huge.precompute()
        sc = scal(1 << 19)
        got = ctx.msm(huge, sc, montgomery=False)
        d = torch.from_numpy(sc.view(np.int64)).cuda()
        ref = ctx.msm_dev(huge, d.data_ptr(), 1 << 19, montgomery=False)
        assert got[1] == ref[1] and (got[0] == ref[0]).all()
        huge.release()
    a, b = scal(n), scal(n)
    ctx.hadamard(sf, a, b); ctx.scale(sf, a, b[0]); ctx.lincomb(sf, [a, b[:100]], scal(2), a[:50])
    ctx.tvecs(sf, [a, b], [b, a], scal(3), n, a, b)
    ctx.compute_coeffs(sf, scal(8)); ctx.combine_check_polys(sf, scal(16).reshape(2, 8, 4), scal(2), scal(2)); ctx.poly_evaluate(sf, a, b[0])
    prod = ctx.hadamard(sf, a, b)
    exp = [ctx.commit(key, v, hiding_index=n, randomizer_mont=r) for v, r in zip((a, b, prod), scal(3))]
    ctx.hp_decide(key, a, b, np.array([e[0] for e in exp]), [e[1] for e in exp], hiding_index=n, randomness=scal(3))
    ctx.hp_product_poly_comm(key, [a, b], [b, a], scal(3), n)
    row_ptr = np.arange(n + 1, dtype=np.uint32); cols = rng.integers(0, n, n).astype(np.uint32)
    mats = [(row_ptr, cols, scal(n))] * 3
    ctx.csr_matvec(sf, mats, a[:6], a[6:])
    h = ctx.register_csr(sf, mats); ctx.csr_matvec_commit(key, h, 3, n, a[:6], a[6:], hiding_index=n, blinders=scal(3)); ctx.release_csr(h)
    k = 6
    hp = ctx.download_bases(key, n, 1).reshape(8)
    s = ctx.ipa_open_begin(key, scal(1 << k), k, scal(1)[0], hp)
    xi = np.array([3, 0, 0, 0], dtype=np.uint64)
    for _ in range(k):
        ctx.ipa_open_round(s); ctx.ipa_open_fold(s, xi, xi)
    ctx.ipa_open_finish(s)
    ctx.set_ipa_fold(2, 2)          # materialise the folded key after rounds 2 and 4 (k_fold_* in ipa.cuh)
    s = ctx.ipa_open_begin(key, scal(1 << k), k, scal(1)[0], hp)
    for _ in range(k):
        ctx.ipa_open_round(s); ctx.ipa_open_fold(s, xi, xi)
    ctx.ipa_open_finish(s)
    ctx.set_ipa_fold()
    key.release()
ctx.close()
print("sanitize_run done")
