// Single-warp latency of the field / group primitives (what bounds the tail kernels: k_finish, k_reduce*, the
// small rounds of the IPA opening).  1, 4, 16 warps per SM, clock64() around a dependent chain.
// Finding (profiles/r01c_latency.jsonl): one warp issues a 255-bit product every ~664 cycles whether or not
// independent products are available to it (2 or 4 independent chains, also with their carry chains interleaved
// by hand, take exactly 2x / 4x as long), while 4+ warps per scheduler reach ~250 cycles per product.  Tail
// kernels therefore gain from more independent warps, not from instruction-level parallelism inside a thread.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../accumulation_b200/csrc/ec.cuh"
using namespace accmsm;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int MODE> __global__ void k_lat(uint32_t *out, long long *cycles, int iters) {
    using F = Fp<0>; using Cv = Curve<0>;
    fe_t a[4], b;
    for (int k = 0; k < 4; k++) for (int i = 0; i < 8; i++) a[k].l[i] = out[(k * 8 + i) & 31] + threadIdx.x * 3 + k;
    for (int i = 0; i < 8; i++) b.l[i] = out[i + 3] | 1;
    for (int k = 0; k < 4; k++) a[k].l[7] &= 0x3fffffff;
    b.l[7] &= 0x3fffffff;
    xyzz_t p; p.x = a[0]; p.y = a[1]; p.zz = a[2]; p.zzz = a[3];
    xyzz_t q; q.x = a[1]; q.y = a[2]; q.zz = a[3]; q.zzz = b;
    affine_t ap; ap.x = a[2]; ap.y = b;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) a[0] = F::mul(a[0], b);
        if (MODE == 1) { a[0] = F::mul(a[0], b); a[1] = F::mul(a[1], b); }
        if (MODE == 2) { a[0] = F::mul(a[0], b); a[1] = F::mul(a[1], b); a[2] = F::mul(a[2], b); a[3] = F::mul(a[3], b); }
        if (MODE == 3) p = Cv::dbl(p);
        if (MODE == 4) Cv::add(p, q);
        if (MODE == 5) Cv::madd(p, ap);
        if (MODE == 6) a[0] = F::add(a[0], b);
        if (MODE == 7) a[0] = F::sub(a[0], b);
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int k = 0; k < 4; k++) for (int i = 0; i < 8; i++) s ^= a[k].l[i] ^ p.x.l[i] ^ p.y.l[i] ^ p.zz.l[i] ^ p.zzz.l[i];
    out[64 + blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE> int run(const char *name, int ops, uint32_t *out, long long *cyc, int warps) {
    int iters = 2000;
    k_lat<MODE><<<148, 32 * warps>>>(out, cyc, iters);
    CK(cudaDeviceSynchronize());
    k_lat<MODE><<<148, 32 * warps>>>(out, cyc, iters);
    CK(cudaDeviceSynchronize());
    long long h[148]; CK(cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
    printf("{\"bench\": \"%s\", \"warps_per_sm\": %d, \"cycles_per_iter\": %.1f, \"cycles_per_op\": %.1f}\n", name, warps, avg / iters, avg / iters / ops);
    return 0;
}
int main() {
    uint32_t *out; long long *cyc;
    CK(cudaMalloc(&out, (64 + 148 * 256) * 4)); CK(cudaMemset(out, 0x5a, (64 + 148 * 256) * 4)); CK(cudaMalloc(&cyc, 148 * 8));
    for (int w = 1; w <= 16; w *= 4) {
        run<0>("mul_chain_x1", 1, out, cyc, w);
        run<1>("mul_chain_x2", 2, out, cyc, w);
        run<2>("mul_chain_x4", 4, out, cyc, w);
        run<3>("xyzz_dbl", 1, out, cyc, w);
        run<4>("xyzz_add", 1, out, cyc, w);
        run<5>("xyzz_madd", 1, out, cyc, w);
        run<6>("fe_add", 1, out, cyc, w);
        run<7>("fe_sub", 1, out, cyc, w);
    }
    return 0;
}
