// Burst micro-benchmark of the 255-bit field product in the regime k_accumulate runs in:
// 256-thread CTAs, 8 / 16 / 32 warps per SM, bursts of a few ms (so the part stays at its boost clock instead of the
// ~1 GHz it power-caps to under a 100 ms saturating IMAD load), effective clock recorded per point.
// Prints one JSON line per (kernel, warps per SM): products per second, cycles per warp-product per SM sub-partition, MHz.
//   fe_mul  : two independent chains of Fp::mul per thread (pure multiplier throughput)
//   fe_sqr  : the same with Fp::sqr
//   madd    : XYZZ mixed additions acc += P_j (10 products each: 7 mul + 2 sqr + 1 dual product), the loop body of
//             k_accumulate without its memory traffic -- the arithmetic ceiling of that kernel
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../accumulation_b200/csrc/ec.cuh"
using namespace accmsm;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int MODE> __global__ void __launch_bounds__(256) k_fe(uint32_t *out, long long *cycles, int iters) {
    fe_t a[2], b;
    for (int k = 0; k < 2; k++) for (int i = 0; i < 8; i++) a[k].l[i] = threadIdx.x * 31 + i + k * 977;
    for (int i = 0; i < 8; i++) b.l[i] = blockIdx.x * 17 + i;
    a[0].l[7] &= 0x3fffffff; a[1].l[7] &= 0x3fffffff; b.l[7] &= 0x3fffffff;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) { a[0] = Fp<0>::mul(a[0], b); a[1] = Fp<0>::mul(a[1], b); }
        else { a[0] = Fp<0>::sqr(a[0]); a[1] = Fp<0>::sqr(a[1]); }
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < 8; i++) s ^= a[0].l[i] ^ a[1].l[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(256, 2) k_madd(uint32_t *out, long long *cycles, int iters) {
    using Cv = Curve<0>;
    affine_t g;
    g.x = Cv::F::neg(Cv::F::one()); g.y = Cv::F::dbl(Cv::F::one());      // (-1, 2)
    xyzz_t acc = Cv::dbl_affine(g);
    for (uint32_t k = 0; k < (threadIdx.x & 7u); k++) acc = Cv::dbl(acc);  // lanes start from different points
    affine_t p = g;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        Cv::madd(acc, p);
        p.y = Cv::F::neg(p.y);                 // alternate +G / -G: acc oscillates between 2^k G and 2^k G + G, never P == +-Q
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < 8; i++) s ^= acc.x.l[i] ^ acc.y.l[i] ^ acc.zz.l[i] ^ acc.zzz.l[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.clockRate);
    uint32_t *out; long long *cyc;
    CK(cudaMalloc(&out, (size_t)sms * 8 * 256 * 4)); CK(cudaMalloc(&cyc, 4096 * 8));
    static long long h[4096];
    const char *names[3] = {"fe_mul", "fe_sqr", "madd"};
    for (int mode = 0; mode < 3; mode++) {
        for (int wpsm = 8; wpsm <= 32; wpsm *= 2) {
            if (mode == 2 && wpsm > 16) continue;       // 128 registers: at most 16 warps per SM
            const int threads = 256, blocks = sms * wpsm / 8;
            // ~2.5 ms per burst at the expected rates
            const int iters = mode == 2 ? (wpsm == 8 ? 700 : 350) : (wpsm == 8 ? 2400 : wpsm == 16 ? 1200 : 600);
            const double per_iter = mode == 2 ? 1.0 : 2.0;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            float best = 1e30f; double cyc_best = 0;
            for (int rep = 0; rep < 6; rep++) {
                CK(cudaDeviceSynchronize());
                cudaEventRecord(e0);
                if (mode == 0) k_fe<0><<<blocks, threads>>>(out, cyc, iters);
                else if (mode == 1) k_fe<1><<<blocks, threads>>>(out, cyc, iters);
                else k_madd<<<blocks, threads>>>(out, cyc, iters);
                cudaEventRecord(e1);
                CK(cudaDeviceSynchronize());
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                CK(cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost));
                double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
                if (rep > 0 && ms < best) { best = ms; cyc_best = avg; }
            }
            const double ops = (double)blocks * threads * iters * per_iter;
            const double warp_ops_per_smsp = (double)wpsm / 4 * iters * per_iter;
            const double gops = ops / best / 1e6;
            printf("{\"bench\": \"%s\", \"warps_per_sm\": %d, \"g_per_s\": %.2f, \"gmul_equiv_per_s\": %.2f, \"ms\": %.4f, "
                   "\"cycles_per_warp_op_per_smsp\": %.1f, \"eff_mhz\": %.0f}\n",
                   names[mode], wpsm, gops, mode == 2 ? gops * 10 : gops, best, cyc_best / warp_ops_per_smsp, cyc_best / (best * 1e3));
        }
    }
    return 0;
}
