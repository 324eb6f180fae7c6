// Integer-pipe micro-benchmark for the B200 roofline denominators that MEASURED_PEAKS.json lacks
// (SURVEY.md 8d asks for a measured IMAD rate).  Prints one JSON line per instruction mix:
// warp-instructions per clock per SM, from clock64() deltas inside the kernel and CUDA events.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../accumulation_b200/csrc/fp.cuh"
using namespace accmsm;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int ITERS = 4096;
constexpr int ACCS = 8;

template <int MODE> __global__ void __launch_bounds__(256) k_bench(uint32_t *out, long long *cycles, uint32_t x, uint32_t y) {
    uint32_t a[ACCS], c[ACCS]; uint64_t w[ACCS];
    for (int i = 0; i < ACCS; i++) { a[i] = threadIdx.x * 7 + i; c[i] = a[i] ^ 0x55; w[i] = a[i] * 0x100000001ull; }
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ACCS; i++) {
            if (MODE == 0) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(x), "r"(y));
            if (MODE == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(x), "r"(y));
            if (MODE == 2) asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0; }" : "+l"(w[i]) : "r"(x));
            if (MODE == 3) asm volatile("add.u32 %0, %0, %1; xor.b32 %0, %0, %2;" : "+r"(a[i]) : "r"(x), "r"(y));
            if (MODE == 4) { asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %0; mad.wide.u32 %0, lo, %1, %0; }" : "+l"(w[i]) : "r"(x));
                             asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(y)); }
            if (MODE == 5) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(x), "r"(y));
                             asm volatile("add.u32 %0, %0, %1;" : "+r"(a[(i + 1) % ACCS]) : "r"(y)); }
        }
        if (MODE == 6) {  // 4-long IMAD.WIDE.U32.X carry chains, two independent
            asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1;"
                         "madc.lo.cc.u32 %2, %8, %10, %2; madc.hi.cc.u32 %3, %8, %10, %3;"
                         "madc.lo.cc.u32 %4, %8, %9, %4; madc.hi.cc.u32 %5, %8, %9, %5;"
                         "madc.lo.cc.u32 %6, %8, %10, %6; madc.hi.u32 %7, %8, %10, %7;"
                         : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7])
                         : "r"(x), "r"(y), "r"(x ^ y));
            asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1;"
                         "madc.lo.cc.u32 %2, %8, %10, %2; madc.hi.cc.u32 %3, %8, %10, %3;"
                         "madc.lo.cc.u32 %4, %8, %9, %4; madc.hi.cc.u32 %5, %8, %9, %5;"
                         "madc.lo.cc.u32 %6, %8, %10, %6; madc.hi.u32 %7, %8, %10, %7;"
                         : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]), "+r"(c[7])
                         : "r"(y), "r"(x), "r"(x ^ y));
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < ACCS; i++) s ^= a[i] ^ c[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// field multiplication throughput: 4 independent dependent-chains per thread
template <int F> __global__ void __launch_bounds__(256) k_femul(uint32_t *out, long long *cycles, int iters) {
    fe_t a[2], b;
    for (int k = 0; k < 2; k++) for (int i = 0; i < 8; i++) a[k].l[i] = threadIdx.x * 31 + i + k * 977;
    for (int i = 0; i < 8; i++) b.l[i] = blockIdx.x * 17 + i;
    a[0].l[7] &= 0x3fffffff; a[1].l[7] &= 0x3fffffff; b.l[7] &= 0x3fffffff;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        a[0] = Fp<F>::mul(a[0], b);
        a[1] = Fp<F>::mul(a[1], b);
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < 8; i++) s ^= a[0].l[i] ^ a[1].l[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE> int run(const char *name, int instr_per_iter, uint32_t *out, long long *cyc, int sms) {
    int blocks = sms * 4, threads = 256;  // 32 warps / SM
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_bench<MODE><<<blocks, threads>>>(out, cyc, 3, 5);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    k_bench<MODE><<<blocks, threads>>>(out, cyc, 3, 5);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[2048]; CK(cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
    double winstr_per_sm = 32.0 * (double)ITERS * instr_per_iter;   // warp-instructions issued per SM
    printf("{\"bench\": \"%s\", \"warp_instr_per_clk_per_sm\": %.3f, \"thread_ops_per_clk_per_sm\": %.1f, \"ms\": %.4f, \"cycles\": %.0f, \"eff_mhz\": %.0f}\n",
           name, winstr_per_sm / avg, 32.0 * winstr_per_sm / avg, ms, avg, avg / (ms * 1e3));
    return 0;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.clockRate);
    uint32_t *out; long long *cyc;
    CK(cudaMalloc(&out, sms * 8 * 256 * 4)); CK(cudaMalloc(&cyc, 4096 * 8));
    run<0>("imad_lo", ACCS, out, cyc, sms);
    run<1>("imad_hi", ACCS, out, cyc, sms);
    run<2>("imad_wide", ACCS, out, cyc, sms);
    run<3>("iadd3+lop3", 2 * ACCS, out, cyc, sms);
    run<4>("imad_wide+iadd", 2 * ACCS, out, cyc, sms);
    run<5>("imad_lo+iadd", 2 * ACCS, out, cyc, sms);
    run<6>("imad_wide_x_chain", 8, out, cyc, sms);
    for (int f = 0; f < 2; f++) {
        for (int wpsm = 8; wpsm <= 64; wpsm *= 2) {
            int threads = 256, blocks = sms * wpsm / 8, iters = 16384;
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            if (f == 0) k_femul<0><<<blocks, threads>>>(out, cyc, iters); else k_femul<1><<<blocks, threads>>>(out, cyc, iters);
            CK(cudaDeviceSynchronize());
            cudaEventRecord(e0);
            if (f == 0) k_femul<0><<<blocks, threads>>>(out, cyc, iters); else k_femul<1><<<blocks, threads>>>(out, cyc, iters);
            cudaEventRecord(e1);
            CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double muls = (double)blocks * threads * iters * 2;
            long long h[4096]; CK(cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost));
            double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
            double warp_muls_per_smsp = (double)wpsm / 4 * iters * 2;
            printf("{\"bench\": \"fe_mul_field%d\", \"warps_per_sm\": %d, \"gmul_per_s\": %.2f, \"ms\": %.4f, \"cycles_per_warp_mul_per_smsp\": %.1f, \"eff_mhz\": %.0f}\n", f, wpsm, muls / ms / 1e6, ms, avg / warp_muls_per_smsp, avg / (ms * 1e3));
        }
    }
    return 0;
}
