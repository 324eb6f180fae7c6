// Sustained-load probe: runs one instruction mix for ~3 s while nvidia-smi samples clocks/power.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include "../accumulation_b200/csrc/fp.cuh"
using namespace accmsm;
template <int MODE> __global__ void __launch_bounds__(256) k_load(uint32_t *out, long long *cycles, int iters) {
    fe_t a[2], b;
    for (int k = 0; k < 2; k++) for (int i = 0; i < 8; i++) a[k].l[i] = threadIdx.x * 31 + i + k * 977;
    for (int i = 0; i < 8; i++) b.l[i] = blockIdx.x * 17 + i;
    a[0].l[7] &= 0x3fffffff; a[1].l[7] &= 0x3fffffff; b.l[7] &= 0x3fffffff;
    uint32_t x = b.l[0] | 1, y = b.l[1]; uint32_t c[8]; for (int i = 0; i < 8; i++) c[i] = x + i;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) { a[0] = Fp<0>::mul(a[0], b); a[1] = Fp<0>::mul(a[1], b); }
        if (MODE == 1) {  // IMAD lo only, 16 independent
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[0].l[i]) : "r"(x), "r"(y));
                                              asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[1].l[i]) : "r"(x), "r"(y)); }
        }
        if (MODE == 2) {  // IADD3/LOP3 only
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) { asm volatile("add.u32 %0, %0, %1; xor.b32 %0, %0, %2;" : "+r"(a[0].l[i]) : "r"(x), "r"(y));
                                              asm volatile("add.u32 %0, %0, %1; xor.b32 %0, %0, %2;" : "+r"(a[1].l[i]) : "r"(x), "r"(y)); }
        }
        if (MODE == 3) {  // IMAD.WIDE.X chains
#pragma unroll
            for (int r = 0; r < 8; r++) {
            asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1;"
                         "madc.lo.cc.u32 %2, %8, %10, %2; madc.hi.cc.u32 %3, %8, %10, %3;"
                         "madc.lo.cc.u32 %4, %8, %9, %4; madc.hi.cc.u32 %5, %8, %9, %5;"
                         "madc.lo.cc.u32 %6, %8, %10, %6; madc.hi.u32 %7, %8, %10, %7;"
                         : "+r"(a[0].l[0]), "+r"(a[0].l[1]), "+r"(a[0].l[2]), "+r"(a[0].l[3]), "+r"(a[0].l[4]), "+r"(a[0].l[5]), "+r"(a[0].l[6]), "+r"(a[0].l[7])
                         : "r"(x), "r"(y), "r"(x ^ y));
            asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1;"
                         "madc.lo.cc.u32 %2, %8, %10, %2; madc.hi.cc.u32 %3, %8, %10, %3;"
                         "madc.lo.cc.u32 %4, %8, %9, %4; madc.hi.cc.u32 %5, %8, %9, %5;"
                         "madc.lo.cc.u32 %6, %8, %10, %6; madc.hi.u32 %7, %8, %10, %7;"
                         : "+r"(a[1].l[0]), "+r"(a[1].l[1]), "+r"(a[1].l[2]), "+r"(a[1].l[3]), "+r"(a[1].l[4]), "+r"(a[1].l[5]), "+r"(a[1].l[6]), "+r"(a[1].l[7])
                         : "r"(y), "r"(x), "r"(x ^ y));
            }
        }
        if (MODE == 4 || MODE == 5 || MODE == 6) {  // IMAD.WIDE.X chains interleaved with independent IADD3
#pragma unroll
            for (int r = 0; r < 8; r++) {
            asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1;"
                         "madc.lo.cc.u32 %2, %8, %10, %2; madc.hi.cc.u32 %3, %8, %10, %3;"
                         "madc.lo.cc.u32 %4, %8, %9, %4; madc.hi.cc.u32 %5, %8, %9, %5;"
                         "madc.lo.cc.u32 %6, %8, %10, %6; madc.hi.u32 %7, %8, %10, %7;"
                         : "+r"(a[0].l[0]), "+r"(a[0].l[1]), "+r"(a[0].l[2]), "+r"(a[0].l[3]), "+r"(a[0].l[4]), "+r"(a[0].l[5]), "+r"(a[0].l[6]), "+r"(a[0].l[7])
                         : "r"(x), "r"(y), "r"(x ^ y));
                if (MODE == 4) {
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[0]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[1]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[2]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[3]) : "r"(y));
                }
                if (MODE == 5) {
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[0]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[1]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[2]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[3]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[4]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[5]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[6]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[7]) : "r"(y));
                }
                if (MODE == 6) {
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[0]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[1]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[2]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[3]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[4]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[5]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[6]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[7]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[0]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[1]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[2]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[3]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[4]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[5]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[6]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[7]) : "r"(y));
                }
            asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1;"
                         "madc.lo.cc.u32 %2, %8, %10, %2; madc.hi.cc.u32 %3, %8, %10, %3;"
                         "madc.lo.cc.u32 %4, %8, %9, %4; madc.hi.cc.u32 %5, %8, %9, %5;"
                         "madc.lo.cc.u32 %6, %8, %10, %6; madc.hi.u32 %7, %8, %10, %7;"
                         : "+r"(a[1].l[0]), "+r"(a[1].l[1]), "+r"(a[1].l[2]), "+r"(a[1].l[3]), "+r"(a[1].l[4]), "+r"(a[1].l[5]), "+r"(a[1].l[6]), "+r"(a[1].l[7])
                         : "r"(x), "r"(y), "r"(x ^ y));
                if (MODE == 4) {
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[0]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[1]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[2]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[3]) : "r"(y));
                }
                if (MODE == 5) {
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[0]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[1]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[2]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[3]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[4]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[5]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[6]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[7]) : "r"(y));
                }
                if (MODE == 6) {
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[0]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[1]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[2]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[3]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[4]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[5]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[6]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[7]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[0]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[1]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[2]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[3]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[4]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[5]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[6]) : "r"(y));
            asm volatile("add.u32 %0, %0, %1;" : "+r"(c[7]) : "r"(y));
                }
            }
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < 8; i++) s ^= a[0].l[i] ^ a[1].l[i] ^ c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
int main(int argc, char **argv) {
    int mode = argc > 1 ? atoi(argv[1]) : 0, wpsm = argc > 2 ? atoi(argv[2]) : 32;
    double secs = argc > 3 ? atof(argv[3]) : 3.0;
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount, blocks = sms * wpsm / 8, iters = 16384;
    uint32_t *out; long long *cyc; cudaMalloc(&out, blocks * 256 * 4); cudaMalloc(&cyc, blocks * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    double total_ms = 0; int launches = 0; double cyc_sum = 0;
    while (total_ms < secs * 1e3) {
        cudaEventRecord(e0);
        switch (mode) { case 0: k_load<0><<<blocks, 256>>>(out, cyc, iters); break; case 1: k_load<1><<<blocks, 256>>>(out, cyc, iters); break;
                        case 2: k_load<2><<<blocks, 256>>>(out, cyc, iters); break; case 3: k_load<3><<<blocks, 256>>>(out, cyc, iters); break; case 4: k_load<4><<<blocks, 256>>>(out, cyc, iters); break; case 5: k_load<5><<<blocks, 256>>>(out, cyc, iters); break; default: k_load<6><<<blocks, 256>>>(out, cyc, iters); }
        cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1); total_ms += ms; launches++;
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); cyc_sum += h;
    }
    printf("{\"mode\": %d, \"warps_per_sm\": %d, \"launches\": %d, \"ms_per_launch\": %.3f, \"cycles_per_launch\": %.0f, \"eff_mhz\": %.0f}\n",
           mode, wpsm, launches, total_ms / launches, cyc_sum / launches, cyc_sum / launches / (total_ms / launches * 1e3));
    return 0;
}
