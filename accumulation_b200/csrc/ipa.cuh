// Device side of IpaPC::open (ark-poly-commit ipa_pc; reference call sites
// src/ipa_pc_as/mod.rs:454-462 (AS prove), :525-534 (index: default proof), examples/scaling-pc.rs:72-81).
// The opening state -- coefficient vector and z-vector (1, z, z^2, ...) -- stays resident in HBM across the log2(D)
// rounds; per round only the two points (l, r) go to the host sponge and one challenge comes back.
//   round:  l = cm_commit(key_l, coeffs_r) + <coeffs_r, z_l> h'     r = cm_commit(key_r, coeffs_l) + <coeffs_l, z_r> h'
//   fold:   coeffs_l += xi^-1 coeffs_r ;  z_l += xi z_r ;  key_l += xi key_r
// The key is never folded explicitly: the folded key of round j is a fixed linear combination of the registered key
// (coefficients = products of the challenges so far), so every round's (l, r) is an MSM over the REGISTERED key with
// scalars generated on the fly (IpaRoundScalars in msm.cuh) -- the window table applies, there is no Horner over
// windows and no per-round scalar multiplication of n/2 generators; final_comm_key is one more such MSM.
#pragma once
#include "vec.cuh"

namespace accmsm {

// z_vec[i] = z^i by square-and-multiply on the index (one thread per element)
template <int FIELD>
__global__ void __launch_bounds__(256) k_powers(const uint8_t *__restrict__ z_ptr, uint32_t n, uint8_t *__restrict__ out) {
    using F = Fp<FIELD>;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe_t zc = load_fe(z_ptr), pw = F::one();
    for (uint32_t e = i; e; e >>= 1) {
        if (e & 1u) pw = F::mul(pw, zc);
        zc = F::sqr(zc);
    }
    store_fe(out + (size_t)i * 32, pw);
}

// blockIdx.y = 0: <coeffs[h..2h), z[0..h)>   blockIdx.y = 1: <coeffs[0..h), z[h..2h)>; per-block partial sums
template <int FIELD>
__global__ void __launch_bounds__(256) k_ipa_inner_partial(const uint8_t *__restrict__ coeffs, const uint8_t *__restrict__ z,
                                                            uint32_t h, uint8_t *__restrict__ partials) {
    using F = Fp<FIELD>;
    __shared__ fe_t sh[256];
    const uint8_t *a = blockIdx.y == 0 ? coeffs + (size_t)h * 32 : coeffs;
    const uint8_t *b = blockIdx.y == 0 ? z : z + (size_t)h * 32;
    fe_t acc = F::zero();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < h; i += gridDim.x * blockDim.x)
        acc = F::add(acc, F::mul(load_fe_nc(a + (size_t)i * 32), load_fe_nc(b + (size_t)i * 32)));
    fe_t s = block_sum<FIELD>(acc, sh);
    if (threadIdx.x == 0) store_fe(partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 32, s);
}
// grid = 2 blocks: out[y] = canonical(scale * sum of partials[y][0..nparts)), scale = 1 when nullptr  -- canonical
// because it feeds either the byte loop of k_ipa_hp_mul or the digit decomposition of the MSM
template <int FIELD>
__global__ void __launch_bounds__(256) k_ipa_inner_final(const uint8_t *__restrict__ partials, uint32_t nparts,
                                                          const uint8_t *__restrict__ scale, uint8_t *__restrict__ out_canon) {
    using F = Fp<FIELD>;
    __shared__ fe_t sh[256];
    fe_t acc = F::zero();
    for (uint32_t i = threadIdx.x; i < nparts; i += blockDim.x)
        acc = F::add(acc, load_fe(partials + ((size_t)blockIdx.x * nparts + i) * 32));
    fe_t s = block_sum<FIELD>(acc, sh);
    if (threadIdx.x == 0) {
        if (scale) s = F::mul(s, load_fe(scale));
        store_fe(out_canon + (size_t)blockIdx.x * 32, F::from_mont(s));
    }
}

// hp_table[j] = 2^(8 j) h' (XYZZ), j < 32: one thread, once per opening session
template <int CURVE>
__global__ void k_ipa_hp_table(const affine_t *__restrict__ hp, xyzz_t *__restrict__ table) {
    using Cv = Curve<CURVE>;
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    xyzz_t cur = Cv::from_affine(load_affine(hp));
    for (int j = 0; j < 32; j++) {
        store_xyzz(table + j, cur);
        if (j < 31) for (int b = 0; b < 8; b++) cur = Cv::dbl(cur);
    }
}
// out[y] = ip[y] * h' for y = 0, 1 (one warp each): lane j multiplies byte j of the scalar into 2^(8j) h',
// then the 32 partial points are summed by a shuffle tree.
template <int CURVE>
__global__ void __launch_bounds__(64) k_ipa_hp_mul(const xyzz_t *__restrict__ table, const uint8_t *__restrict__ ip_canon,
                                                    xyzz_t *__restrict__ out) {
    using Cv = Curve<CURVE>;
    const uint32_t y = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t byte = ip_canon[y * 32 + lane];
    xyzz_t base = load_xyzz(table + lane);
    xyzz_t acc = Cv::identity();
#pragma unroll 1
    for (int b = 7; b >= 0; b--) {
        acc = Cv::dbl(acc);
        if ((byte >> b) & 1u) Cv::add(acc, base);
    }
#pragma unroll 1
    for (int d = 16; d >= 1; d >>= 1) {
        xyzz_t o = shfl_down_xyzz(acc, d);
        if (lane < (uint32_t)d) Cv::add(acc, o);
    }
    if (lane == 0) store_xyzz(out + y, acc);
}

// coeffs[i] += xi_inv * coeffs[i + h] ;  z[i] += xi * z[i + h]
template <int FIELD>
__global__ void __launch_bounds__(256) k_ipa_fold_scalars(uint8_t *__restrict__ coeffs, uint8_t *__restrict__ z, uint32_t h,
                                                           const uint8_t *__restrict__ xi, const uint8_t *__restrict__ xi_inv) {
    using F = Fp<FIELD>;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= h) return;
    fe_t x = load_fe(xi), xinv = load_fe(xi_inv);
    uint8_t *c0 = coeffs + (size_t)i * 32, *z0 = z + (size_t)i * 32;
    store_fe(c0, F::add(load_fe(c0), F::mul(xinv, load_fe(coeffs + (size_t)(i + h) * 32))));
    store_fe(z0, F::add(load_fe(z0), F::mul(x, load_fe(z + (size_t)(i + h) * 32))));
}

}  // namespace accmsm
