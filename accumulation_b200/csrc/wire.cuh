// ark-serialize 0.2 wire images on the device: a commitment key that arrives in its canonical
// compressed form -- what `CanonicalSerialize` writes for `Vec<GroupAffine<P>>` minus the 8-byte length prefix -- is
// decompressed straight into the HBM-resident key, one thread per point (one Tonelli-Shanks square root each: ~750 field
// products, data-parallel, no host arithmetic).
//   field element        32 B  little-endian canonical integer (into_repr)
//   compressed SW point  33 B  x as above, then one flag byte: bit 7 = "PositiveY" (y is the larger of y, -y as canonical
//                              integers), bit 6 = point at infinity (x = 0); both set is invalid
// (ark-ff 0.2 `serialize_with_flags`: buffer_byte_size(255 + 2 flag bits) = 33; ark-ec 0.2 short_weierstrass_jacobian
// `GroupAffine::{serialize, deserialize}` with `SWFlags`; `get_point_from_x(x, greatest)` picks the root.)
#pragma once
#include "msm.cuh"

namespace accmsm {

constexpr uint32_t WIRE_POINT_BYTES = 33;

ACC_D fe_t wire_load_fe(const uint8_t *p) {          // unaligned little-endian bytes -> limbs
    fe_t r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) | ((uint32_t)p[4 * i + 3] << 24);
    return r;
}
ACC_D void wire_store_fe(uint8_t *p, const fe_t &r) {
#pragma unroll
    for (int i = 0; i < 8; i++) { p[4 * i] = (uint8_t)r.l[i]; p[4 * i + 1] = (uint8_t)(r.l[i] >> 8); p[4 * i + 2] = (uint8_t)(r.l[i] >> 16); p[4 * i + 3] = (uint8_t)(r.l[i] >> 24); }
}

// in: n x 33 B.  out: Montgomery x || y records + identity bytes; *n_bad counts invalid encodings (x not canonical,
// x^3 + 5 not a square, contradictory flags) -- ark-serialize answers those with SerializationError::InvalidData.
template <int CURVE>
__global__ void __launch_bounds__(128) k_wire_decompress(const uint8_t *__restrict__ in, uint32_t n, affine_t *__restrict__ out,
                                                          uint8_t *__restrict__ out_inf, uint32_t *__restrict__ n_bad, uint32_t *__restrict__ n_inf) {
    using F = typename Curve<CURVE>::F;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *p = in + (size_t)i * WIRE_POINT_BYTES;
    const fe_t xc = wire_load_fe(p);
    const uint32_t flags = p[32];
    const bool positive = (flags >> 7) & 1u, infinity = (flags >> 6) & 1u;
    affine_t r;
    r.x = F::zero(); r.y = F::one();
    bool bad = (flags & 0x3fu) != 0 || (positive && infinity) || !F::is_canonical(xc);
    if (!bad && infinity) { bad = !F::is_zero(xc); atomicAdd(n_inf, 1u); }
    if (!bad && !infinity) {
        const fe_t x = F::to_mont(xc);
        fe_t five = F::zero(); five.l[0] = 5u;
        const fe_t rhs = F::add(F::mul(F::sqr(x), x), F::to_mont(five));
        fe_t y;
        if (!F::sqrt(rhs, y)) bad = true;
        else {
            const fe_t ny = F::neg(y);
            const bool y_is_larger = F::lt(F::from_mont(ny), F::from_mont(y));
            r.x = x; r.y = (y_is_larger == positive) ? y : ny;
        }
    }
    if (bad) atomicAdd(n_bad, 1u);
    store_fe(&out[i].x, r.x); store_fe(&out[i].y, r.y);
    out_inf[i] = (!bad && infinity) ? 1 : 0;
}

// the inverse: registered bases -> n x 33 B
template <int CURVE>
__global__ void __launch_bounds__(128) k_wire_compress(const affine_t *__restrict__ in, const uint8_t *__restrict__ in_inf, uint32_t n,
                                                        uint8_t *__restrict__ out) {
    using F = typename Curve<CURVE>::F;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t *p = out + (size_t)i * WIRE_POINT_BYTES;
    if (in_inf && in_inf[i]) { wire_store_fe(p, F::zero()); p[32] = 1u << 6; return; }
    const fe_t x = load_fe(&in[i].x), y = load_fe(&in[i].y);
    wire_store_fe(p, F::from_mont(x));
    p[32] = F::lt(F::from_mont(F::neg(y)), F::from_mont(y)) ? (1u << 7) : 0u;
}

}  // namespace accmsm
