// Test-only host build of the device headers (carry flag emulated), so the limb-level algorithms
// can be checked against the oracle on a box without a GPU.  Never part of the shipped library.
#include <cstring>
#include "../../accumulation_b200/csrc/fp.cuh"
using namespace accmsm;

template <int F> static void binop(int op, const uint32_t *a, const uint32_t *b, uint32_t *o, size_t n) {
    for (size_t i = 0; i < n; i++) {
        fe_t x, y, r;
        memcpy(x.l, a + 8 * i, 32);
        if (b) memcpy(y.l, b + 8 * i, 32);
        switch (op) {
            case 0: r = Fp<F>::mul(x, y); break;
            case 1: r = Fp<F>::add(x, y); break;
            case 2: r = Fp<F>::sub(x, y); break;
            case 3: r = Fp<F>::sqr(x); break;
            case 4: r = Fp<F>::inv(x); break;
            case 5: r = Fp<F>::from_mont(x); break;
            case 6: r = Fp<F>::to_mont(x); break;
            case 7: r = Fp<F>::neg(x); break;
            case 8: r = Fp<F>::inv_gcd(x); break;
            case 9: { fe_t t; bool ok = Fp<F>::sqrt(x, t); r = ok ? Fp<F>::mul(t, t) : Fp<F>::zero(); if (!ok) r.l[0] = 0xdeadbeefu; break; }   // sqrt(x)^2, or a marker
            default: r = x;
        }
        memcpy(o + 8 * i, r.l, 32);
    }
}
extern "C" void host_fe_op(int field, int op, const uint32_t *a, const uint32_t *b, uint32_t *o, size_t n) {
    if (field == 0) binop<0>(op, a, b, o, n); else binop<1>(op, a, b, o, n);
}

template <int F> static void mul2op(int sub, const uint32_t *a, const uint32_t *b, const uint32_t *c, const uint32_t *d, uint32_t *o, size_t n) {
    for (size_t i = 0; i < n; i++) {
        fe_t x, y, z, w;
        memcpy(x.l, a + 8 * i, 32); memcpy(y.l, b + 8 * i, 32); memcpy(z.l, c + 8 * i, 32); memcpy(w.l, d + 8 * i, 32);
        fe_t r = sub ? Fp<F>::mul2sub(x, y, z, w) : Fp<F>::mul2(x, y, z, w);
        memcpy(o + 8 * i, r.l, 32);
    }
}
extern "C" void host_fe_mul2(int field, int sub, const uint32_t *a, const uint32_t *b, const uint32_t *c, const uint32_t *d, uint32_t *o, size_t n) {
    if (field == 0) mul2op<0>(sub, a, b, c, d, o, n); else mul2op<1>(sub, a, b, c, d, o, n);
}

#include "../../accumulation_b200/csrc/ec.cuh"
template <int C> static void ec_sum(const uint32_t *xy, const uint8_t *neg, size_t n, int mode, uint32_t *out_xy, uint8_t *out_inf) {
    using Cv = Curve<C>;
    xyzz_t acc = Cv::identity(), acc2 = Cv::identity();
    for (size_t i = 0; i < n; i++) {
        affine_t p; memcpy(&p, xy + 16 * i, 64);
        if (neg && neg[i]) p = Cv::neg(p);
        if (mode == 0) Cv::madd(acc, p);                       // mixed adds
        else if (mode == 1) { xyzz_t q = Cv::from_affine(p); Cv::add(acc, q); }   // full adds
        else { if (i & 1) Cv::madd(acc2, p); else Cv::madd(acc, p); }             // two partials, then merged
    }
    if (mode == 2) Cv::add(acc, acc2);
    if (mode == 3) { acc = Cv::identity(); affine_t p; memcpy(&p, xy, 64); Cv::madd(acc, p); for (size_t i = 0; i < n; i++) acc = Cv::dbl(acc); }
    affine_t o; uint32_t inf; Cv::to_affine(acc, o, inf);
    memcpy(out_xy, &o, 64); *out_inf = (uint8_t)inf;
}
extern "C" void host_ec_sum(int curve, const uint32_t *xy, const uint8_t *neg, size_t n, int mode, uint32_t *out_xy, uint8_t *out_inf) {
    if (curve == 0) ec_sum<0>(xy, neg, n, mode, out_xy, out_inf); else ec_sum<1>(xy, neg, n, mode, out_xy, out_inf);
}

// the library's host-side 4 x 64-bit field code (hostfp.hpp): products, inversion, XYZZ -> affine with one inversion
#include "../../accumulation_b200/csrc/hostfp.hpp"
extern "C" void host_fast_op(int field, int op, const uint64_t *a, const uint64_t *b, uint64_t *o, size_t n) {
    const hostfp::Modulus &M = hostfp::modulus(field);
    for (size_t i = 0; i < n; i++) {
        if (op == 0) hostfp::mul(M, a + 4 * i, b + 4 * i, o + 4 * i);
        else hostfp::inv(M, a + 4 * i, o + 4 * i);
    }
}
extern "C" void host_fast_xyzz_to_affine(int field, const uint64_t *raw, size_t k, uint64_t *out_xy, uint8_t *out_inf) {
    hostfp::xyzz_to_affine(field, raw, k, out_xy, out_inf);
}
