// Group law for the Pallas/Vesta cycle (y^2 = x^3 + 5, a = 0, prime order) in XYZZ coordinates:
// x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2, identity <=> ZZ == 0.  Replaces ark-ec 0.2
// `short_weierstrass_jacobian::{GroupAffine, GroupProjective}` on the device.  Any correct addition law
// gives the same affine result, so bucket sums are bit-exact with ark-ec's Jacobian path after
// normalisation.  Formulas: EFD madd-2008-s / add-2008-s / dbl-2008-s-1 / mdbl-2008-s-1.
#pragma once
#include "fp.cuh"

namespace accmsm {

struct alignas(16) affine_t { fe_t x, y; };            // 64 B, Montgomery coordinates, never the identity
struct alignas(16) xyzz_t { fe_t x, y, zz, zzz; };     // 128 B

// CURVE 0 = Pallas (coordinates in Fp = field 0), CURVE 1 = Vesta (coordinates in Fq = field 1)
// FT selects the field implementation: Fp (products inlined; throughput kernels) or FpCall (products behind a
// call; latency-bound tail kernels, see fp.cuh).
template <int CURVE, template <int> class FT = Fp> struct Curve {
    using F = FT<CURVE == 0 ? 0 : 1>;

    static ACC_HD xyzz_t identity() {
        xyzz_t r;
        r.x = F::zero(); r.y = F::one(); r.zz = F::zero(); r.zzz = F::zero();
        return r;
    }
    static ACC_HD bool is_identity(const xyzz_t &p) { return F::is_zero(p.zz); }
    static ACC_HD xyzz_t from_affine(const affine_t &p) {
        xyzz_t r;
        r.x = p.x; r.y = p.y; r.zz = F::one(); r.zzz = F::one();
        return r;
    }
    static ACC_HD affine_t neg(const affine_t &p) {
        affine_t r;
        r.x = p.x; r.y = F::neg(p.y);
        return r;
    }

    // 2 * (affine p) -> XYZZ   (mdbl-2008-s-1, a = 0)
    static ACC_HD xyzz_t dbl_affine(const affine_t &p) {
        xyzz_t r;
        fe_t u = F::dbl(p.y);
        fe_t v = F::sqr(u);
        fe_t w = F::mul(u, v);
        fe_t s = F::mul(p.x, v);
        fe_t xx = F::sqr(p.x);
        fe_t m = F::add(F::dbl(xx), xx);
        r.x = F::sub(F::sub(F::sqr(m), s), s);
        r.y = F::mul2sub(m, F::sub(s, r.x), w, p.y);
        r.zz = v; r.zzz = w;
        return r;
    }
    // 2 * p   (dbl-2008-s-1, a = 0); identity and y == 0 cannot occur on a prime-order curve except identity
    static ACC_HD xyzz_t dbl(const xyzz_t &p) {
        if (is_identity(p)) return p;
        xyzz_t r;
        fe_t u = F::dbl(p.y);
        fe_t v = F::sqr(u);
        fe_t w = F::mul(u, v);
        fe_t s = F::mul(p.x, v);
        fe_t xx = F::sqr(p.x);
        fe_t m = F::add(F::dbl(xx), xx);
        r.x = F::sub(F::sub(F::sqr(m), s), s);
        r.y = F::mul2sub(m, F::sub(s, r.x), w, p.y);
        r.zz = F::mul(v, p.zz);
        r.zzz = F::mul(w, p.zzz);
        return r;
    }

    // acc += p (p affine, not the identity)   madd-2008-s with the exceptional cases peeled:
    // acc empty -> load; same point -> doubling; opposite point -> identity.
    static ACC_HD void madd(xyzz_t &acc, const affine_t &p) {
        if (is_identity(acc)) { acc = from_affine(p); return; }
        fe_t u2 = F::mul(p.x, acc.zz);
        fe_t s2 = F::mul(p.y, acc.zzz);
        fe_t pp_ = F::sub(u2, acc.x);   // P
        fe_t r = F::sub(s2, acc.y);     // R
        if (F::is_zero(pp_)) {
            if (F::is_zero(r)) acc = dbl_affine(p); else acc = identity();
            return;
        }
        fe_t pp = F::sqr(pp_);
        fe_t ppp = F::mul(pp_, pp);
        fe_t q = F::mul(acc.x, pp);
        fe_t x3 = F::sub(F::sub(F::sub(F::sqr(r), ppp), q), q);
        fe_t y3 = F::mul2sub(r, F::sub(q, x3), acc.y, ppp);
        acc.zz = F::mul(acc.zz, pp);
        acc.zzz = F::mul(acc.zzz, ppp);
        acc.x = x3; acc.y = y3;
    }

    // acc += q   add-2008-s with the exceptional cases peeled
    static ACC_HD void add(xyzz_t &acc, const xyzz_t &q) {
        if (is_identity(q)) return;
        if (is_identity(acc)) { acc = q; return; }
        fe_t u1 = F::mul(acc.x, q.zz);
        fe_t u2 = F::mul(q.x, acc.zz);
        fe_t s1 = F::mul(acc.y, q.zzz);
        fe_t s2 = F::mul(q.y, acc.zzz);
        fe_t pp_ = F::sub(u2, u1);
        fe_t r = F::sub(s2, s1);
        if (F::is_zero(pp_)) {
            if (F::is_zero(r)) acc = dbl(acc); else acc = identity();
            return;
        }
        fe_t pp = F::sqr(pp_);
        fe_t ppp = F::mul(pp_, pp);
        fe_t qq = F::mul(u1, pp);
        fe_t x3 = F::sub(F::sub(F::sub(F::sqr(r), ppp), qq), qq);
        fe_t y3 = F::mul2sub(r, F::sub(qq, x3), s1, ppp);
        acc.zz = F::mul(F::mul(acc.zz, q.zz), pp);
        acc.zzz = F::mul(F::mul(acc.zzz, q.zzz), ppp);
        acc.x = x3; acc.y = y3;
    }

    // Normalise to ark-ec's affine image: (x, y, infinity); the identity is (0, 1, true).
    // With t = ZZZ^-1:  ZZ^-1 = ZZ^2 t^2 (since ZZ^3 = ZZZ^2), so x = X ZZ^2 t^2, y = Y t.
    // GCD = true: binary-GCD inversion (single-thread tails); false: Fermat ladder (no divergence across lanes)
    template <bool GCD = false>
    static ACC_HD void to_affine(const xyzz_t &p, affine_t &out, uint32_t &inf) {
        if (is_identity(p)) { out.x = F::zero(); out.y = F::one(); inf = 1; return; }
        fe_t t = GCD ? F::inv_gcd(p.zzz) : F::inv(p.zzz);
        fe_t zt = F::mul(p.zz, t);
        out.x = F::mul(p.x, F::sqr(zt));
        out.y = F::mul(p.y, t);
        inf = 0;
    }
};

}  // namespace accmsm
