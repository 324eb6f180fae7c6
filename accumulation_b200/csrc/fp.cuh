// 255-bit Montgomery field arithmetic on 8 x 32-bit limbs for the Pallas/Vesta cycle (K1 of
// SURVEY.md 2b).  Replaces ark-ff 0.2 `Fp256<P>` (4 x u64 Montgomery, R = 2^256) on the device: the
// memory image is identical (little-endian limbs), so field elements cross the C-ABI untouched.
//
// Both moduli are 2^254 + t with limbs [1, m1, m2, m3, 0, 0, 0, 2^30] and -m^{-1} mod 2^32 = -1
// (SURVEY.md App. B): the Montgomery quotient digit is a negation and each reduction round needs
// three real limb products.
//
// Carry chains are written as one PTX instruction per `asm volatile` statement (add.cc / madc.*),
// which ptxas fuses into IMAD.WIDE + IADD3.X; the same source compiles for the host with an emulated
// carry flag so tests can check the arithmetic without a GPU (tests/host/…).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define ACC_HD __host__ __device__ __forceinline__
#define ACC_D __device__ __forceinline__
#else
#define ACC_HD inline
#define ACC_D inline
#endif

namespace accmsm {

// ---------------------------------------------------------------------------------------------
// carry-chain primitives
// ---------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
#define ACC_ASM_R1(name, ptx)                                                       \
    ACC_D uint32_t name(uint32_t a, uint32_t b) {                                   \
        uint32_t r;                                                                 \
        asm volatile(ptx " %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));                \
        return r;                                                                   \
    }
#define ACC_ASM_R3(name, ptx)                                                       \
    ACC_D uint32_t name(uint32_t a, uint32_t b, uint32_t c) {                       \
        uint32_t r;                                                                 \
        asm volatile(ptx " %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));    \
        return r;                                                                   \
    }
ACC_ASM_R1(add_cc, "add.cc.u32")
ACC_ASM_R1(addc_cc, "addc.cc.u32")
ACC_ASM_R1(addc, "addc.u32")
ACC_ASM_R1(sub_cc, "sub.cc.u32")
ACC_ASM_R1(subc_cc, "subc.cc.u32")
ACC_ASM_R1(subc, "subc.u32")
ACC_ASM_R3(mad_lo_cc, "mad.lo.cc.u32")
ACC_ASM_R3(madc_lo_cc, "madc.lo.cc.u32")
ACC_ASM_R3(mad_hi_cc, "mad.hi.cc.u32")
ACC_ASM_R3(madc_hi_cc, "madc.hi.cc.u32")
ACC_ASM_R3(madc_hi, "madc.hi.u32")
ACC_ASM_R3(madc_lo, "madc.lo.u32")
ACC_D uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
ACC_D uint32_t mul_hi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
// 64-bit (register-pair) forms: `mul.wide.u32` feeding `add.cc.u64` is fused by ptxas into a single
// IMAD.WIDE.U32(.X) with carry-in/out predicates; keeping the accumulators as 64-bit virtual registers
// makes the pair alignment explicit, which is what lets the whole multiplier stay on wide IMADs.
ACC_D uint64_t mulw(uint32_t a, uint32_t b) { uint64_t r; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r; }
ACC_D uint64_t add_cc64(uint64_t a, uint64_t b) { uint64_t r; asm volatile("add.cc.u64 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
ACC_D uint64_t addc_cc64(uint64_t a, uint64_t b) { uint64_t r; asm volatile("addc.cc.u64 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
ACC_D uint64_t addc64(uint64_t a, uint64_t b) { uint64_t r; asm volatile("addc.u64 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
ACC_D uint32_t lo32(uint64_t a) { uint32_t l, h; asm("mov.b64 {%0, %1}, %2;" : "=r"(l), "=r"(h) : "l"(a)); return l; }
ACC_D uint32_t hi32(uint64_t a) { uint32_t l, h; asm("mov.b64 {%0, %1}, %2;" : "=r"(l), "=r"(h) : "l"(a)); return h; }
ACC_D uint64_t pack64(uint32_t l, uint32_t h) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(l), "r"(h)); return r; }
#undef ACC_ASM_R1
#undef ACC_ASM_R3
#else
// Host emulation (unit tests only): one carry/borrow flag per thread, PTX semantics.
namespace hostcc { static thread_local uint32_t cf = 0; }
inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b; hostcc::cf = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b + hostcc::cf; hostcc::cf = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + hostcc::cf; }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t d = (uint64_t)a - b; hostcc::cf = (uint32_t)(d >> 63); return (uint32_t)d; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t d = (uint64_t)a - b - hostcc::cf; hostcc::cf = (uint32_t)(d >> 63); return (uint32_t)d; }
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - hostcc::cf; }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint64_t mulw(uint32_t a, uint32_t b) { return (uint64_t)a * b; }
inline uint64_t add_cc64(uint64_t a, uint64_t b) { unsigned __int128 s = (unsigned __int128)a + b; hostcc::cf = (uint32_t)(s >> 64); return (uint64_t)s; }
inline uint64_t addc_cc64(uint64_t a, uint64_t b) { unsigned __int128 s = (unsigned __int128)a + b + hostcc::cf; hostcc::cf = (uint32_t)(s >> 64); return (uint64_t)s; }
inline uint64_t addc64(uint64_t a, uint64_t b) { return a + b + hostcc::cf; }
inline uint32_t lo32(uint64_t a) { return (uint32_t)a; }
inline uint32_t hi32(uint64_t a) { return (uint32_t)(a >> 32); }
inline uint64_t pack64(uint32_t l, uint32_t h) { return ((uint64_t)h << 32) | l; }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_lo(a, b), c); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_lo(a, b), c); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return addc(mul_hi(a, b), c); }
inline uint32_t madc_lo(uint32_t a, uint32_t b, uint32_t c) { return addc(mul_lo(a, b), c); }
#endif

// ---------------------------------------------------------------------------------------------
// field parameters. FIELD 0 = Fp (Pallas base / Vesta scalar), FIELD 1 = Fq (Pallas scalar / Vesta base)
// ---------------------------------------------------------------------------------------------
template <int FIELD> struct FieldParams;
template <> struct FieldParams<0> {
    static constexpr uint32_t M1 = 0x992d30edu, M2 = 0x094cf91bu, M3 = 0x224698fcu;
    static constexpr uint32_t R[8] = {0xfffffffdu, 0x34786d38u, 0xe41914adu, 0x992c350bu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0x3fffffffu};
    static constexpr uint32_t R2[8] = {0x0000000fu, 0x8c78ecb3u, 0x8b0de0e7u, 0xd7d30dbdu, 0xc3c95d18u, 0x7797a99bu, 0x7b9cb714u, 0x096d41afu};
};
template <> struct FieldParams<1> {
    static constexpr uint32_t M1 = 0x8c46eb21u, M2 = 0x0994a8ddu, M3 = 0x224698fcu;
    static constexpr uint32_t R[8] = {0xfffffffdu, 0x5b2b3e9cu, 0xe3420567u, 0x992c350bu, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0x3fffffffu};
    static constexpr uint32_t R2[8] = {0x0000000fu, 0xfc9678ffu, 0x891a16e3u, 0x67bb433du, 0x04ccf590u, 0x7fae2310u, 0x7ccfdaa9u, 0x096d41afu};
};
constexpr uint32_t MOD_L0 = 1u, MOD_L7 = 0x40000000u;

// A zero the compiler cannot fold (constant bank, never written): multiplying by it turns a two-limb
// carry ripple into one IMAD.WIDE.U32.X on the multiplier pipe instead of two IADD3.X on the ALU pipe.
// ACC_MUL_WIDE_RIPPLES picks how many of the 3 ripples per reduction round take that route, to balance
// the two half-rate pipes (tools/powerprobe.cu measures both at 2 clk per warp instruction per SMSP).
#if defined(__CUDACC__)
static __constant__ uint32_t ACC_OPAQUE_ZERO;
#endif
// ACC_MUL_SHIFT_TOP = 1 computes q * 2^30 with two shifts on the ALU pipe instead of an IMAD.WIDE (96 -> 88 wide
// multiplies per product).  Measured slower inside k_accumulate (2.30 vs 2.19 ms at 2^20): the kernel is bound by
// issue slots of the whole instruction mix, not by the multiplier pipe alone.  Kept for the record, off.
#ifndef ACC_MUL_SHIFT_TOP
#define ACC_MUL_SHIFT_TOP 0
#endif
// ACC_SQR_DEDICATED = 1: Fp::sqr skips the symmetric limb products (36 instead of 64 IMAD.WIDE; the skipped ones become
// carry ripples on the ALU pipe).  In isolation that loses (tools/ubench2.cu, profiles/r02a_ubench2.jsonl: 73.7 G
// squarings/s against 78.8 G products/s at 16 warps per SM -- the part is power-limited under pure multiplier load), but
// inside k_accumulate, where the multiplier pipe is the busiest unit and the ALU pipe has slack, it wins: 2.109 vs
// 2.163 ms at 2^20 points (same box, same run, profiles/r02b_bench*.jsonl).  ON.
#ifndef ACC_SQR_DEDICATED
#define ACC_SQR_DEDICATED 1
#endif
#ifndef ACC_MUL_WIDE_RIPPLES
#define ACC_MUL_WIDE_RIPPLES 0
#endif
ACC_HD uint32_t opaque_zero() {
#if defined(__CUDA_ARCH__)
    return ACC_OPAQUE_ZERO;
#else
    return 0u;
#endif
}

// A field element in registers: 8 little-endian 32-bit limbs, Montgomery form, canonical (< m).
struct alignas(16) fe_t {
    uint32_t l[8];
};

template <int FIELD> struct Fp {
    using P = FieldParams<FIELD>;

    static ACC_HD uint32_t mod_limb(int i) {
        return i == 0 ? MOD_L0 : i == 1 ? P::M1 : i == 2 ? P::M2 : i == 3 ? P::M3 : i == 7 ? MOD_L7 : 0u;
    }
    static ACC_HD fe_t zero() { fe_t r; for (int i = 0; i < 8; i++) r.l[i] = 0; return r; }
    static ACC_HD fe_t one() {
        fe_t r;
        r.l[0] = 0xfffffffdu; r.l[1] = FIELD == 0 ? 0x34786d38u : 0x5b2b3e9cu;
        r.l[2] = FIELD == 0 ? 0xe41914adu : 0xe3420567u; r.l[3] = 0x992c350bu;
        r.l[4] = r.l[5] = r.l[6] = 0xffffffffu; r.l[7] = 0x3fffffffu;
        return r;
    }
    static ACC_HD fe_t r2() {
        fe_t r;
        if (FIELD == 0) {
            r.l[0] = 0x0000000fu; r.l[1] = 0x8c78ecb3u; r.l[2] = 0x8b0de0e7u; r.l[3] = 0xd7d30dbdu;
            r.l[4] = 0xc3c95d18u; r.l[5] = 0x7797a99bu; r.l[6] = 0x7b9cb714u; r.l[7] = 0x096d41afu;
        } else {
            r.l[0] = 0x0000000fu; r.l[1] = 0xfc9678ffu; r.l[2] = 0x891a16e3u; r.l[3] = 0x67bb433du;
            r.l[4] = 0x04ccf590u; r.l[5] = 0x7fae2310u; r.l[6] = 0x7ccfdaa9u; r.l[7] = 0x096d41afu;
        }
        return r;
    }
    static ACC_HD bool is_zero(const fe_t &a) {
        return (a.l[0] | a.l[1] | a.l[2] | a.l[3] | a.l[4] | a.l[5] | a.l[6] | a.l[7]) == 0;
    }
    static ACC_HD bool eq(const fe_t &a, const fe_t &b) {
        uint32_t d = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) d |= a.l[i] ^ b.l[i];
        return d == 0;
    }

    // r = a - m if a >= m else a   (a < 2m)
    static ACC_HD void reduce_once(fe_t &a) {
        uint32_t u[8];
        u[0] = sub_cc(a.l[0], MOD_L0);
        u[1] = subc_cc(a.l[1], P::M1);
        u[2] = subc_cc(a.l[2], P::M2);
        u[3] = subc_cc(a.l[3], P::M3);
        u[4] = subc_cc(a.l[4], 0);
        u[5] = subc_cc(a.l[5], 0);
        u[6] = subc_cc(a.l[6], 0);
        u[7] = subc_cc(a.l[7], MOD_L7);
        uint32_t borrow = subc(0, 0);  // 0xffffffff if a < m
#pragma unroll
        for (int i = 0; i < 8; i++) a.l[i] = borrow ? a.l[i] : u[i];
    }

    static ACC_HD fe_t add(const fe_t &a, const fe_t &b) {
        fe_t r;
        r.l[0] = add_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < 7; i++) r.l[i] = addc_cc(a.l[i], b.l[i]);
        r.l[7] = addc(a.l[7], b.l[7]);  // both < 2^255: no carry out
        reduce_once(r);
        return r;
    }
    static ACC_HD fe_t dbl(const fe_t &a) { return add(a, a); }
    static ACC_HD fe_t sub(const fe_t &a, const fe_t &b) {
        fe_t r;
        r.l[0] = sub_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < 8; i++) r.l[i] = subc_cc(a.l[i], b.l[i]);
        uint32_t mask = subc(0, 0);  // all ones if a < b
        r.l[0] = add_cc(r.l[0], mask & MOD_L0);
        r.l[1] = addc_cc(r.l[1], mask & P::M1);
        r.l[2] = addc_cc(r.l[2], mask & P::M2);
        r.l[3] = addc_cc(r.l[3], mask & P::M3);
        r.l[4] = addc_cc(r.l[4], 0);
        r.l[5] = addc_cc(r.l[5], 0);
        r.l[6] = addc_cc(r.l[6], 0);
        r.l[7] = addc(r.l[7], mask & MOD_L7);
        return r;
    }
    static ACC_HD fe_t neg(const fe_t &a) { return sub(zero(), a); }

    // Interleaved (CIOS) Montgomery product on two column-parity accumulators of 64-bit registers:
    //   V = EV + OD * 2^32,   EV = sum ev[k] 2^(64k) (k < 5),   OD = sum od[k] 2^(64k) (k < 4).
    // Products a[j]*b[i] with j even land 64-bit aligned in EV, j odd in OD, so every limb product is one
    // IMAD.WIDE.U32.X in a carry chain.  After each round V is divided by 2^32, which swaps the roles of
    // the two accumulators (pure renaming): EV' = OD + hi32(ev[0]), OD' = EV >> 64; the carry of that one
    // 32-bit add is handed to the OD' chain, whose first limb has the same weight (2^32).
    // Invariant: V < a + m after every round, so EV, OD < 2^256 and no chain overflows its top register.
    // The result is canonicalised by one conditional subtraction.
    static ACC_HD fe_t mul(const fe_t &A, const fe_t &B) {
        const uint32_t *a = A.l, *b = B.l;
        uint64_t ev[5], od[4];
#pragma unroll
        for (int s = 0; s < 4; s++) { ev[s] = mulw(a[2 * s], b[0]); od[s] = mulw(a[2 * s + 1], b[0]); }
        ev[4] = 0;
        const uint32_t k0 = opaque_zero();
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (i > 0) {
                uint64_t nev[5], nod[4];
                uint32_t t_lo = add_cc(lo32(od[0]), hi32(ev[0]));
                nod[0] = addc_cc64(ev[1], mulw(a[1], b[i]));
                nod[1] = addc_cc64(ev[2], mulw(a[3], b[i]));
                nod[2] = addc_cc64(ev[3], mulw(a[5], b[i]));
                nod[3] = addc64(ev[4], mulw(a[7], b[i]));
                nev[0] = add_cc64(pack64(t_lo, hi32(od[0])), mulw(a[0], b[i]));
                nev[1] = addc_cc64(od[1], mulw(a[2], b[i]));
                nev[2] = addc_cc64(od[2], mulw(a[4], b[i]));
                nev[3] = addc_cc64(od[3], mulw(a[6], b[i]));
                nev[4] = addc64(0, 0);
#pragma unroll
                for (int k = 0; k < 4; k++) { ev[k] = nev[k]; od[k] = nod[k]; }
                ev[4] = nev[4];
            }
            // V += q * m, q = -V mod 2^32, m = [1, M1, M2, M3, 0, 0, 0, 2^30]
            reduce_round(ev, od, k0);
        }
        return assemble(ev, od);
    }
    // One Montgomery reduction round on the two accumulators: V += q * m with q = -V mod 2^32 (the low word becomes 0).
    static ACC_HD void reduce_round(uint64_t (&ev)[5], uint64_t (&od)[4], uint32_t k0) {
        uint32_t q = 0u - lo32(ev[0]);
        ev[0] = add_cc64(ev[0], (uint64_t)q);
        ev[1] = addc_cc64(ev[1], mulw(q, P::M2));
        ev[2] = addc_cc64(ev[2], ACC_MUL_WIDE_RIPPLES >= 2 ? mulw(q, k0) : 0ull);
        ev[3] = addc_cc64(ev[3], ACC_MUL_WIDE_RIPPLES >= 3 ? mulw(q, k0) : 0ull);
        ev[4] = addc64(ev[4], 0);
        od[0] = add_cc64(od[0], mulw(q, P::M1));
        od[1] = addc_cc64(od[1], mulw(q, P::M3));
        od[2] = addc_cc64(od[2], ACC_MUL_WIDE_RIPPLES >= 1 ? mulw(q, k0) : 0ull);
        od[3] = addc64(od[3], mulw(q, MOD_L7));
    }
    // (EV + OD 2^32) / 2^32 with lo32(ev[0]) == 0, then one conditional subtraction (value < 2m)
    static ACC_HD fe_t assemble(const uint64_t (&ev)[5], const uint64_t (&od)[4]) {
        fe_t r;
        r.l[0] = add_cc(hi32(ev[0]), lo32(od[0]));
        r.l[1] = addc_cc(lo32(ev[1]), hi32(od[0]));
        r.l[2] = addc_cc(hi32(ev[1]), lo32(od[1]));
        r.l[3] = addc_cc(lo32(ev[2]), hi32(od[1]));
        r.l[4] = addc_cc(hi32(ev[2]), lo32(od[2]));
        r.l[5] = addc_cc(lo32(ev[3]), hi32(od[2]));
        r.l[6] = addc_cc(hi32(ev[3]), lo32(od[3]));
        r.l[7] = addc(lo32(ev[4]), hi32(od[3]));
        reduce_once(r);
        return r;
    }

    // Squaring: a^2 = sum_i a_i 2^(32 i) * (a_i 2^(32 i) + 2 A_{>i}) with A_{>i} the limbs above i.  2 A_{>i} has the limbs
    // [a_{i+1} << 1, d_{i+2}, .., d_7] where d = limbs of 2a (2a < 2^256), so row i of the interleaved product needs only
    // the 8 - i limb products at positions >= i; the skipped ones leave plain carry ripples.  36 instead of 64 limb
    // products, same reduction.  Bounds: the row operand is < 2a, so V < 2a + m < 2^256 after every round and the result
    // is < a^2 / R + m < 2m.
    static ACC_HD fe_t sqr(const fe_t &A) {
#if !ACC_SQR_DEDICATED
        return mul(A, A);
#endif
        const uint32_t *a = A.l;
        uint32_t d[8], s[8];      // d = 2a, s[j] = a[j] << 1 (d[j] with the bit shifted in from limb j - 1 cleared)
        d[0] = s[0] = a[0] << 1;
#pragma unroll
        for (int j = 1; j < 8; j++) { s[j] = a[j] << 1; d[j] = s[j] | (a[j - 1] >> 31); }
        uint64_t ev[5], od[4];
        // row 0: a_0 * [a_0, s_1, d_2, .., d_7]
        ev[0] = mulw(a[0], a[0]); od[0] = mulw(s[1], a[0]);
        ev[1] = mulw(d[2], a[0]); od[1] = mulw(d[3], a[0]);
        ev[2] = mulw(d[4], a[0]); od[2] = mulw(d[5], a[0]);
        ev[3] = mulw(d[6], a[0]); od[3] = mulw(d[7], a[0]);
        ev[4] = 0;
        const uint32_t k0 = opaque_zero();
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (i > 0) {
                // operand limb at position j of row i: 0 (j < i), a_i (j == i), s_{i+1} (j == i + 1), d_j (j > i + 1)
                auto x = [&](int j) -> uint64_t {
                    if (j < i) return 0ull;
                    return mulw(j == i ? a[i] : j == i + 1 ? s[j] : d[j], a[i]);
                };
                uint64_t nev[5], nod[4];
                uint32_t t_lo = add_cc(lo32(od[0]), hi32(ev[0]));
                nod[0] = addc_cc64(ev[1], x(1));
                nod[1] = addc_cc64(ev[2], x(3));
                nod[2] = addc_cc64(ev[3], x(5));
                nod[3] = addc64(ev[4], x(7));
                nev[0] = add_cc64(pack64(t_lo, hi32(od[0])), x(0));
                nev[1] = addc_cc64(od[1], x(2));
                nev[2] = addc_cc64(od[2], x(4));
                nev[3] = addc_cc64(od[3], x(6));
                nev[4] = addc64(0, 0);
#pragma unroll
                for (int k = 0; k < 4; k++) { ev[k] = nev[k]; od[k] = nod[k]; }
                ev[4] = nev[4];
            }
            reduce_round(ev, od, k0);
        }
        return assemble(ev, od);
    }

    // (a b + c d) R^-1 with ONE reduction: both product rows are accumulated in every round.  V < a + c + m < 2^256
    // after every round (inputs canonical), the result is < (a b + c d) / R + m < 2m.  96 + 64 limb products instead of
    // 2 x 96, one conditional subtraction instead of two plus a field addition.
    static ACC_HD fe_t mul2(const fe_t &A, const fe_t &B, const fe_t &C, const fe_t &D) {
        const uint32_t *a = A.l, *b = B.l, *c = C.l, *dd = D.l;
        uint64_t ev[5], od[4];
#pragma unroll
        for (int s = 0; s < 4; s++) { ev[s] = mulw(a[2 * s], b[0]); od[s] = mulw(a[2 * s + 1], b[0]); }
        ev[0] = add_cc64(ev[0], mulw(c[0], dd[0]));
        ev[1] = addc_cc64(ev[1], mulw(c[2], dd[0]));
        ev[2] = addc_cc64(ev[2], mulw(c[4], dd[0]));
        ev[3] = addc_cc64(ev[3], mulw(c[6], dd[0]));
        ev[4] = addc64(0, 0);
        od[0] = add_cc64(od[0], mulw(c[1], dd[0]));
        od[1] = addc_cc64(od[1], mulw(c[3], dd[0]));
        od[2] = addc_cc64(od[2], mulw(c[5], dd[0]));
        od[3] = addc64(od[3], mulw(c[7], dd[0]));
        const uint32_t k0 = opaque_zero();
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (i > 0) {
                uint64_t nev[5], nod[4];
                uint32_t t_lo = add_cc(lo32(od[0]), hi32(ev[0]));
                nod[0] = addc_cc64(ev[1], mulw(a[1], b[i]));
                nod[1] = addc_cc64(ev[2], mulw(a[3], b[i]));
                nod[2] = addc_cc64(ev[3], mulw(a[5], b[i]));
                nod[3] = addc64(ev[4], mulw(a[7], b[i]));
                nev[0] = add_cc64(pack64(t_lo, hi32(od[0])), mulw(a[0], b[i]));
                nev[1] = addc_cc64(od[1], mulw(a[2], b[i]));
                nev[2] = addc_cc64(od[2], mulw(a[4], b[i]));
                nev[3] = addc_cc64(od[3], mulw(a[6], b[i]));
                nev[4] = addc64(0, 0);
                nev[0] = add_cc64(nev[0], mulw(c[0], dd[i]));
                nev[1] = addc_cc64(nev[1], mulw(c[2], dd[i]));
                nev[2] = addc_cc64(nev[2], mulw(c[4], dd[i]));
                nev[3] = addc_cc64(nev[3], mulw(c[6], dd[i]));
                nev[4] = addc64(nev[4], 0);
                nod[0] = add_cc64(nod[0], mulw(c[1], dd[i]));
                nod[1] = addc_cc64(nod[1], mulw(c[3], dd[i]));
                nod[2] = addc_cc64(nod[2], mulw(c[5], dd[i]));
                nod[3] = addc64(nod[3], mulw(c[7], dd[i]));
#pragma unroll
                for (int k = 0; k < 4; k++) { ev[k] = nev[k]; od[k] = nod[k]; }
                ev[4] = nev[4];
            }
            reduce_round(ev, od, k0);
        }
        return assemble(ev, od);
    }
    // a b - c d
    static ACC_HD fe_t mul2sub(const fe_t &a, const fe_t &b, const fe_t &c, const fe_t &d) { return mul2(a, b, neg(c), d); }

    // into_repr(): Montgomery image -> canonical integer = a * 1 * R^-1
    static ACC_HD fe_t from_mont(const fe_t &a) {
        fe_t o = zero();
        o.l[0] = 1u;
        return mul(a, o);
    }
    static ACC_HD fe_t to_mont(const fe_t &a) { return mul(a, r2()); }

    // R^3 mod m (for inv_gcd: the inverse of the Montgomery image x = a R is a^-1 R^-1; times R^3 / R gives a^-1 R)
    static ACC_HD fe_t r3() {
        fe_t r;
        if (FIELD == 0) {
            r.l[0] = 0x3a9e10f9u; r.l[1] = 0xf185a599u; r.l[2] = 0x6ac5b1d1u; r.l[3] = 0xf6a68f3bu;
            r.l[4] = 0x353fd42cu; r.l[5] = 0xdf8d1014u; r.l[6] = 0x2d2d9910u; r.l[7] = 0x2ae30922u;
        } else {
            r.l[0] = 0x249dae4cu; r.l[1] = 0x008b421cu; r.l[2] = 0xdba41326u; r.l[3] = 0xe13bda50u;
            r.l[4] = 0x8e15cb63u; r.l[5] = 0x88fececbu; r.l[6] = 0x6e6792c8u; r.l[7] = 0x07dd97a0u;
        }
        return r;
    }
    // Inversion by the binary extended Euclidean algorithm (Stein): ~380 shift steps and ~180 subtractions on 8 limbs,
    // about a third of the ~254 squarings + ~64 products of the Fermat ladder below.  Data-dependent control flow: meant
    // for single-thread tails (k_finish normalises one point per MSM); kernels where every lane inverts its own value
    // keep inv() (no divergence).  inv_gcd(0) = 0.
    static ACC_HD fe_t inv_gcd(const fe_t &a) {
        if (is_zero(a)) return zero();
        fe_t u = a, v, b = zero(), c = zero();
#pragma unroll
        for (int i = 0; i < 8; i++) v.l[i] = mod_limb(i);
        b.l[0] = 1u;                                   // b x == u, c x == v (mod m) with x the integer image of a
        auto is_one = [](const fe_t &t) { return t.l[0] == 1u && (t.l[1] | t.l[2] | t.l[3] | t.l[4] | t.l[5] | t.l[6] | t.l[7]) == 0u; };
        auto shr1 = [](fe_t &t, uint32_t top) {
#pragma unroll
            for (int i = 0; i < 7; i++) t.l[i] = (t.l[i] >> 1) | (t.l[i + 1] << 31);
            t.l[7] = (t.l[7] >> 1) | (top << 31);
        };
        auto halve_mod = [&](fe_t &t) {               // t / 2 mod m for t < m
            uint32_t carry = 0;
            if (t.l[0] & 1u) {
                t.l[0] = add_cc(t.l[0], MOD_L0);
                t.l[1] = addc_cc(t.l[1], P::M1);
                t.l[2] = addc_cc(t.l[2], P::M2);
                t.l[3] = addc_cc(t.l[3], P::M3);
                t.l[4] = addc_cc(t.l[4], 0u);
                t.l[5] = addc_cc(t.l[5], 0u);
                t.l[6] = addc_cc(t.l[6], 0u);
                t.l[7] = addc_cc(t.l[7], MOD_L7);
                carry = addc(0u, 0u);
            }
            shr1(t, carry);
        };
        while (!is_one(u) && !is_one(v)) {
            while (!(u.l[0] & 1u)) { shr1(u, 0u); halve_mod(b); }
            while (!(v.l[0] & 1u)) { shr1(v, 0u); halve_mod(c); }
            fe_t d;                                    // d = u - v, borrow tells which is larger
            d.l[0] = sub_cc(u.l[0], v.l[0]);
#pragma unroll
            for (int i = 1; i < 8; i++) d.l[i] = subc_cc(u.l[i], v.l[i]);
            uint32_t borrow = subc(0u, 0u);
            if (!borrow) { u = d; b = sub(b, c); }
            else {
                v.l[0] = sub_cc(v.l[0], u.l[0]);
#pragma unroll
                for (int i = 1; i < 7; i++) v.l[i] = subc_cc(v.l[i], u.l[i]);
                v.l[7] = subc(v.l[7], u.l[7]);
                c = sub(c, b);
            }
        }
        return mul(is_one(u) ? b : c, r3());
    }

    // a^(m-2) (Fermat); inv(0) = 0.  m - 2 has limbs [0xffffffff, M1-1, M2, M3, 0, 0, 0, 2^30].
    static ACC_HD fe_t inv(const fe_t &a) {
        fe_t acc = one();
        const uint32_t e[8] = {0xffffffffu, P::M1 - 1u, P::M2, P::M3, 0u, 0u, 0u, MOD_L7};
        for (int i = 254; i >= 0; i--) {
            acc = sqr(acc);
            if ((e[i >> 5] >> (i & 31)) & 1u) acc = mul(acc, a);
        }
        return acc;
    }

    // ---- square roots (Tonelli-Shanks; both fields have two-adicity 32: m - 1 = 2^32 T) --------------------------
    // Used to decompress ark-serialize points on the device (wire.cuh): y = sqrt(x^3 + 5).
    static ACC_HD fe_t sqrt_exp() {      // (T - 1) / 2
        fe_t e = zero();
        e.l[0] = FIELD == 0 ? 0xcc969876u : 0xc6237590u; e.l[1] = FIELD == 0 ? 0x04a67c8du : 0x04ca546eu;
        e.l[2] = 0x11234c7eu; e.l[6] = 0x20000000u;
        return e;
    }
    static ACC_HD fe_t root_of_unity() {  // 5^T: a primitive 2^32-th root of unity (5 is a non-residue in both fields), Montgomery
        fe_t r;
        if (FIELD == 0) {
            r.l[0] = 0xbad6dbf0u; r.l[1] = 0xa28db849u; r.l[2] = 0xd3b539dfu; r.l[3] = 0x9083cd03u;
            r.l[4] = 0x9dc8448eu; r.l[5] = 0xfba6b9cau; r.l[6] = 0x7b89c6dau; r.l[7] = 0x3ec92874u;
        } else {
            r.l[0] = 0x8c9942deu; r.l[1] = 0x21807742u; r.l[2] = 0x21b60494u; r.l[3] = 0xcc495789u;
            r.l[4] = 0xb2efbee2u; r.l[5] = 0xac2e5d27u; r.l[6] = 0x7f2db056u; r.l[7] = 0x0b79fa89u;
        }
        return r;
    }
    // out = a square root of a (either one); false when a is not a square
    static ACC_HD bool sqrt(const fe_t &a, fe_t &out) {
        if (is_zero(a)) { out = zero(); return true; }
        const fe_t e = sqrt_exp(), o = one();
        fe_t w = o;
        for (int i = 221; i >= 0; i--) {
            w = mul(w, w);
            if ((e.l[i >> 5] >> (i & 31)) & 1u) w = mul(w, a);
        }
        fe_t x = mul(a, w), b = mul(x, w), z = root_of_unity();
        uint32_t v = 32;
        while (!eq(b, o)) {
            uint32_t k = 0;
            fe_t b2 = b;
            while (!eq(b2, o) && k < v) { b2 = mul(b2, b2); k++; }
            if (k == v) return false;
            fe_t t = z;
            for (uint32_t j = 0; j + k + 1 < v; j++) t = mul(t, t);
            z = mul(t, t); b = mul(b, z); x = mul(x, t); v = k;
        }
        out = x;
        return true;
    }
    // a < b as canonical integers (inputs canonical, NOT Montgomery images)
    static ACC_HD bool lt(const fe_t &a, const fe_t &b) {
        for (int i = 7; i >= 0; i--) { if (a.l[i] != b.l[i]) return a.l[i] < b.l[i]; }
        return false;
    }
    // a canonical 256-bit integer is a valid field element
    static ACC_HD bool is_canonical(const fe_t &a) {
        fe_t m;
        for (int i = 0; i < 8; i++) m.l[i] = mod_limb(i);
        return lt(a, m);
    }
};

// Fp with the product behind a real call.  The latency-bound tail kernels (one or a few warps walking through
// dozens of point operations once) are dominated by instruction fetch when every product is inlined: an XYZZ add is
// ~2500 instructions = 40 KB of straight-line code executed exactly once.  With the product out of line an add is a
// few hundred instructions that stay in the instruction cache.  Throughput kernels keep the inlined Fp (measured:
// k_accumulate is 1.5x slower with calls).
template <int FIELD> struct FpCall : Fp<FIELD> {
#if defined(__CUDACC__)
    // operands and result travel in registers (scalars / a small struct by value), not through the local-memory stack
    struct Regs8 { uint32_t r0, r1, r2, r3, r4, r5, r6, r7; };
    static __device__ __noinline__ Regs8 mul_regs(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5,
                                                  uint32_t a6, uint32_t a7, uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3,
                                                  uint32_t b4, uint32_t b5, uint32_t b6, uint32_t b7) {
        fe_t a, b;
        a.l[0] = a0; a.l[1] = a1; a.l[2] = a2; a.l[3] = a3; a.l[4] = a4; a.l[5] = a5; a.l[6] = a6; a.l[7] = a7;
        b.l[0] = b0; b.l[1] = b1; b.l[2] = b2; b.l[3] = b3; b.l[4] = b4; b.l[5] = b5; b.l[6] = b6; b.l[7] = b7;
        fe_t r = Fp<FIELD>::mul(a, b);
        Regs8 o{r.l[0], r.l[1], r.l[2], r.l[3], r.l[4], r.l[5], r.l[6], r.l[7]};
        return o;
    }
    static __device__ __forceinline__ fe_t mul(const fe_t &a, const fe_t &b) {
        Regs8 o = mul_regs(a.l[0], a.l[1], a.l[2], a.l[3], a.l[4], a.l[5], a.l[6], a.l[7],
                           b.l[0], b.l[1], b.l[2], b.l[3], b.l[4], b.l[5], b.l[6], b.l[7]);
        fe_t r;
        r.l[0] = o.r0; r.l[1] = o.r1; r.l[2] = o.r2; r.l[3] = o.r3; r.l[4] = o.r4; r.l[5] = o.r5; r.l[6] = o.r6; r.l[7] = o.r7;
        return r;
    }
#else
    static fe_t mul(const fe_t &a, const fe_t &b) { return Fp<FIELD>::mul(a, b); }
#endif
    static ACC_HD fe_t sqr(const fe_t &a) { return mul(a, a); }
    // the latency-bound kernels keep ONE product body in the instruction cache: no dedicated squaring / dual product
    static ACC_HD fe_t mul2(const fe_t &a, const fe_t &b, const fe_t &c, const fe_t &d) { return Fp<FIELD>::add(mul(a, b), mul(c, d)); }
    static ACC_HD fe_t mul2sub(const fe_t &a, const fe_t &b, const fe_t &c, const fe_t &d) { return Fp<FIELD>::sub(mul(a, b), mul(c, d)); }
    static ACC_HD fe_t inv(const fe_t &a) {
        using P = FieldParams<FIELD>;
        fe_t acc = Fp<FIELD>::one();
        const uint32_t e[8] = {0xffffffffu, P::M1 - 1u, P::M2, P::M3, 0u, 0u, 0u, MOD_L7};
        for (int i = 254; i >= 0; i--) {
            acc = mul(acc, acc);
            if ((e[i >> 5] >> (i & 31)) & 1u) acc = mul(acc, a);
        }
        return acc;
    }
};

}  // namespace accmsm
