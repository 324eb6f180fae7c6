// Cooperative point arithmetic for the latency-bound tail of the MSM (bucket reduction leaf, small sums).
//
// One warp needs ~10 k cycles for an XYZZ addition because its 14 field products run back to back and a single warp
// cannot overlap independent products (tools/latbench.cu: the warp is bound by its own IMAD issue rate).  Four warps
// sitting on the four sub-partitions of an SM can: a GROUP of four warps holds four identical copies ("replicas") of a
// warp's 32 points; an addition runs in four phases, in each phase warp `role` computes ONE of up to four
// independent products for all 32 lanes, publishes it through shared memory, and after a named barrier every replica
// reads all four results.  Critical path: 4 products + 4 barriers instead of 14 products (doubling: 3 instead of 9).
// Everything that is not a product (field add / sub, shuffles, control flow) is simply done by all four replicas.
// The exceptional cases of the group law (identity operands, P == Q, P == -Q) are resolved per lane after the
// phases by the plain single-warp formulas, identically in every replica.
#pragma once
#include "ec.cuh"

namespace accmsm {

constexpr int COOP_WARPS = 4;
struct alignas(16) CoopScratch { fe_t v[2][COOP_WARPS][32]; };   // two generations: one barrier per phase is enough

struct CoopCtx {
    CoopScratch *sm;     // this group's scratch
    uint32_t role;       // warp index inside the group, 0..3
    uint32_t lane;
    uint32_t bar;        // named barrier id of the group (1..15)
    int gen;             // scratch generation of the next phase
};

ACC_D void coop_bar(uint32_t id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(COOP_WARPS * 32) : "memory"); }

template <int CURVE> struct Coop {
    using Cv = Curve<CURVE, FpCall>;
    using F = typename Cv::F;

    // publish this warp's product, fetch all four
    static ACC_D void exchange(CoopCtx &c, const fe_t &mine, fe_t out[COOP_WARPS]) {
        store_fe(&c.sm->v[c.gen][c.role][c.lane], mine);
        coop_bar(c.bar);
#pragma unroll
        for (int j = 0; j < COOP_WARPS; j++) out[j] = load_fe(&c.sm->v[c.gen][j][c.lane]);
        c.gen ^= 1;
    }

    // acc += q   (add-2008-s); every thread of the group must call it
    static ACC_D void add(CoopCtx &c, xyzz_t &acc, const xyzz_t &q) {
        const bool q_id = Cv::is_identity(q), acc_id = Cv::is_identity(acc);
        fe_t o[COOP_WARPS], a, b;
        switch (c.role) {
            case 0: a = acc.x; b = q.zz; break;
            case 1: a = q.x; b = acc.zz; break;
            case 2: a = acc.y; b = q.zzz; break;
            default: a = q.y; b = acc.zzz; break;
        }
        exchange(c, F::mul(a, b), o);
        const fe_t u1 = o[0], s1 = o[2];
        const fe_t p = F::sub(o[1], o[0]), r = F::sub(o[3], o[2]);
        switch (c.role) {
            case 0: a = p; b = p; break;
            case 1: a = r; b = r; break;
            case 2: a = acc.zz; b = q.zz; break;
            default: a = acc.zzz; b = q.zzz; break;
        }
        exchange(c, F::mul(a, b), o);
        const fe_t pp = o[0], rr = o[1];
        switch (c.role) {
            case 0: a = p; b = pp; break;           // PPP
            case 1: a = u1; b = pp; break;          // Q
            case 2: a = o[2]; b = pp; break;        // ZZ3 = ZZ1 ZZ2 PP
            default: a = o[3]; b = p; break;        // ZZZ1 ZZZ2 P  (x PP in the last phase)
        }
        exchange(c, F::mul(a, b), o);
        const fe_t ppp = o[0], qq = o[1], zz3 = o[2];
        const fe_t x3 = F::sub(F::sub(F::sub(rr, ppp), qq), qq);
        switch (c.role) {
            case 0: a = r; b = F::sub(qq, x3); break;
            case 1: a = s1; b = ppp; break;
            default: a = o[3]; b = pp; break;       // ZZZ3 (roles 2 and 3 compute the same value)
        }
        exchange(c, F::mul(a, b), o);
        xyzz_t res;
        res.x = x3; res.y = F::sub(o[0], o[1]); res.zz = zz3; res.zzz = o[2];
        if (q_id) res = acc;
        else if (acc_id) res = q;
        else if (F::is_zero(p)) res = F::is_zero(r) ? Cv::dbl(acc) : Cv::identity();
        acc = res;
    }

    // 2 p   (dbl-2008-s-1, a = 0); every thread of the group must call it
    static ACC_D xyzz_t dbl(CoopCtx &c, const xyzz_t &p) {
        fe_t o[COOP_WARPS], a, b;
        const fe_t u = F::dbl(p.y);
        switch (c.role) {
            case 0: a = u; b = u; break;             // V
            default: a = p.x; b = p.x; break;        // XX
        }
        exchange(c, F::mul(a, b), o);
        const fe_t v = o[0];
        const fe_t m = F::add(F::dbl(o[1]), o[1]);
        switch (c.role) {
            case 0: a = u; b = v; break;             // W
            case 1: a = p.x; b = v; break;           // S
            case 2: a = m; b = m; break;             // M^2
            default: a = v; b = p.zz; break;         // ZZ3
        }
        exchange(c, F::mul(a, b), o);
        const fe_t w = o[0], s = o[1];
        xyzz_t r;
        r.x = F::sub(F::sub(o[2], s), s);
        r.zz = o[3];
        switch (c.role) {
            case 0: a = m; b = F::sub(s, r.x); break;
            case 1: a = w; b = p.y; break;
            default: a = w; b = p.zzz; break;        // ZZZ3
        }
        exchange(c, F::mul(a, b), o);
        r.y = F::sub(o[0], o[1]);
        r.zzz = o[2];
        return Cv::is_identity(p) ? p : r;
    }
};

}  // namespace accmsm
