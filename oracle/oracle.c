/* TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, 4x64-bit Montgomery limbs, unsigned __int128) of the arithmetic that
 * arkworks-rs/accumulation's hot path reaches.  None of that arithmetic is under /root/reference:
 * it lives in the un-vendored, un-pinned dependencies named by /root/reference/Cargo.toml:15-19,
 * 33-36,39 (ark-ec ^0.2.0, ark-ff ^0.2.0, ark-poly ^0.2.0, ark-pallas ^0.2.0,
 * ark-poly-commit@accumulation-experimental).  So each function below restates the PUBLISHED
 * algorithm of that dependency and cites the reference call sites it serves.
 *
 * PARITY UNPINNED: the reference has no golden vectors / KATs for this path.  This
 * file is pinned by (1) the curve KATs of SURVEY.md App. B, (2) bit-agreement with the
 * independent Python big-int oracle oracle/pyref.py on committed fixtures (tests/golden/), and
 * (3) the algebraic identities of SURVEY.md 8c(4).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (accumulation_b200/) never links or calls it.
 */
#include "oracle.h"
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;
typedef struct {
    fe m;          /* modulus */
    uint64_t inv;  /* -m^{-1} mod 2^64 */
    fe r;          /* R mod m  (Montgomery one) */
    fe r2;         /* R^2 mod m */
} field_t;

static field_t FIELDS[2];
static int INITIALISED = 0;

/* ---------------------------------------------------------------- big-int helpers */
static inline int ge(const fe *a, const fe *b) {
    for (int i = 3; i >= 0; i--) {
        if (a->l[i] > b->l[i]) return 1;
        if (a->l[i] < b->l[i]) return 0;
    }
    return 1;
}
static inline uint64_t add4(fe *o, const fe *a, const fe *b) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a->l[i] + b->l[i]; o->l[i] = (uint64_t)c; c >>= 64; }
    return (uint64_t)c;
}
static inline uint64_t sub4(fe *o, const fe *a, const fe *b) {
    uint64_t br = 0;
    for (int i = 0; i < 4; i++) {
        u128 t = (u128)a->l[i] - b->l[i] - br;
        o->l[i] = (uint64_t)t; br = (uint64_t)(t >> 64) & 1;
    }
    return br;
}
static inline int is_zero(const fe *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int eq(const fe *a, const fe *b) { return memcmp(a, b, sizeof(fe)) == 0; }

/* ---------------------------------------------------------------- ark-ff 0.2 Fp256 semantics */
static inline void f_add(const field_t *F, fe *o, const fe *a, const fe *b) {
    fe t; add4(&t, a, b);           /* both < m < 2^255: no carry out */
    if (ge(&t, &F->m)) sub4(&t, &t, &F->m);
    *o = t;
}
static inline void f_sub(const field_t *F, fe *o, const fe *a, const fe *b) {
    fe t; if (sub4(&t, a, b)) add4(&t, &t, &F->m);
    *o = t;
}
static inline void f_dbl(const field_t *F, fe *o, const fe *a) { f_add(F, o, a, a); }
static inline void f_neg(const field_t *F, fe *o, const fe *a) {
    if (is_zero(a)) { *o = *a; return; }
    sub4(o, &F->m, a);
}
/* CIOS Montgomery product, R = 2^256 */
static inline void f_mul(const field_t *F, fe *o, const fe *a, const fe *b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a->l[j] * b->l[i] + t[j];
            t[j] = (uint64_t)c; c >>= 64;
        }
        c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
        uint64_t mm = t[0] * F->inv;
        c = (u128)mm * F->m.l[0] + t[0]; c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)mm * F->m.l[j] + t[j];
            t[j - 1] = (uint64_t)c; c >>= 64;
        }
        c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
    }
    fe r = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || ge(&r, &F->m)) sub4(&r, &r, &F->m);
    *o = r;
}
static inline void f_sqr(const field_t *F, fe *o, const fe *a) { f_mul(F, o, a, a); }
static void f_pow(const field_t *F, fe *o, const fe *a, const fe *e) {
    fe acc = F->r;
    for (int i = 255; i >= 0; i--) {
        f_sqr(F, &acc, &acc);
        if ((e->l[i / 64] >> (i % 64)) & 1) f_mul(F, &acc, &acc, a);
    }
    *o = acc;
}
static void f_inv(const field_t *F, fe *o, const fe *a) { /* Fermat; inv(0) = 0 */
    fe e = F->m; fe two = {{2, 0, 0, 0}}; sub4(&e, &e, &two);
    f_pow(F, o, a, &e);
}
static inline void f_from_mont(const field_t *F, fe *o, const fe *a) { /* into_repr() */
    fe one = {{1, 0, 0, 0}}; f_mul(F, o, a, &one);
}
static inline void f_to_mont(const field_t *F, fe *o, const fe *a) { f_mul(F, o, a, &F->r2); }

static void field_init(field_t *F, const uint64_t m[4], uint64_t inv) {
    memcpy(F->m.l, m, 32); F->inv = inv;
    /* R mod m by 256 modular doublings of 1, R^2 by 256 more */
    fe x = {{1, 0, 0, 0}};
    for (int i = 0; i < 512; i++) {
        f_add(F, &x, &x, &x);
        if (i == 255) F->r = x;
    }
    F->r2 = x;
}
static void init_once(void) {
    if (INITIALISED) return;
#pragma omp critical(oracle_init)
    {
        if (!INITIALISED) {
            /* SURVEY.md App. B */
            static const uint64_t P[4] = {0x992d30ed00000001ULL, 0x224698fc094cf91bULL, 0, 0x4000000000000000ULL};
            static const uint64_t Q[4] = {0x8c46eb2100000001ULL, 0x224698fc0994a8ddULL, 0, 0x4000000000000000ULL};
            field_init(&FIELDS[0], P, 0x992d30ecffffffffULL);
            field_init(&FIELDS[1], Q, 0x8c46eb20ffffffffULL);
            INITIALISED = 1;
        }
    }
}
static const field_t *field_of(int id) { init_once(); return &FIELDS[id & 1]; }
static const field_t *base_field(int curve) { return field_of(curve == 0 ? 0 : 1); }
static const field_t *scalar_field(int curve) { return field_of(curve == 0 ? 1 : 0); }

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
/* threads of the following calls from this thread (torchrun exports OMP_NUM_THREADS=1 to every rank) */
void oracle_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }


/* ---------------------------------------------------------------- vector field ops */
#define FE(p, i) ((const fe *)((p) + 4 * (i)))
#define FEO(p, i) ((fe *)((p) + 4 * (i)))
void oracle_fe_mul(int f, const uint64_t *a, const uint64_t *b, uint64_t *o, size_t n) {
    const field_t *F = field_of(f);
    for (size_t i = 0; i < n; i++) f_mul(F, FEO(o, i), FE(a, i), FE(b, i));
}
void oracle_fe_add(int f, const uint64_t *a, const uint64_t *b, uint64_t *o, size_t n) {
    const field_t *F = field_of(f);
    for (size_t i = 0; i < n; i++) f_add(F, FEO(o, i), FE(a, i), FE(b, i));
}
void oracle_fe_sub(int f, const uint64_t *a, const uint64_t *b, uint64_t *o, size_t n) {
    const field_t *F = field_of(f);
    for (size_t i = 0; i < n; i++) f_sub(F, FEO(o, i), FE(a, i), FE(b, i));
}
void oracle_fe_inv(int f, const uint64_t *a, uint64_t *o, size_t n) {
    const field_t *F = field_of(f);
    for (size_t i = 0; i < n; i++) f_inv(F, FEO(o, i), FE(a, i));
}
void oracle_fe_to_mont(int f, const uint64_t *a, uint64_t *o, size_t n) {
    const field_t *F = field_of(f);
    for (size_t i = 0; i < n; i++) f_to_mont(F, FEO(o, i), FE(a, i));
}
void oracle_fe_from_mont(int f, const uint64_t *a, uint64_t *o, size_t n) {
    const field_t *F = field_of(f);
    for (size_t i = 0; i < n; i++) f_from_mont(F, FEO(o, i), FE(a, i));
}

/* ---------------------------------------------------------------- SplitMix64 inputs */
static inline uint64_t splitmix(uint64_t *s) {
    uint64_t z = (*s += 0x9E3779B97F4A7C15ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
static void rand_fe(const field_t *F, uint64_t *s, fe *o) { /* canonical-range 255-bit value */
    do {
        for (int i = 0; i < 4; i++) o->l[i] = splitmix(s);
        o->l[3] &= 0x7fffffffffffffffULL;
    } while (ge(o, &F->m));
}
/* Same stream as pyref.SplitMix64(seed).field(): value v; stored as v (canonical) or v*R (mont). */
void oracle_gen_scalars(int f, uint64_t seed, size_t n, int montgomery, uint64_t *out) {
    const field_t *F = field_of(f);
    uint64_t s = seed;
    for (size_t i = 0; i < n; i++) {
        fe v; rand_fe(F, &s, &v);
        if (montgomery) f_to_mont(F, &v, &v);
        *FEO(out, i) = v;
    }
}

/* ---------------------------------------------------------------- sqrt (Tonelli-Shanks, 2-adicity 32) */
static int f_sqrt(const field_t *F, fe *o, const fe *a) {
    if (is_zero(a)) { *o = *a; return 1; }
    fe e, one = {{1, 0, 0, 0}}, t, legendre;
    /* (m-1)/2 */
    sub4(&e, &F->m, &one);
    for (int i = 0; i < 4; i++) e.l[i] = (e.l[i] >> 1) | (i < 3 ? e.l[i + 1] << 63 : 0);
    f_pow(F, &legendre, a, &e);
    if (!eq(&legendre, &F->r)) return 0;
    /* m-1 = 2^32 * tt */
    fe tt; sub4(&tt, &F->m, &one);
    for (int i = 0; i < 4; i++) tt.l[i] = (tt.l[i] >> 32) | (i < 3 ? tt.l[i + 1] << 32 : 0);
    /* non-residue z: smallest integer with legendre -1 */
    fe z, zm, c, x, b;
    for (uint64_t zi = 2;; zi++) {
        fe zc = {{zi, 0, 0, 0}}; f_to_mont(F, &zm, &zc);
        f_pow(F, &t, &zm, &e);
        if (!eq(&t, &F->r)) { z = zm; break; }
    }
    f_pow(F, &c, &z, &tt);
    fe tp1h; add4(&tp1h, &tt, &one);
    for (int i = 0; i < 4; i++) tp1h.l[i] = (tp1h.l[i] >> 1) | (i < 3 ? tp1h.l[i + 1] << 63 : 0);
    f_pow(F, &x, a, &tp1h);
    f_pow(F, &b, a, &tt);
    int s = 32;
    while (!eq(&b, &F->r)) {
        int i = 0; fe b2 = b;
        while (!eq(&b2, &F->r)) { f_sqr(F, &b2, &b2); i++; }
        fe ee = c;
        for (int j = 0; j < s - i - 1; j++) f_sqr(F, &ee, &ee);
        f_mul(F, &x, &x, &ee);
        f_sqr(F, &c, &ee);
        f_mul(F, &b, &b, &c);
        s = i;
    }
    *o = x;
    return 1;
}

/* ---------------------------------------------------------------- group law (ark-ec 0.2
 * short_weierstrass_jacobian, a = 0): Jacobian (X,Y,Z), identity <=> Z == 0 */
typedef struct { fe x, y, z; } jac;
typedef struct { fe x, y; int inf; } aff;

static inline void j_zero(const field_t *F, jac *p) { memset(p, 0, sizeof *p); p->y = F->r; }
static inline int j_is_zero(const jac *p) { return is_zero(&p->z); }

static void j_double(const field_t *F, jac *p) { /* dbl-2009-l */
    if (j_is_zero(p)) return;
    fe a, b, c, d, e, f, t;
    f_sqr(F, &a, &p->x); f_sqr(F, &b, &p->y); f_sqr(F, &c, &b);
    f_add(F, &t, &p->x, &b); f_sqr(F, &t, &t); f_sub(F, &t, &t, &a); f_sub(F, &t, &t, &c);
    f_dbl(F, &d, &t);
    f_dbl(F, &e, &a); f_add(F, &e, &e, &a);
    f_sqr(F, &f, &e);
    f_mul(F, &p->z, &p->z, &p->y); f_dbl(F, &p->z, &p->z);
    f_sub(F, &p->x, &f, &d); f_sub(F, &p->x, &p->x, &d);
    f_sub(F, &t, &d, &p->x); f_mul(F, &t, &t, &e);
    f_dbl(F, &c, &c); f_dbl(F, &c, &c); f_dbl(F, &c, &c);
    f_sub(F, &p->y, &t, &c);
}
static void j_add_mixed(const field_t *F, jac *p, const aff *q) { /* madd-2007-bl */
    if (q->inf) return;
    if (j_is_zero(p)) { p->x = q->x; p->y = q->y; p->z = F->r; return; }
    fe z1z1, u2, s2, h, hh, i, j, r, v, t;
    f_sqr(F, &z1z1, &p->z);
    f_mul(F, &u2, &q->x, &z1z1);
    f_mul(F, &s2, &q->y, &p->z); f_mul(F, &s2, &s2, &z1z1);
    if (eq(&p->x, &u2) && eq(&p->y, &s2)) { j_double(F, p); return; }
    f_sub(F, &h, &u2, &p->x);
    f_sqr(F, &hh, &h);
    f_dbl(F, &i, &hh); f_dbl(F, &i, &i);
    f_mul(F, &j, &h, &i);
    f_sub(F, &r, &s2, &p->y); f_dbl(F, &r, &r);
    f_mul(F, &v, &p->x, &i);
    fe x3, y3, z3;
    f_sqr(F, &x3, &r); f_sub(F, &x3, &x3, &j); f_sub(F, &x3, &x3, &v); f_sub(F, &x3, &x3, &v);
    f_mul(F, &t, &p->y, &j); f_dbl(F, &t, &t);
    f_sub(F, &y3, &v, &x3); f_mul(F, &y3, &y3, &r); f_sub(F, &y3, &y3, &t);
    f_add(F, &z3, &p->z, &h); f_sqr(F, &z3, &z3); f_sub(F, &z3, &z3, &z1z1); f_sub(F, &z3, &z3, &hh);
    p->x = x3; p->y = y3; p->z = z3;
}
static void j_add(const field_t *F, jac *p, const jac *q) { /* add-2007-bl */
    if (j_is_zero(q)) return;
    if (j_is_zero(p)) { *p = *q; return; }
    fe z1z1, z2z2, u1, u2, s1, s2, h, i, j, r, v, t;
    f_sqr(F, &z1z1, &p->z); f_sqr(F, &z2z2, &q->z);
    f_mul(F, &u1, &p->x, &z2z2); f_mul(F, &u2, &q->x, &z1z1);
    f_mul(F, &s1, &p->y, &q->z); f_mul(F, &s1, &s1, &z2z2);
    f_mul(F, &s2, &q->y, &p->z); f_mul(F, &s2, &s2, &z1z1);
    if (eq(&u1, &u2) && eq(&s1, &s2)) { j_double(F, p); return; }
    f_sub(F, &h, &u2, &u1);
    f_dbl(F, &i, &h); f_sqr(F, &i, &i);
    f_mul(F, &j, &h, &i);
    f_sub(F, &r, &s2, &s1); f_dbl(F, &r, &r);
    f_mul(F, &v, &u1, &i);
    fe x3, y3, z3;
    f_sqr(F, &x3, &r); f_sub(F, &x3, &x3, &j); f_sub(F, &x3, &x3, &v); f_sub(F, &x3, &x3, &v);
    f_mul(F, &t, &s1, &j); f_dbl(F, &t, &t);
    f_sub(F, &y3, &v, &x3); f_mul(F, &y3, &y3, &r); f_sub(F, &y3, &y3, &t);
    f_add(F, &z3, &p->z, &q->z); f_sqr(F, &z3, &z3); f_sub(F, &z3, &z3, &z1z1);
    f_sub(F, &z3, &z3, &z2z2); f_mul(F, &z3, &z3, &h);
    p->x = x3; p->y = y3; p->z = z3;
}
/* into_affine(): identity encoded as ark-ec does: (0, 1, infinity = true) */
static void j_to_affine(const field_t *F, const jac *p, uint64_t *out_xy, uint8_t *out_inf) {
    if (j_is_zero(p)) {
        memset(out_xy, 0, 64); memcpy(out_xy + 4, F->r.l, 32); *out_inf = 1; return;
    }
    fe zi, zi2, zi3, x, y;
    f_inv(F, &zi, &p->z); f_sqr(F, &zi2, &zi); f_mul(F, &zi3, &zi2, &zi);
    f_mul(F, &x, &p->x, &zi2); f_mul(F, &y, &p->y, &zi3);
    memcpy(out_xy, x.l, 32); memcpy(out_xy + 4, y.l, 32); *out_inf = 0;
}
static inline void load_aff(aff *a, const uint64_t *xy, int inf) {
    memcpy(a->x.l, xy, 32); memcpy(a->y.l, xy + 4, 32); a->inf = inf;
}
/* variable-base scalar mul: double-and-add, MSB first (GroupAffine::mul semantics) */
static void j_mul(const field_t *F, jac *out, const aff *b, const fe *k_canon) {
    j_zero(F, out);
    int started = 0;
    for (int i = 255; i >= 0; i--) {
        if (started) j_double(F, out);
        if ((k_canon->l[i / 64] >> (i % 64)) & 1) { j_add_mixed(F, out, b); started = 1; }
    }
}

int oracle_on_curve(int curve, const uint64_t *xy) {
    const field_t *F = base_field(curve);
    fe x, y, l, r, five = {{5, 0, 0, 0}};
    memcpy(x.l, xy, 32); memcpy(y.l, xy + 4, 32);
    f_to_mont(F, &five, &five);
    f_sqr(F, &l, &y);
    f_sqr(F, &r, &x); f_mul(F, &r, &r, &x); f_add(F, &r, &r, &five);
    return eq(&l, &r);
}
void oracle_point_mul(int curve, const uint64_t *xy, uint8_t inf, const uint64_t *k,
                      uint64_t *out_xy, uint8_t *out_inf) {
    const field_t *F = base_field(curve);
    aff a; load_aff(&a, xy, inf); jac r; fe kk; memcpy(kk.l, k, 32);
    j_mul(F, &r, &a, &kk); j_to_affine(F, &r, out_xy, out_inf);
}
void oracle_point_add(int curve, const uint64_t *a_xy, uint8_t a_inf, const uint64_t *b_xy,
                      uint8_t b_inf, uint64_t *out_xy, uint8_t *out_inf) {
    const field_t *F = base_field(curve);
    aff a, b; load_aff(&a, a_xy, a_inf); load_aff(&b, b_xy, b_inf);
    jac r; j_zero(F, &r); j_add_mixed(F, &r, &a); j_add_mixed(F, &r, &b);
    j_to_affine(F, &r, out_xy, out_inf);
}

/* Synthetic bases: 64 anchor points from x-sampling + sqrt, then each of T chunks
 * walks P_{i+1} = P_i + anchor[r_i] so 2^24 points cost ~one mixed add each; batch-normalised. */
static void sample_point(const field_t *F, uint64_t *s, aff *o) {
    fe five = {{5, 0, 0, 0}}; f_to_mont(F, &five, &five);
    for (;;) {
        fe xc, x, rhs, y;
        rand_fe(F, s, &xc); f_to_mont(F, &x, &xc);
        f_sqr(F, &rhs, &x); f_mul(F, &rhs, &rhs, &x); f_add(F, &rhs, &rhs, &five);
        if (!f_sqrt(F, &y, &rhs)) continue;
        if (splitmix(s) & 1) f_neg(F, &y, &y);
        o->x = x; o->y = y; o->inf = 0; return;
    }
}
void oracle_gen_points(int curve, uint64_t seed, size_t n, uint64_t *out_xy) {
    const field_t *F = base_field(curve);
    enum { NA = 64, CHUNK = 4096 };
    aff anchors[NA];
    uint64_t s = seed ^ 0xA5A5A5A5DEADBEEFULL;
    for (int i = 0; i < NA; i++) sample_point(F, &s, &anchors[i]);
    size_t nchunks = (n + CHUNK - 1) / CHUNK;
#pragma omp parallel for schedule(dynamic, 1)
    for (size_t ch = 0; ch < nchunks; ch++) {
        size_t lo = ch * CHUNK, hi = lo + CHUNK < n ? lo + CHUNK : n, m = hi - lo;
        uint64_t cs = seed + 0x1234567ULL * (ch + 1);
        jac *pts = (jac *)malloc(m * sizeof(jac));
        fe *pref = (fe *)malloc(m * sizeof(fe));
        aff start; sample_point(F, &cs, &start);
        jac cur; j_zero(F, &cur); j_add_mixed(F, &cur, &start);
        for (size_t i = 0; i < m; i++) {
            pts[i] = cur;
            j_add_mixed(F, &cur, &anchors[splitmix(&cs) % NA]);
            if (j_is_zero(&cur)) j_add_mixed(F, &cur, &start); /* never in practice */
        }
        /* Montgomery batch inversion of z */
        fe acc = F->r;
        for (size_t i = 0; i < m; i++) { pref[i] = acc; f_mul(F, &acc, &acc, &pts[i].z); }
        fe inv; f_inv(F, &inv, &acc);
        for (size_t i = m; i-- > 0;) {
            fe zi, zi2, zi3, x, y;
            f_mul(F, &zi, &inv, &pref[i]); f_mul(F, &inv, &inv, &pts[i].z);
            f_sqr(F, &zi2, &zi); f_mul(F, &zi3, &zi2, &zi);
            f_mul(F, &x, &pts[i].x, &zi2); f_mul(F, &y, &pts[i].y, &zi3);
            memcpy(out_xy + 8 * (lo + i), x.l, 32); memcpy(out_xy + 8 * (lo + i) + 4, y.l, 32);
        }
        free(pts); free(pref);
    }
}

/* ---------------------------------------------------------------- ark-ec 0.2.0 VariableBaseMSM
 * (SURVEY App. A.1).  Reached from the reference via PedersenCommitment::commit
 * (src/hp_as/mod.rs:196,197,214,377,910-918; src/r1cs_nark_as/r1cs_nark/mod.rs:216-218,...) and
 * IpaPC::cm_commit (src/ipa_pc_as/mod.rs:155,454-462,836-845). */
static unsigned ceil_log2(size_t a) { unsigned l = 0; while (((size_t)1 << l) < a) l++; return l; }
static void msm_ark(const field_t *F, const uint64_t *bases_xy, const uint8_t *bases_inf,
                    const uint64_t *scalars, size_t size, jac *out) {
    unsigned c = size < 32 ? 3 : (ceil_log2(size) * 69 / 100) + 2;
    const unsigned num_bits = 255;
    unsigned nwin = (num_bits + c - 1) / c;
    jac *window_sums = (jac *)malloc(nwin * sizeof(jac));
    const fe one = {{1, 0, 0, 0}};
#pragma omp parallel for schedule(dynamic, 1)
    for (unsigned w = 0; w < nwin; w++) {
        unsigned w_start = w * c;
        size_t nb = ((size_t)1 << c) - 1;
        jac res; j_zero(F, &res);
        jac *buckets = (jac *)malloc(nb * sizeof(jac));
        for (size_t b = 0; b < nb; b++) j_zero(F, &buckets[b]);
        for (size_t i = 0; i < size; i++) {
            const fe *s = FE(scalars, i);
            if (is_zero(s)) continue;
            aff base; load_aff(&base, bases_xy + 8 * i, bases_inf ? bases_inf[i] : 0);
            if (eq(s, &one)) {
                if (w_start == 0) j_add_mixed(F, &res, &base);
            } else {
                /* (scalar >> w_start) % 2^c on the low limb after the shift */
                unsigned limb = w_start / 64, off = w_start % 64;
                uint64_t v = s->l[limb] >> off;
                if (off && limb + 1 < 4) v |= s->l[limb + 1] << (64 - off);
                v &= ((uint64_t)1 << c) - 1;
                if (v) j_add_mixed(F, &buckets[v - 1], &base);
            }
        }
        jac running; j_zero(F, &running);
        for (size_t b = nb; b-- > 0;) { j_add(F, &running, &buckets[b]); j_add(F, &res, &running); }
        free(buckets);
        window_sums[w] = res;
    }
    jac total; j_zero(F, &total);
    for (unsigned w = nwin - 1; w >= 1; w--) {
        j_add(F, &total, &window_sums[w]);
        for (unsigned k = 0; k < c; k++) j_double(F, &total);
    }
    j_add(F, &total, &window_sums[0]);
    free(window_sums);
    *out = total;
}
void oracle_msm_ark(int curve, const uint64_t *bases_xy, const uint8_t *bases_inf, size_t n_bases,
                    const uint64_t *scalars, size_t n_scalars, uint64_t *out_xy, uint8_t *out_inf) {
    const field_t *F = base_field(curve);
    size_t size = n_bases < n_scalars ? n_bases : n_scalars;
    jac r; msm_ark(F, bases_xy, bases_inf, scalars, size, &r);
    j_to_affine(F, &r, out_xy, out_inf);
}
void oracle_commit(int curve, const uint64_t *bases_xy, size_t n_bases, const uint64_t *elems,
                   size_t n_elems, const uint64_t *hiding_xy, const uint64_t *randomizer,
                   uint64_t *out_xy, uint8_t *out_inf) {
    const field_t *F = base_field(curve), *S = scalar_field(curve);
    size_t size = n_bases < n_elems ? n_bases : n_elems;
    uint64_t *repr = (uint64_t *)malloc(size ? size * 32 : 32);
#pragma omp parallel for
    for (size_t i = 0; i < size; i++) f_from_mont(S, FEO(repr, i), FE(elems, i));
    jac r; msm_ark(F, bases_xy, NULL, repr, size, &r);
    free(repr);
    if (randomizer && hiding_xy) {
        fe k; f_from_mont(S, &k, (const fe *)randomizer);
        aff h; load_aff(&h, hiding_xy, 0);
        jac t; j_mul(F, &t, &h, &k); j_add(F, &r, &t);
    }
    j_to_affine(F, &r, out_xy, out_inf);
}

/* ---------------------------------------------------------------- SuccinctCheckPolynomial
 * (called at src/ipa_pc_as/mod.rs:400,418 and inside IpaPC::check :836) */
void oracle_compute_coeffs(int f, const uint64_t *ch, int k, uint64_t *coeffs) {
    const field_t *F = field_of(f);
    size_t n = (size_t)1 << k;
    for (size_t j = 0; j < n; j++) *FEO(coeffs, j) = F->r;
    for (int i = 1; i <= k; i++) {
        size_t e = (size_t)1 << (k - i);
        const fe *xi = FE(ch, i - 1);
        for (size_t start = e; start < n; start += 2 * e)
            for (size_t j = start; j < start + e; j++) f_mul(F, FEO(coeffs, j), FE(coeffs, j), xi);
    }
}
void oracle_succinct_evaluate(int f, const uint64_t *ch, int k, const uint64_t *z, uint64_t *out) {
    const field_t *F = field_of(f);
    /* prod_i (1 + xi_i * z^(2^(k-i))) : walk i = k..1 squaring z */
    fe zp; memcpy(zp.l, z, 32);
    fe acc = F->r, t;
    for (int i = k; i >= 1; i--) {
        f_mul(F, &t, FE(ch, i - 1), &zp); f_add(F, &t, &t, &F->r);
        f_mul(F, &acc, &acc, &t);
        f_sqr(F, &zp, &zp);
    }
    memcpy(out, acc.l, 32);
}
void oracle_poly_evaluate(int f, const uint64_t *coeffs, size_t n, const uint64_t *z, uint64_t *out) {
    const field_t *F = field_of(f);
    fe acc; memset(&acc, 0, sizeof acc);
    for (size_t i = n; i-- > 0;) { f_mul(F, &acc, &acc, (const fe *)z); f_add(F, &acc, &acc, FE(coeffs, i)); }
    memcpy(out, acc.l, 32);
}
int oracle_ipa_check_final_key(int curve, const uint64_t *key_xy, size_t n_key, const uint64_t *ch,
                               int k, const uint64_t *exp_xy, uint8_t exp_inf, uint64_t *out_xy,
                               uint8_t *out_inf) {
    size_t n = (size_t)1 << k;
    uint64_t *coeffs = (uint64_t *)malloc(n * 32);
    oracle_compute_coeffs(curve == 0 ? 1 : 0, ch, k, coeffs);
    oracle_commit(curve, key_xy, n_key, coeffs, n, NULL, NULL, out_xy, out_inf);
    free(coeffs);
    if (*out_inf || exp_inf) return *out_inf == exp_inf;
    return memcmp(out_xy, exp_xy, 64) == 0;
}
void oracle_ipa_fold_key(int curve, const uint64_t *key_xy, size_t n_key, const uint64_t *ch, int k,
                         uint64_t *out_xy, uint8_t *out_inf) {
    const field_t *F = base_field(curve), *S = scalar_field(curve);
    size_t n = n_key;
    uint64_t *cur = (uint64_t *)malloc(n * 64);
    uint8_t *inf = (uint8_t *)calloc(n, 1);
    memcpy(cur, key_xy, n * 64);
    for (int r = 0; r < k; r++) {
        size_t h = n / 2;
        fe xi; f_from_mont(S, &xi, FE(ch, r));
#pragma omp parallel for
        for (size_t i = 0; i < h; i++) {
            aff l, rr; load_aff(&l, cur + 8 * i, inf[i]); load_aff(&rr, cur + 8 * (i + h), inf[i + h]);
            jac t; j_mul(F, &t, &rr, &xi); j_add_mixed(F, &t, &l);
            j_to_affine(F, &t, cur + 8 * i, &inf[i]);
        }
        n = h;
    }
    memcpy(out_xy, cur, 64); *out_inf = inf[0];
    free(cur); free(inf);
}
/* ---------------------------------------------------------------- IpaPC::open / succinct_check
 * (ark-poly-commit ipa_pc; reference call sites src/ipa_pc_as/mod.rs:454-462 (prove),
 * :525-534 (index default proof), :198-205 (succinct_check), examples/scaling-pc.rs:72-81).  The round
 * challenges come from the host sponge, so the restatement is per round: the caller squeezes xi from
 * (l, r) between oracle_ipa_open_round_lr and oracle_ipa_open_fold. */
static void inner_product(const field_t *S, const uint64_t *a, const uint64_t *b, size_t n, fe *out) {
    fe acc; memset(&acc, 0, sizeof acc);
    for (size_t i = 0; i < n; i++) { fe t; f_mul(S, &t, FE(a, i), FE(b, i)); f_add(S, &acc, &acc, &t); }
    *out = acc;
}
/* z_vec = (1, z, z^2, ...) */
void oracle_powers(int f, const uint64_t *z, size_t n, uint64_t *out) {
    const field_t *F = field_of(f);
    fe cur = F->r;
    for (size_t i = 0; i < n; i++) { *FEO(out, i) = cur; f_mul(F, &cur, &cur, (const fe *)z); }
}
/* state: key (n points, affine), coeffs (n), z (n); n even.
 *   l = cm_commit(key_l, coeffs_r) + <coeffs_r, z_l> h'      r = cm_commit(key_r, coeffs_l) + <coeffs_l, z_r> h' */
void oracle_ipa_open_round_lr(int curve, const uint64_t *key_xy, const uint64_t *coeffs,
                              const uint64_t *z, size_t n, const uint64_t *h_prime_xy,
                              uint64_t *l_xy, uint8_t *l_inf, uint64_t *r_xy, uint8_t *r_inf) {
    const field_t *S = scalar_field(curve);
    size_t h = n / 2;
    fe ipl, ipr;
    inner_product(S, coeffs + 4 * h, z, h, &ipl);
    inner_product(S, coeffs, z + 4 * h, h, &ipr);
    oracle_commit(curve, key_xy, h, coeffs + 4 * h, h, h_prime_xy, ipl.l, l_xy, l_inf);
    oracle_commit(curve, key_xy + 8 * h, h, coeffs, h, h_prime_xy, ipr.l, r_xy, r_inf);
}
/* coeffs_l += xi^-1 coeffs_r ; z_l += xi z_r ; key_l += xi key_r (normalised); the first n/2 entries hold
 * the folded state afterwards */
void oracle_ipa_open_fold(int curve, uint64_t *key_xy, uint64_t *coeffs, uint64_t *z, size_t n,
                          const uint64_t *xi_mont, const uint64_t *xi_inv_mont) {
    const field_t *F = base_field(curve), *S = scalar_field(curve);
    size_t h = n / 2;
    fe xi; f_from_mont(S, &xi, (const fe *)xi_mont);
#pragma omp parallel for
    for (size_t i = 0; i < h; i++) {
        fe t;
        f_mul(S, &t, (const fe *)xi_inv_mont, FE(coeffs, i + h)); f_add(S, FEO(coeffs, i), FE(coeffs, i), &t);
        f_mul(S, &t, (const fe *)xi_mont, FE(z, i + h)); f_add(S, FEO(z, i), FE(z, i), &t);
        aff l, rr; load_aff(&l, key_xy + 8 * i, 0); load_aff(&rr, key_xy + 8 * (i + h), 0);
        jac p; j_mul(F, &p, &rr, &xi); j_add_mixed(F, &p, &l);
        uint8_t inf; j_to_affine(F, &p, key_xy + 8 * i, &inf);
    }
}
/* succinct_check's group equation with the transcript values given:
 *   C' = C + v h' + sum_i (xi_i^-1 l_i + xi_i r_i) ;  accept iff C' == c final_key + (h(z) c) h' */
int oracle_ipa_succinct_check(int curve, const uint64_t *comm_xy, uint8_t comm_inf, const uint64_t *z,
                              const uint64_t *v, const uint64_t *l_xy, const uint64_t *r_xy, int k,
                              const uint64_t *xi_mont, const uint64_t *h_prime_xy,
                              const uint64_t *final_key_xy, const uint64_t *c_mont) {
    const field_t *F = base_field(curve), *S = scalar_field(curve);
    int sf = curve == 0 ? 1 : 0;
    jac acc; j_zero(F, &acc);
    aff a; load_aff(&a, comm_xy, comm_inf);
    if (!comm_inf) j_add_mixed(F, &acc, &a);
    aff hp; load_aff(&hp, h_prime_xy, 0);
    fe s; jac t;
    f_from_mont(S, &s, (const fe *)v); j_mul(F, &t, &hp, &s); j_add(F, &acc, &t);
    for (int i = 0; i < k; i++) {
        fe xinv; f_inv(S, &xinv, FE(xi_mont, i));
        aff l, r; load_aff(&l, l_xy + 8 * i, 0); load_aff(&r, r_xy + 8 * i, 0);
        f_from_mont(S, &s, &xinv); j_mul(F, &t, &l, &s); j_add(F, &acc, &t);
        f_from_mont(S, &s, FE(xi_mont, i)); j_mul(F, &t, &r, &s); j_add(F, &acc, &t);
    }
    uint64_t hz[4]; oracle_succinct_evaluate(sf, xi_mont, k, z, hz);
    fe vp; f_mul(S, &vp, (const fe *)hz, (const fe *)c_mont);
    jac rhs; j_zero(F, &rhs);
    aff fk; load_aff(&fk, final_key_xy, 0);
    f_from_mont(S, &s, (const fe *)c_mont); j_mul(F, &t, &fk, &s); j_add(F, &rhs, &t);
    f_from_mont(S, &s, &vp); j_mul(F, &t, &hp, &s); j_add(F, &rhs, &t);
    uint64_t a_xy[8], b_xy[8]; uint8_t a_inf, b_inf;
    j_to_affine(F, &acc, a_xy, &a_inf); j_to_affine(F, &rhs, b_xy, &b_inf);
    if (a_inf || b_inf) return a_inf == b_inf;
    return memcmp(a_xy, b_xy, 64) == 0;
}

void oracle_combine_check_polys(int f, const uint64_t *ch, int m, int k, const uint64_t *alphas,
                                const uint64_t *random_poly, size_t n_random, uint64_t *out) {
    const field_t *F = field_of(f);
    size_t n = (size_t)1 << k;
    memset(out, 0, n * 32);
    if (random_poly) memcpy(out, random_poly, n_random * 32);
    uint64_t *tmp = (uint64_t *)malloc(n * 32);
    for (int j = 0; j < m; j++) {
        oracle_compute_coeffs(f, ch + 4 * (size_t)j * k, k, tmp);
        for (size_t i = 0; i < n; i++) {
            fe t; f_mul(F, &t, FE(tmp, i), FE(alphas, j)); f_add(F, FEO(out, i), FE(out, i), &t);
        }
    }
    free(tmp);
}

/* ---------------------------------------------------------------- hp_as vector ops */
void oracle_hadamard(int f, const uint64_t *a, const uint64_t *b, uint64_t *o, size_t n) {
    oracle_fe_mul(f, a, b, o, n); /* src/hp_as/mod.rs:278-285 */
}
void oracle_scale(int f, const uint64_t *v, const uint64_t *c, uint64_t *o, size_t n) {
    const field_t *F = field_of(f); /* src/hp_as/mod.rs:482-489 */
    for (size_t i = 0; i < n; i++) f_mul(F, FEO(o, i), FE(v, i), (const fe *)c);
}
void oracle_combine_vectors(int f, const uint64_t *const *vecs, const size_t *lens, int m,
                            const uint64_t *ch, const uint64_t *hiding, size_t n_hiding,
                            uint64_t *out, size_t out_len) {
    const field_t *F = field_of(f); /* src/hp_as/mod.rs:492-512 */
    memset(out, 0, out_len * 32);
    if (hiding) memcpy(out, hiding, n_hiding * 32);
    for (int ni = 0; ni < m; ni++)
        for (size_t li = 0; li < lens[ni] && li < out_len; li++) {
            fe t; f_mul(F, &t, FE(ch, ni), FE(vecs[ni], li)); f_add(F, FEO(out, li), FE(out, li), &t);
        }
}
void oracle_tvecs(int f, const uint64_t *const *a_vecs, const size_t *a_lens,
                  const uint64_t *const *b_vecs, const size_t *b_lens, int n, const uint64_t *mu,
                  size_t len, const uint64_t *ha, size_t n_ha, const uint64_t *hb, size_t n_hb,
                  uint64_t *out) {
    const field_t *F = field_of(f); /* src/hp_as/mod.rs:288-349 */
    fe *ac = (fe *)malloc(n * sizeof(fe)), *bc = (fe *)malloc(n * sizeof(fe));
    fe *t = (fe *)malloc((2 * n - 1) * sizeof(fe));
    for (size_t li = 0; li < len; li++) {
        for (int i = 0; i < n; i++) {
            if (li < a_lens[i]) f_mul(F, &ac[i], FE(mu, i), FE(a_vecs[i], li)); else memset(&ac[i], 0, 32);
            int rj = n - 1 - i; /* b reversed (:320) */
            if (li < b_lens[i]) bc[rj] = *FE(b_vecs[i], li); else memset(&bc[rj], 0, 32);
        }
        if (ha && li < n_ha) { fe x; f_mul(F, &x, FE(ha, li), FE(mu, n)); f_add(F, &ac[0], &ac[0], &x); }
        if (hb && li < n_hb) { fe x; f_mul(F, &x, FE(hb, li), FE(mu, 1)); f_add(F, &bc[0], &bc[0], &x); }
        memset(t, 0, (2 * n - 1) * sizeof(fe));
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) { fe x; f_mul(F, &x, &ac[i], &bc[j]); f_add(F, &t[i + j], &t[i + j], &x); }
        for (int k = 0; k < 2 * n - 1; k++) *FEO(out, (size_t)k * len + li) = t[k];
    }
    free(ac); free(bc); free(t);
}

/* ---------------------------------------------------------------- r1cs_nark matrix_vec_mul */
void oracle_csr_matvec(int f, const uint32_t *row_ptr, const uint32_t *cols, const uint64_t *coeffs,
                       size_t n_rows, const uint64_t *input, size_t n_input, const uint64_t *witness,
                       size_t n_witness, uint64_t *out) {
    const field_t *F = field_of(f); /* src/r1cs_nark_as/r1cs_nark/mod.rs:443-462, rayon over rows */
    (void)n_witness;
#pragma omp parallel for
    for (size_t r = 0; r < n_rows; r++) {
        fe acc; memset(&acc, 0, 32);
        for (uint32_t e = row_ptr[r]; e < row_ptr[r + 1]; e++) {
            size_t col = cols[e];
            const fe *z = col < n_input ? FE(input, col) : FE(witness, col - n_input);
            fe t;
            if (eq(FE(coeffs, e), &F->r)) t = *z; else f_mul(F, &t, z, FE(coeffs, e));
            f_add(F, &acc, &acc, &t);
        }
        *FEO(out, r) = acc;
    }
}
