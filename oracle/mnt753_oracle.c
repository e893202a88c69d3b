/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference's CPU MSM path.
 * See mnt753_oracle.h for scope, wire formats and the parity-pinning statement (PINNED against
 * the reference's libff through tests/golden/).  Each function cites the reference lines it follows
 * (paths relative to /root/reference/depends/libff/libff unless stated).
 */
#include "mnt753_oracle.h"

#include <omp.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "oracle_constants.h"

#define NL 12
typedef unsigned __int128 u128;

typedef struct { uint64_t l[NL]; } fp;
typedef struct { uint64_t p[NL], r1[NL], r2[NL], inv; } fpar;

static const fpar PAR_A = {MNT753_MOD_A_U64, MNT753_R1_A_U64, MNT753_R2_A_U64, MNT753_INV_A_U64};
static const fpar PAR_B = {MNT753_MOD_B_U64, MNT753_R1_B_U64, MNT753_R2_B_U64, MNT753_INV_B_U64};

/* ---------------------------------------------------------------- bigint helpers (mpn_* stand-ins) */
static int bn_cmp(const uint64_t *a, const uint64_t *b) {
    for (int i = NL - 1; i >= 0; --i)
        if (a[i] != b[i]) return a[i] > b[i] ? 1 : -1;
    return 0;
}
static uint64_t bn_add(uint64_t *r, const uint64_t *a, const uint64_t *b) {
    u128 c = 0;
    for (int i = 0; i < NL; ++i) { c += (u128)a[i] + b[i]; r[i] = (uint64_t)c; c >>= 64; }
    return (uint64_t)c;
}
static uint64_t bn_sub(uint64_t *r, const uint64_t *a, const uint64_t *b) {
    uint64_t bw = 0;
    for (int i = 0; i < NL; ++i) {
        u128 d = (u128)a[i] - b[i] - bw;
        r[i] = (uint64_t)d;
        bw = (uint64_t)(d >> 64) & 1;
    }
    return bw;
}
static int bn_is_zero(const uint64_t *a) {
    uint64_t o = 0;
    for (int i = 0; i < NL; ++i) o |= a[i];
    return o == 0;
}
static int bn_test_bit(const uint64_t *a, size_t bit) { return bit < 64 * NL && ((a[bit / 64] >> (bit % 64)) & 1); }
/* bigint<n>::num_bits, algebra/fields/bigint.tcc:103-130 */
static size_t bn_num_bits(const uint64_t *a) {
    for (int i = NL - 1; i >= 0; --i)
        if (a[i]) return (size_t)(i + 1) * 64 - (size_t)__builtin_clzll(a[i]);
    return 0;
}

/* ---------------------------------------------------------------- Fp (algebra/fields/fp.tcc) */
/* Fp_model::mul_reduce generic path, fp.tcc:161-186: full product, then HAC 14.32 reduction. */
static void fp_mul_raw(const fpar *P, uint64_t *r, const uint64_t *a, const uint64_t *b) {
    uint64_t res[2 * NL + 1];
    memset(res, 0, sizeof res);
    for (int i = 0; i < NL; ++i) { /* mpn_mul_n */
        u128 c = 0;
        for (int j = 0; j < NL; ++j) {
            c += (u128)a[j] * b[i] + res[i + j];
            res[i + j] = (uint64_t)c;
            c >>= 64;
        }
        res[i + NL] = (uint64_t)c;
    }
    for (int i = 0; i < NL; ++i) {
        uint64_t k = P->inv * res[i];
        u128 c = 0;
        for (int j = 0; j < NL; ++j) { /* mpn_addmul_1(res+i, mod, n, k) */
            c += (u128)P->p[j] * k + res[i + j];
            res[i + j] = (uint64_t)c;
            c >>= 64;
        }
        for (int j = i + NL; j < 2 * NL + 1 && c; ++j) { /* mpn_add_1 */
            c += res[j];
            res[j] = (uint64_t)c;
            c >>= 64;
        }
    }
    if (res[2 * NL] || bn_cmp(res + NL, P->p) >= 0) bn_sub(res + NL, res + NL, P->p);
    memcpy(r, res + NL, NL * 8);
}
static void fp_mul(const fpar *P, fp *r, const fp *a, const fp *b) { fp_mul_raw(P, r->l, a->l, b->l); }
static void fp_sqr(const fpar *P, fp *r, const fp *a) { fp_mul_raw(P, r->l, a->l, a->l); }
/* Fp_model::operator+=, fp.tcc:310-433 (generic tail: add, then subtract modulus on carry or >=) */
static void fp_add(const fpar *P, fp *r, const fp *a, const fp *b) {
    uint64_t c = bn_add(r->l, a->l, b->l);
    if (c || bn_cmp(r->l, P->p) >= 0) bn_sub(r->l, r->l, P->p);
}
/* Fp_model::operator-=, fp.tcc:436-530 (generic tail: add modulus first when a < b) */
static void fp_sub(const fpar *P, fp *r, const fp *a, const fp *b) {
    if (bn_cmp(a->l, b->l) < 0) {
        uint64_t t[NL];
        bn_add(t, a->l, P->p); /* carry, if any, is cancelled by the borrow below */
        bn_sub(r->l, t, b->l);
    } else {
        bn_sub(r->l, a->l, b->l);
    }
}
/* Fp_model::operator-(), fp.tcc:606-622: 0 -> 0, else p - x */
static void fp_neg(const fpar *P, fp *r, const fp *a) {
    if (bn_is_zero(a->l)) { *r = *a; return; }
    bn_sub(r->l, P->p, a->l);
}
static int fp_is_zero(const fp *a) { return bn_is_zero(a->l); }
static int fp_eq(const fp *a, const fp *b) { return bn_cmp(a->l, b->l) == 0; }
static void fp_one(const fpar *P, fp *r) { memcpy(r->l, P->r1, NL * 8); }
static void fp_zero(fp *r) { memset(r, 0, sizeof *r); }
/* Fp_model::as_bigint, fp.tcc:246-261: Montgomery-reduce by multiplying with the integer 1 */
static void fp_from_mont(const fpar *P, uint64_t *r, const fp *a) {
    uint64_t one[NL] = {1};
    fp_mul_raw(P, r, a->l, one);
}
/* Fp_model(bigint), fp.tcc:189-194: multiply by R^2 */
static void fp_to_mont(const fpar *P, fp *r, const uint64_t *a) { fp_mul_raw(P, r->l, a, P->r2); }
/* Fp_model::inverse, fp.tcc:700-745, computes the unique x^-1 with mpn_gcdext; the same field
 * element is obtained here as x^(p-2) (Fermat), result in Montgomery form. */
static void fp_inv(const fpar *P, fp *r, const fp *a) {
    uint64_t e[NL], two[NL] = {2};
    bn_sub(e, P->p, two);
    fp acc, base = *a;
    fp_one(P, &acc);
    for (size_t i = 0; i < 64 * NL; ++i) {
        if (bn_test_bit(e, i)) fp_mul(P, &acc, &acc, &base);
        fp_sqr(P, &base, &base);
    }
    *r = acc;
}

/* ---------------------------------------------------------------- Fp2 / Fp3 towers */
typedef struct { fp c[3]; } fe;
typedef struct {
    const fpar *fq, *fr;
    int deg;      /* 1, 2, 3 */
    fp nr;        /* non-residue, Montgomery (13 for MNT4753 Fq2, 11 for MNT6753 Fq3) */
    fp mba[3];    /* G1: coeff_a; G2: twist_mul_by_a_c{0,1,2} (mnt4753_init.cpp:125-126, mnt6753_init.cpp:138-140) */
} gctx;

static void fe_zero(fe *r) { memset(r, 0, sizeof *r); }
static void fe_one(const gctx *G, fe *r) { fe_zero(r); fp_one(G->fq, &r->c[0]); }
static int fe_is_zero(const gctx *G, const fe *a) {
    for (int i = 0; i < G->deg; ++i) if (!fp_is_zero(&a->c[i])) return 0;
    return 1;
}
static int fe_eq(const gctx *G, const fe *a, const fe *b) {
    for (int i = 0; i < G->deg; ++i) if (!fp_eq(&a->c[i], &b->c[i])) return 0;
    return 1;
}
static void fe_add(const gctx *G, fe *r, const fe *a, const fe *b) { for (int i = 0; i < G->deg; ++i) fp_add(G->fq, &r->c[i], &a->c[i], &b->c[i]); }
static void fe_sub(const gctx *G, fe *r, const fe *a, const fe *b) { for (int i = 0; i < G->deg; ++i) fp_sub(G->fq, &r->c[i], &a->c[i], &b->c[i]); }
static void fe_neg(const gctx *G, fe *r, const fe *a) { for (int i = 0; i < G->deg; ++i) fp_neg(G->fq, &r->c[i], &a->c[i]); }

static void fe_mul(const gctx *G, fe *r, const fe *x, const fe *y) {
    const fpar *P = G->fq;
    if (G->deg == 1) { fp_mul(P, &r->c[0], &x->c[0], &y->c[0]); return; }
    if (G->deg == 2) { /* Fp2_model::operator*, fp2.tcc:79-90 (Karatsuba) */
        fp aA, bB, s, t, u;
        fp_mul(P, &aA, &x->c[0], &y->c[0]);
        fp_mul(P, &bB, &x->c[1], &y->c[1]);
        fp_add(P, &s, &x->c[0], &x->c[1]);
        fp_add(P, &t, &y->c[0], &y->c[1]);
        fp_mul(P, &u, &s, &t);
        fp_sub(P, &u, &u, &aA);
        fp_sub(P, &u, &u, &bB);
        fp_mul(P, &s, &G->nr, &bB);
        fp_add(P, &r->c[0], &aA, &s);
        r->c[1] = u;
        return;
    }
    /* Fp3_model::operator*, fp3.tcc:83-97 (Karatsuba) */
    fp aA, bB, cC, s, t, u, v, w;
    fp_mul(P, &aA, &x->c[0], &y->c[0]);
    fp_mul(P, &bB, &x->c[1], &y->c[1]);
    fp_mul(P, &cC, &x->c[2], &y->c[2]);
    fp_add(P, &s, &x->c[1], &x->c[2]); fp_add(P, &t, &y->c[1], &y->c[2]);
    fp_mul(P, &u, &s, &t); fp_sub(P, &u, &u, &bB); fp_sub(P, &u, &u, &cC); /* (b+c)(B+C)-bB-cC */
    fp_add(P, &s, &x->c[0], &x->c[1]); fp_add(P, &t, &y->c[0], &y->c[1]);
    fp_mul(P, &v, &s, &t); fp_sub(P, &v, &v, &aA); fp_sub(P, &v, &v, &bB); /* (a+b)(A+B)-aA-bB */
    fp_add(P, &s, &x->c[0], &x->c[2]); fp_add(P, &t, &y->c[0], &y->c[2]);
    fp_mul(P, &w, &s, &t); fp_sub(P, &w, &w, &aA); fp_add(P, &w, &w, &bB); fp_sub(P, &w, &w, &cC); /* (a+c)(A+C)-aA+bB-cC */
    fp_mul(P, &s, &G->nr, &u); fp_add(P, &r->c[0], &aA, &s);
    fp_mul(P, &s, &G->nr, &cC); fp_add(P, &r->c[1], &v, &s);
    r->c[2] = w;
}

static void fe_sqr(const gctx *G, fe *r, const fe *x) {
    const fpar *P = G->fq;
    if (G->deg == 1) { fp_sqr(P, &r->c[0], &x->c[0]); return; }
    if (G->deg == 2) { /* Fp2_model::squared_complex, fp2.tcc:113-122 */
        fp ab, s, t, u;
        fp_mul(P, &ab, &x->c[0], &x->c[1]);
        fp_add(P, &s, &x->c[0], &x->c[1]);
        fp_mul(P, &t, &G->nr, &x->c[1]); fp_add(P, &t, &x->c[0], &t);
        fp_mul(P, &u, &s, &t);
        fp_sub(P, &u, &u, &ab);
        fp_mul(P, &s, &G->nr, &ab);
        fp_sub(P, &r->c[0], &u, &s);
        fp_add(P, &r->c[1], &ab, &ab);
        return;
    }
    /* Fp3_model::squared, fp3.tcc:107-123 (CH-SQR2) */
    fp s0, ab, s1, s2, bc, s3, s4, t;
    fp_sqr(P, &s0, &x->c[0]);
    fp_mul(P, &ab, &x->c[0], &x->c[1]); fp_add(P, &s1, &ab, &ab);
    fp_sub(P, &t, &x->c[0], &x->c[1]); fp_add(P, &t, &t, &x->c[2]); fp_sqr(P, &s2, &t);
    fp_mul(P, &bc, &x->c[1], &x->c[2]); fp_add(P, &s3, &bc, &bc);
    fp_sqr(P, &s4, &x->c[2]);
    fp_mul(P, &t, &G->nr, &s3); fp_add(P, &r->c[0], &s0, &t);
    fp_mul(P, &t, &G->nr, &s4); fp_add(P, &r->c[1], &s1, &t);
    fp_add(P, &t, &s1, &s2); fp_add(P, &t, &t, &s3); fp_sub(P, &t, &t, &s0); fp_sub(P, &r->c[2], &t, &s4);
}

static void fe_inv(const gctx *G, fe *r, const fe *x) {
    const fpar *P = G->fq;
    if (G->deg == 1) { fp_inv(P, &r->c[0], &x->c[0]); return; }
    if (G->deg == 2) { /* Fp2_model::inverse, fp2.tcc:129-142 */
        fp t0, t1, t2, t3;
        fp_sqr(P, &t0, &x->c[0]); fp_sqr(P, &t1, &x->c[1]);
        fp_mul(P, &t2, &G->nr, &t1); fp_sub(P, &t2, &t0, &t2);
        fp_inv(P, &t3, &t2);
        fp_mul(P, &r->c[0], &x->c[0], &t3);
        fp_mul(P, &t0, &x->c[1], &t3); fp_neg(P, &r->c[1], &t0);
        return;
    }
    /* Fp3_model::inverse, fp3.tcc:126-143 */
    fp t0, t1, t2, t3, t4, t5, c0, c1, c2, t6, u, v;
    const fp *a = &x->c[0], *b = &x->c[1], *c = &x->c[2];
    fp_sqr(P, &t0, a); fp_sqr(P, &t1, b); fp_sqr(P, &t2, c);
    fp_mul(P, &t3, a, b); fp_mul(P, &t4, a, c); fp_mul(P, &t5, b, c);
    fp_mul(P, &u, &G->nr, &t5); fp_sub(P, &c0, &t0, &u);
    fp_mul(P, &u, &G->nr, &t2); fp_sub(P, &c1, &u, &t3);
    fp_sub(P, &c2, &t1, &t4);
    fp_mul(P, &u, c, &c1); fp_mul(P, &v, b, &c2); fp_add(P, &u, &u, &v); fp_mul(P, &u, &G->nr, &u);
    fp_mul(P, &v, a, &c0); fp_add(P, &u, &v, &u);
    fp_inv(P, &t6, &u);
    fp_mul(P, &r->c[0], &t6, &c0); fp_mul(P, &r->c[1], &t6, &c1); fp_mul(P, &r->c[2], &t6, &c2);
}

/* coeff_a * x (G1: mnt4753_g1.cpp:323) / mul_by_a (mnt4753_g2.cpp:31-34, mnt6753_g2.cpp:38-41) */
static void fe_mul_by_a(const gctx *G, fe *r, const fe *x) {
    const fpar *P = G->fq;
    fe t;
    if (G->deg == 1) fp_mul(P, &t.c[0], &G->mba[0], &x->c[0]);
    else if (G->deg == 2) { fp_mul(P, &t.c[0], &G->mba[0], &x->c[0]); fp_mul(P, &t.c[1], &G->mba[1], &x->c[1]); }
    else { fp_mul(P, &t.c[0], &G->mba[0], &x->c[1]); fp_mul(P, &t.c[1], &G->mba[1], &x->c[2]); fp_mul(P, &t.c[2], &G->mba[2], &x->c[0]); }
    *r = t;
}

/* ---------------------------------------------------------------- curve contexts */
static gctx CTX[2][2];
static int ctx_ready = 0;
static void ctx_init(void) {
    if (ctx_ready) return;
#pragma omp critical(orc_ctx)
    {
        if (!ctx_ready) {
            static const uint64_t a2[NL] = MNT753_MONT2_A_U64, a13[NL] = MNT753_MONT13_A_U64, a26[NL] = MNT753_MONT26_A_U64;
            static const uint64_t b11[NL] = MNT753_MONT11_B_U64, b121[NL] = MNT753_MONT121_B_U64;
            memset(CTX, 0, sizeof CTX);
            gctx *g;
            g = &CTX[0][0]; g->fq = &PAR_A; g->fr = &PAR_B; g->deg = 1; memcpy(g->mba[0].l, a2, 96);
            g = &CTX[0][1]; g->fq = &PAR_A; g->fr = &PAR_B; g->deg = 2; memcpy(g->nr.l, a13, 96);
            memcpy(g->mba[0].l, a26, 96); memcpy(g->mba[1].l, a26, 96);
            g = &CTX[1][0]; g->fq = &PAR_B; g->fr = &PAR_A; g->deg = 1; memcpy(g->mba[0].l, b11, 96);
            g = &CTX[1][1]; g->fq = &PAR_B; g->fr = &PAR_A; g->deg = 3; memcpy(g->nr.l, b11, 96);
            memcpy(g->mba[0].l, b121, 96); memcpy(g->mba[1].l, b121, 96); memcpy(g->mba[2].l, b11, 96);
            ctx_ready = 1;
        }
    }
}
static const gctx *get_ctx(int curve, int group) {
    ctx_init();
    if (curve < 0 || curve > 1 || group < 1 || group > 2) return NULL;
    return &CTX[curve][group - 1];
}

/* ---------------------------------------------------------------- group law: libff homogeneous
 * projective coordinates (X/Z, Y/Z); zero = (0,1,0); is_zero <=> X == 0 && Z == 0. */
typedef struct { fe X, Y, Z; } pt;

static void pt_zero(const gctx *G, pt *r) { fe_zero(&r->X); fe_one(G, &r->Y); fe_zero(&r->Z); }
static int pt_is_zero(const gctx *G, const pt *a) { return fe_is_zero(G, &a->X) && fe_is_zero(G, &a->Z); }

/* mnt4753_G1::dbl, mnt4753_g1.cpp:305-337 (dbl-2007-bl); same shape in the G2 / mnt6753 twins */
static void pt_dbl(const gctx *G, pt *r, const pt *a) {
    if (pt_is_zero(G, a)) { *r = *a; return; }
    fe XX, ZZ, w, s, ss, sss, R, RR, B, h, t, X3, Y3;
    fe_sqr(G, &XX, &a->X);
    fe_sqr(G, &ZZ, &a->Z);
    fe_mul_by_a(G, &w, &ZZ);
    fe_add(G, &t, &XX, &XX); fe_add(G, &t, &t, &XX); fe_add(G, &w, &w, &t);
    fe_mul(G, &s, &a->Y, &a->Z); fe_add(G, &s, &s, &s);
    fe_sqr(G, &ss, &s);
    fe_mul(G, &sss, &s, &ss);
    fe_mul(G, &R, &a->Y, &s);
    fe_sqr(G, &RR, &R);
    fe_add(G, &t, &a->X, &R); fe_sqr(G, &B, &t); fe_sub(G, &B, &B, &XX); fe_sub(G, &B, &B, &RR);
    fe_sqr(G, &h, &w); fe_add(G, &t, &B, &B); fe_sub(G, &h, &h, &t);
    fe_mul(G, &X3, &h, &s);
    fe_sub(G, &t, &B, &h); fe_mul(G, &Y3, &w, &t); fe_add(G, &t, &RR, &RR); fe_sub(G, &Y3, &Y3, &t);
    r->X = X3; r->Y = Y3; r->Z = sss;
}

/* shared tail of operator+ / add / mixed_add (add-1998-cmo-2), mnt4753_g1.cpp:187-206 */
static void pt_add_tail(const gctx *G, pt *r, const fe *X1Z2, const fe *X2Z1, const fe *Y1Z2, const fe *Y2Z1, const fe *Z1Z2) {
    fe u, uu, v, vv, vvv, R, A, t, X3, Y3, Z3;
    fe_sub(G, &u, Y2Z1, Y1Z2);
    fe_sqr(G, &uu, &u);
    fe_sub(G, &v, X2Z1, X1Z2);
    fe_sqr(G, &vv, &v);
    fe_mul(G, &vvv, &v, &vv);
    fe_mul(G, &R, &vv, X1Z2);
    fe_mul(G, &A, &uu, Z1Z2); fe_add(G, &t, &vvv, &R); fe_add(G, &t, &t, &R); fe_sub(G, &A, &A, &t);
    fe_mul(G, &X3, &v, &A);
    fe_sub(G, &t, &R, &A); fe_mul(G, &Y3, &u, &t); fe_mul(G, &t, &vvv, Y1Z2); fe_sub(G, &Y3, &Y3, &t);
    fe_mul(G, &Z3, &vvv, Z1Z2);
    r->X = X3; r->Y = Y3; r->Z = Z3;
}

/* mnt4753_G1::operator+, mnt4753_g1.cpp:134-207 */
static void pt_add(const gctx *G, pt *r, const pt *a, const pt *b) {
    if (pt_is_zero(G, a)) { *r = *b; return; }
    if (pt_is_zero(G, b)) { *r = *a; return; }
    fe X1Z2, X2Z1, Y1Z2, Y2Z1, Z1Z2;
    fe_mul(G, &X1Z2, &a->X, &b->Z);
    fe_mul(G, &X2Z1, &a->Z, &b->X);
    fe_mul(G, &Y1Z2, &a->Y, &b->Z);
    fe_mul(G, &Y2Z1, &a->Z, &b->Y);
    if (fe_eq(G, &X1Z2, &X2Z1) && fe_eq(G, &Y1Z2, &Y2Z1)) { pt_dbl(G, r, a); return; }
    fe_mul(G, &Z1Z2, &a->Z, &b->Z);
    pt_add_tail(G, r, &X1Z2, &X2Z1, &Y1Z2, &Y2Z1, &Z1Z2);
}

/* mnt4753_G1::mixed_add, mnt4753_g1.cpp:254-303 (other must have Z == 1 or be zero) */
static void pt_mixed_add(const gctx *G, pt *r, const pt *a, const pt *b) {
    if (pt_is_zero(G, a)) { *r = *b; return; }
    if (pt_is_zero(G, b)) { *r = *a; return; }
    fe X2Z1, Y2Z1;
    fe_mul(G, &X2Z1, &a->Z, &b->X);
    fe_mul(G, &Y2Z1, &a->Z, &b->Y);
    if (fe_eq(G, &a->X, &X2Z1) && fe_eq(G, &a->Y, &Y2Z1)) { pt_dbl(G, r, a); return; }
    pt_add_tail(G, r, &a->X, &X2Z1, &a->Y, &Y2Z1, &a->Z);
}

static void pt_neg(const gctx *G, pt *r, const pt *a) { r->X = a->X; fe_neg(G, &r->Y, &a->Y); r->Z = a->Z; }

/* to_affine_coordinates, mnt4753_g1.cpp:68-83 */
static void pt_to_affine(const gctx *G, pt *a) {
    if (pt_is_zero(G, a)) { pt_zero(G, a); return; }
    fe zi;
    fe_inv(G, &zi, &a->Z);
    fe_mul(G, &a->X, &a->X, &zi);
    fe_mul(G, &a->Y, &a->Y, &zi);
    fe_one(G, &a->Z);
}

/* scalar_mul, algebra/curves/curve_utils.tcc:14-34 (plain double-and-add over the integer scalar) */
static void pt_scalar_mul(const gctx *G, pt *r, const pt *base, const uint64_t *k) {
    pt acc;
    pt_zero(G, &acc);
    int found = 0;
    for (long i = 64 * NL - 1; i >= 0; --i) {
        if (found) pt_dbl(G, &acc, &acc);
        if (bn_test_bit(k, (size_t)i)) { found = 1; pt_add(G, &acc, &acc, base); }
    }
    *r = acc;
}

/* ---------------------------------------------------------------- wire codecs (libsnark/serialization.hpp) */
static void rd_fe(const gctx *G, fe *r, const uint64_t *p) { fe_zero(r); for (int i = 0; i < G->deg; ++i) memcpy(r->c[i].l, p + NL * i, 96); }
static void wr_fe(const gctx *G, uint64_t *p, const fe *a) { for (int i = 0; i < G->deg; ++i) memcpy(p + NL * i, a->c[i].l, 96); }
/* read_g1/read_g2, serialization.hpp:83-111: y == 0 => zero, else (x, y, 1) */
static void rd_affine(const gctx *G, pt *r, const uint64_t *p) {
    fe x, y;
    rd_fe(G, &x, p);
    rd_fe(G, &y, p + NL * G->deg);
    if (fe_is_zero(G, &y)) { pt_zero(G, r); return; }
    r->X = x; r->Y = y; fe_one(G, &r->Z);
}
/* write_g1/write_g2, serialization.hpp:43-67 */
static void wr_affine(const gctx *G, uint64_t *p, const pt *a) {
    pt t = *a;
    if (pt_is_zero(G, &t)) { memset(p, 0, (size_t)2 * G->deg * 96); return; }
    pt_to_affine(G, &t);
    wr_fe(G, p, &t.X);
    wr_fe(G, p + NL * G->deg, &t.Y);
}
/* read_pt, libsnark/prover_reference_functions.cpp:106-115: Jacobian (X,Y,Z) -> (X*Z, Y, Z^3) */
static void rd_jacobian(const gctx *G, pt *r, const uint64_t *p) {
    fe x, y, z, zz;
    rd_fe(G, &x, p); rd_fe(G, &y, p + NL * G->deg); rd_fe(G, &z, p + 2 * NL * G->deg);
    fe_mul(G, &r->X, &x, &z);
    r->Y = y;
    fe_mul(G, &zz, &z, &z);
    fe_mul(G, &r->Z, &zz, &z);
}

/* ---------------------------------------------------------------- MSM (algebra/scalar_multiplication/multiexp.tcc) */
/* libff::log2, common/utils.cpp:32-45 (ceil log2) */
static size_t ff_log2(size_t n) {
    size_t r = ((n & (n - 1)) == 0 ? 0 : 1);
    while (n > 1) { n >>= 1; r++; }
    return r;
}

/* multi_exp_inner<naive>, multiexp.tcc:143-162 */
static void msm_naive(const gctx *G, pt *res, const pt *bases, const uint64_t *exps, size_t n) {
    pt acc, t;
    pt_zero(G, &acc);
    for (size_t i = 0; i < n; ++i) {
        pt_scalar_mul(G, &t, &bases[i], exps + i * NL);
        pt_add(G, &acc, &acc, &t);
    }
    *res = acc;
}

/* multi_exp_inner<BDLO12>, multiexp.tcc:165-282 (USE_MIXED_ADDITION is off in the reference build) */
static void msm_bdlo12(const gctx *G, pt *res, const pt *bases, const uint64_t *exps, size_t length) {
    size_t log2_length = ff_log2(length);
    size_t c = log2_length - (log2_length / 3 - 2);
    size_t num_bits = 0;
    for (size_t i = 0; i < length; ++i) {
        size_t b = bn_num_bits(exps + i * NL);
        if (b > num_bits) num_bits = b;
    }
    size_t num_groups = (num_bits + c - 1) / c;
    pt result;
    pt_zero(G, &result);
    int result_nonzero = 0;
    pt *buckets = (pt *)malloc(sizeof(pt) << c);
    unsigned char *nz = (unsigned char *)malloc((size_t)1 << c);
    for (size_t k = num_groups - 1; k <= num_groups; k--) {
        if (result_nonzero)
            for (size_t i = 0; i < c; ++i) pt_dbl(G, &result, &result);
        memset(nz, 0, (size_t)1 << c);
        for (size_t i = 0; i < length; ++i) {
            size_t id = 0;
            for (size_t j = 0; j < c; ++j)
                if (bn_test_bit(exps + i * NL, k * c + j)) id |= (size_t)1 << j;
            if (id == 0) continue;
            if (nz[id]) pt_add(G, &buckets[id], &buckets[id], &bases[i]);
            else { buckets[id] = bases[i]; nz[id] = 1; }
        }
        pt running;
        int running_nonzero = 0;
        for (size_t i = ((size_t)1 << c) - 1; i > 0; --i) {
            if (nz[i]) {
                if (running_nonzero) pt_add(G, &running, &running, &buckets[i]);
                else { running = buckets[i]; running_nonzero = 1; }
            }
            if (running_nonzero) {
                if (result_nonzero) pt_add(G, &result, &result, &running);
                else { result = running; result_nonzero = 1; }
            }
        }
    }
    free(buckets);
    free(nz);
    *res = result;
}

/* multi_exp, multiexp.tcc:402-441: split into `chunks` ranges, OpenMP, serial fold */
static void msm_chunked(const gctx *G, pt *res, const pt *bases, const uint64_t *exps, size_t total, size_t chunks, int method) {
    if (total == 0) { pt_zero(G, res); return; }
    if (total < chunks || chunks == 1) {
        if (method == 0) msm_naive(G, res, bases, exps, total); else msm_bdlo12(G, res, bases, exps, total);
        return;
    }
    size_t one = total / chunks;
    pt *partial = (pt *)malloc(sizeof(pt) * chunks);
#pragma omp parallel for
    for (size_t i = 0; i < chunks; ++i) {
        size_t lo = i * one, hi = (i == chunks - 1) ? total : (i + 1) * one;
        if (method == 0) msm_naive(G, &partial[i], bases + lo, exps + lo * NL, hi - lo);
        else msm_bdlo12(G, &partial[i], bases + lo, exps + lo * NL, hi - lo);
    }
    pt fin;
    pt_zero(G, &fin);
    for (size_t i = 0; i < chunks; ++i) pt_add(G, &fin, &fin, &partial[i]);
    free(partial);
    *res = fin;
}

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* ---------------------------------------------------------------- exported API */
int orc_num_threads(void) { return omp_get_max_threads(); }
void orc_set_num_threads(int t) { omp_set_num_threads(t); }

int orc_field_op(int curve, int field, int op, size_t n, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    const gctx *G0 = get_ctx(curve, field == 0 ? 1 : 2);
    if (!G0) return -1;
    size_t stride = (size_t)NL * G0->deg;
    for (size_t i = 0; i < n; ++i) {
        fe x, y, r;
        rd_fe(G0, &x, a + i * stride);
        if (b) rd_fe(G0, &y, b + i * stride); else fe_zero(&y);
        switch (op) {
            case 0: fe_mul(G0, &r, &x, &y); break;
            case 1: fe_add(G0, &r, &x, &y); break;
            case 2: fe_sub(G0, &r, &x, &y); break;
            case 3: fe_sqr(G0, &r, &x); break;
            case 4: if (fe_is_zero(G0, &x)) fe_zero(&r); else fe_inv(G0, &r, &x); break;
            case 5: fe_neg(G0, &r, &x); break;
            default: return -1;
        }
        wr_fe(G0, out + i * stride, &r);
    }
    return 0;
}

int orc_fr_from_mont(int curve, size_t n, const uint64_t *in, uint64_t *out) {
    const gctx *G = get_ctx(curve, 1);
    if (!G) return -1;
    for (size_t i = 0; i < n; ++i) { fp t; memcpy(t.l, in + i * NL, 96); fp_from_mont(G->fr, out + i * NL, &t); }
    return 0;
}
int orc_fr_to_mont(int curve, size_t n, const uint64_t *in, uint64_t *out) {
    const gctx *G = get_ctx(curve, 1);
    if (!G) return -1;
    for (size_t i = 0; i < n; ++i) { fp t; fp_to_mont(G->fr, &t, in + i * NL); memcpy(out + i * NL, t.l, 96); }
    return 0;
}

int orc_point_op(int curve, int group, int op, const uint64_t *a, const uint64_t *b, const uint64_t *k, uint64_t *out) {
    const gctx *G = get_ctx(curve, group);
    if (!G) return -1;
    pt A, B, R;
    rd_affine(G, &A, a);
    if (b) rd_affine(G, &B, b); else pt_zero(G, &B);
    switch (op) {
        case 0: case 5: pt_add(G, &R, &A, &B); break;
        case 1: pt_dbl(G, &R, &A); break;
        case 2: pt_mixed_add(G, &R, &A, &B); break;
        case 3: { fp s; uint64_t e[NL]; memcpy(s.l, k, 96); fp_from_mont(G->fr, e, &s); pt_scalar_mul(G, &R, &A, e); } break;
        case 4: pt_neg(G, &R, &A); break;
        default: return -1;
    }
    wr_affine(G, out, &R);
    return 0;
}

/* batch_to_special, multiexp.tcc:683-715 idea (Montgomery's trick) applied per OpenMP chunk */
int orc_gen_bases(int curve, int group, size_t n, const uint64_t *p0, const uint64_t *q, uint64_t *out) {
    const gctx *G = get_ctx(curve, group);
    if (!G) return -1;
    pt P0, Q;
    rd_affine(G, &P0, p0);
    rd_affine(G, &Q, q);
    size_t aff = (size_t)2 * G->deg * NL;
    int T = omp_get_max_threads();
    size_t per = (n + (size_t)T - 1) / (size_t)T;
    if (per < 1) per = 1;
#pragma omp parallel for
    for (int t = 0; t < T; ++t) {
        size_t lo = (size_t)t * per, hi = lo + per < n ? lo + per : n;
        if (lo >= hi) continue;
        size_t m = hi - lo;
        pt *v = (pt *)malloc(sizeof(pt) * m);
        fe *pre = (fe *)malloc(sizeof(fe) * m);
        uint64_t e[NL] = {lo};
        pt cur, t0;
        pt_scalar_mul(G, &t0, &Q, e);
        pt_add(G, &cur, &P0, &t0);
        for (size_t i = 0; i < m; ++i) { v[i] = cur; pt_mixed_add(G, &cur, &cur, &Q); }
        /* prefix products of the non-zero Z's */
        fe acc; fe_one(G, &acc);
        for (size_t i = 0; i < m; ++i) { pre[i] = acc; if (!pt_is_zero(G, &v[i])) fe_mul(G, &acc, &acc, &v[i].Z); }
        fe inv; fe_inv(G, &inv, &acc);
        for (size_t i = m; i-- > 0;) {
            if (pt_is_zero(G, &v[i])) { memset(out + (lo + i) * aff, 0, aff * 8); continue; }
            fe zi; fe_mul(G, &zi, &inv, &pre[i]);
            fe_mul(G, &inv, &inv, &v[i].Z);
            fe x, y;
            fe_mul(G, &x, &v[i].X, &zi); fe_mul(G, &y, &v[i].Y, &zi);
            wr_fe(G, out + (lo + i) * aff, &x); wr_fe(G, out + (lo + i) * aff + NL * G->deg, &y);
        }
        free(v); free(pre);
    }
    return 0;
}

double orc_msm(int curve, int group, size_t n, const uint64_t *bases, const uint64_t *scalars, uint64_t *out,
               int method, int chunks, int prefilter) {
    const gctx *G = get_ctx(curve, group);
    if (!G || method < 0 || method > 1) return -1.0;
    size_t aff = (size_t)2 * G->deg * NL;
    pt *g = (pt *)malloc(sizeof(pt) * (n ? n : 1));
    uint64_t *e = (uint64_t *)malloc(96 * (n ? n : 1));
    fp one; fp_one(G->fr, &one);
    if (chunks <= 0) chunks = omp_get_max_threads();
    double t0 = now_s();
    /* multi_exp_with_mixed_addition prefilter, multiexp.tcc:443-496: skip 0, add 1 directly */
    pt acc; pt_zero(G, &acc);
    size_t m = 0;
    for (size_t i = 0; i < n; ++i) {
        fp s; memcpy(s.l, scalars + i * NL, 96);
        if (prefilter && fp_is_zero(&s)) continue;
        pt b; rd_affine(G, &b, bases + i * aff);
        if (prefilter && fp_eq(&s, &one)) { pt_add(G, &acc, &acc, &b); continue; }
        g[m] = b;
        fp_from_mont(G->fr, e + m * NL, &s);
        ++m;
    }
    pt r;
    msm_chunked(G, &r, g, e, m, (size_t)chunks, method);
    pt_add(G, &r, &acc, &r);
    double t1 = now_s();
    wr_affine(G, out, &r);
    free(g); free(e);
    return t1 - t0;
}

int orc_msm_closed_form(int curve, int group, size_t n, const uint64_t *p0, const uint64_t *q,
                        const uint64_t *scalars, uint64_t *out) {
    const gctx *G = get_ctx(curve, group);
    if (!G) return -1;
    const fpar *R = G->fr;
    fp s0, s1; fp_zero(&s0); fp_zero(&s1);
    for (size_t i = 0; i < n; ++i) {
        fp s, idx, t; uint64_t ii[NL] = {i};
        memcpy(s.l, scalars + i * NL, 96);
        fp_add(R, &s0, &s0, &s);
        fp_to_mont(R, &idx, ii);
        fp_mul(R, &t, &idx, &s);
        fp_add(R, &s1, &s1, &t);
    }
    uint64_t e0[NL], e1[NL];
    fp_from_mont(R, e0, &s0); fp_from_mont(R, e1, &s1);
    pt P0, Q, a, b, r;
    rd_affine(G, &P0, p0); rd_affine(G, &Q, q);
    pt_scalar_mul(G, &a, &P0, e0); pt_scalar_mul(G, &b, &Q, e1);
    pt_add(G, &r, &a, &b);
    wr_affine(G, out, &r);
    return 0;
}

int orc_jacobian_to_affine(int curve, int group, const uint64_t *xyz, uint64_t *out) {
    const gctx *G = get_ctx(curve, group);
    if (!G) return -1;
    pt p; rd_jacobian(G, &p, xyz);
    wr_affine(G, out, &p);
    return 0;
}

int orc_fold_jacobian(int curve, int group, size_t n, const uint64_t *xyz, uint64_t *out) {
    const gctx *G = get_ctx(curve, group);
    if (!G) return -1;
    pt acc, p; pt_zero(G, &acc);
    for (size_t i = 0; i < n; ++i) { rd_jacobian(G, &p, xyz + i * (size_t)3 * G->deg * NL); pt_add(G, &acc, &acc, &p); }
    wr_affine(G, out, &acc);
    return 0;
}

/* ---------------------------------------------------------------- H polynomial (compute_H)
 * Restatement of compute_H<B> (cuda_prover_piecewise.cu:14-49) on libfqfft's basic_radix2_domain
 * (depends/libfqfft/libfqfft/evaluation_domain/domains/basic_radix2_domain.tcc:63-126 and
 * basic_radix2_domain_aux.tcc: _basic_serial_radix2_FFT = bit-reversal + Cooley-Tukey butterflies,
 * _multiply_by_coset), over Fr of the curve. */
static const uint64_t ROOT_A[NL] = MNT753_ROOT_OF_UNITY_A_U64, ROOT_B[NL] = MNT753_ROOT_OF_UNITY_B_U64;
static const uint64_t GEN_A[NL] = MNT753_MONT17_A_U64, GEN_B[NL] = MNT753_MONT17_B_U64;

static size_t bitreverse(size_t n, size_t l) {
    size_t r = 0;
    for (size_t k = 0; k < l; ++k) { r = (r << 1) | (n & 1); n >>= 1; }
    return r;
}
/* basic_radix2_domain_aux.tcc:_basic_serial_radix2_FFT */
static void serial_fft(const fpar *P, fp *a, size_t n, const fp *omega) {
    const size_t logn = ff_log2(n);
    for (size_t k = 0; k < n; ++k) {
        const size_t rk = bitreverse(k, logn);
        if (k < rk) { fp t = a[k]; a[k] = a[rk]; a[rk] = t; }
    }
    size_t m = 1;
    for (size_t s = 1; s <= logn; ++s) {
        fp w_m = *omega;                      /* omega^(n / 2m) by repeated squaring */
        for (size_t e = n / (2 * m); e > 1; e >>= 1) fp_sqr(P, &w_m, &w_m);
        for (size_t k = 0; k < n; k += 2 * m) {
            fp w; fp_one(P, &w);
            for (size_t j = 0; j < m; ++j) {
                fp t, u = a[k + j];
                fp_mul(P, &t, &w, &a[k + j + m]);
                fp_add(P, &a[k + j], &u, &t);
                fp_sub(P, &a[k + j + m], &u, &t);
                fp_mul(P, &w, &w, &w_m);
            }
        }
        m *= 2;
    }
}
static void fft_scale(const fpar *P, fp *a, size_t n, const fp *c) { for (size_t i = 0; i < n; ++i) fp_mul(P, &a[i], &a[i], c); }
/* _multiply_by_coset: a[i] *= g^i */
static void mul_coset(const fpar *P, fp *a, size_t n, const fp *g) {
    fp u = *g;
    for (size_t i = 1; i < n; ++i) { fp_mul(P, &a[i], &a[i], &u); fp_mul(P, &u, &u, g); }
}

int orc_compute_h(int curve, size_t d, const uint64_t *ca, const uint64_t *cb, const uint64_t *cc, uint64_t *out) {
    if (curve != 0 && curve != 1) return -1;
    const fpar *P = curve == 0 ? &PAR_B : &PAR_A;          /* Fr of the curve */
    const size_t s_adic = curve == 0 ? MNT753_TWO_ADICITY_B : MNT753_TWO_ADICITY_A;
    const size_t m = d + 1, logm = ff_log2(m);
    if (((size_t)1 << logm) != m || logm > s_adic) return -2;
    fp omega, omega_inv, g, g_inv, m_inv, z_inv, one;
    memcpy(omega.l, curve == 0 ? ROOT_B : ROOT_A, 96);
    memcpy(g.l, curve == 0 ? GEN_B : GEN_A, 96);
    fp_one(P, &one);
    for (size_t i = logm; i < s_adic; ++i) fp_sqr(P, &omega, &omega);     /* get_root_of_unity(m) */
    fp_inv(P, &omega_inv, &omega);
    fp_inv(P, &g_inv, &g);
    { uint64_t mi[NL] = {0}; mi[logm / 64] = (uint64_t)1 << (logm % 64); fp mm; fp_to_mont(P, &mm, mi); fp_inv(P, &m_inv, &mm); }
    { fp z = g; for (size_t i = 0; i < logm; ++i) fp_sqr(P, &z, &z); fp_sub(P, &z, &z, &one); fp_inv(P, &z_inv, &z); }
    fp *A = malloc(m * sizeof(fp)), *B = malloc(m * sizeof(fp)), *C = malloc(m * sizeof(fp));
    if (!A || !B || !C) { free(A); free(B); free(C); return -3; }
    memcpy(A, ca, m * 96); memcpy(B, cb, m * 96); memcpy(C, cc, m * 96);
    fp *v[3] = {A, B, C};
    for (int k = 0; k < 3; ++k) {                           /* iFFT then cosetFFT of each */
        serial_fft(P, v[k], m, &omega_inv);
        fft_scale(P, v[k], m, &m_inv);
        mul_coset(P, v[k], m, &g);
        serial_fft(P, v[k], m, &omega);
    }
    for (size_t i = 0; i < m; ++i) {                        /* H_tmp = (ca * cb - cc) / Z(g) */
        fp_mul(P, &A[i], &A[i], &B[i]);
        fp_sub(P, &A[i], &A[i], &C[i]);
        fp_mul(P, &A[i], &A[i], &z_inv);
    }
    serial_fft(P, A, m, &omega_inv);                        /* icosetFFT */
    fft_scale(P, A, m, &m_inv);
    mul_coset(P, A, m, &g_inv);
    memcpy(out, A, m * 96);
    memset(out + m * NL, 0, 96);                            /* vector_Fr_zeros(m + 1) */
    free(A); free(B); free(C);
    return 0;
}

