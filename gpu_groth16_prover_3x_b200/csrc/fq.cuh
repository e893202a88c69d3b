// 753-bit Montgomery arithmetic on 24 x 32-bit limbs held in registers (R = 2^768).
//
// Replaces multiexp/arith.cu:219-363 (Fp<modulus_info>: add/sub/neg/mul CIOS on one 64-bit limb per
// lane of a 16-lane tile) and multiexp/fixnum.cu / primitives.cu of the reference.  Here one
// thread owns a whole element; the product loop is a CIOS whose 32x32->64 products are accumulated
// into two interleaved 64-bit-aligned accumulators ("even" at limb 0, "odd" at limb 1) so that
// every product is ONE IMAD.WIDE.U32.X on an add-with-carry chain, the reduction multiplies by
// compile-time modulus limbs (immediates), and the per-row >>32 is register renaming.
//
// fq_dot computes  (sum_k a_k * b_k) / R  mod p  for K <= 3 operand pairs with a single
// interleaved reduction ("lazy reduction"): that is what the Fq2/Fq3 towers in fe.cuh are built
// from (one dot product per output coefficient).
#pragma once
#include "mnt753_constants.h"
#include "prim.cuh"

namespace mnt753 {

constexpr int NLIMB = 24;

// modulus A: Fq(MNT4753) = Fr(MNT6753); modulus B: Fr(MNT4753) = Fq(MNT6753)
struct ModA {
    MSM_HD static constexpr uint32_t P(int j) { constexpr uint32_t t[NLIMB] = MNT753_MOD_A_U32; return t[j]; }
    MSM_HD static constexpr uint32_t R1(int j) { constexpr uint32_t t[NLIMB] = MNT753_R1_A_U32; return t[j]; }
    MSM_HD static constexpr uint32_t R2(int j) { constexpr uint32_t t[NLIMB] = MNT753_R2_A_U32; return t[j]; }
    MSM_HD static constexpr uint32_t ROOT(int j) { constexpr uint32_t t[NLIMB] = MNT753_ROOT_OF_UNITY_A_U32; return t[j]; }
    MSM_HD static constexpr uint32_t G17(int j) { constexpr uint32_t t[NLIMB] = MNT753_MONT17_A_U32; return t[j]; }
    static constexpr int TWO_ADICITY = MNT753_TWO_ADICITY_A;
    static constexpr uint32_t INV = MNT753_INV_A_U32;
};
struct ModB {
    MSM_HD static constexpr uint32_t P(int j) { constexpr uint32_t t[NLIMB] = MNT753_MOD_B_U32; return t[j]; }
    MSM_HD static constexpr uint32_t R1(int j) { constexpr uint32_t t[NLIMB] = MNT753_R1_B_U32; return t[j]; }
    MSM_HD static constexpr uint32_t R2(int j) { constexpr uint32_t t[NLIMB] = MNT753_R2_B_U32; return t[j]; }
    MSM_HD static constexpr uint32_t ROOT(int j) { constexpr uint32_t t[NLIMB] = MNT753_ROOT_OF_UNITY_B_U32; return t[j]; }
    MSM_HD static constexpr uint32_t G17(int j) { constexpr uint32_t t[NLIMB] = MNT753_MONT17_B_U32; return t[j]; }
    static constexpr int TWO_ADICITY = MNT753_TWO_ADICITY_B;
    static constexpr uint32_t INV = MNT753_INV_B_U32;
};

typedef uint32_t fq_t[NLIMB];

MSM_DEVICE bool fq_is_zero(const fq_t &a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) o |= a[i];
    return o == 0;
}

// r = a - p if a >= p else a        (a < 2p)
template <class M>
MSM_DEVICE void fq_cond_sub(fq_t &r, const fq_t &a) {
    fq_t t;
    t[0] = prim::sub_cc(a[0], M::P(0));
#pragma unroll
    for (int i = 1; i < NLIMB; ++i) t[i] = prim::subc_cc(a[i], M::P(i));
    uint32_t borrow = prim::subc(0, 0);  // 0xffffffff when a < p
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) r[i] = borrow ? a[i] : t[i];
}

// modular add: a, b canonical -> canonical (reference: arith.cu:258-267)
template <class M>
MSM_DEVICE void fq_add(fq_t &r, const fq_t &a, const fq_t &b) {
    fq_t s;
    s[0] = prim::add_cc(a[0], b[0]);
#pragma unroll
    for (int i = 1; i < NLIMB; ++i) s[i] = prim::addc_cc(a[i], b[i]);
    // 2p < 2^754 so no carry out of limb 23
    fq_cond_sub<M>(r, s);
}

// modular sub (reference: arith.cu:276-285)
template <class M>
MSM_DEVICE void fq_sub(fq_t &r, const fq_t &a, const fq_t &b) {
    fq_t d;
    d[0] = prim::sub_cc(a[0], b[0]);
#pragma unroll
    for (int i = 1; i < NLIMB; ++i) d[i] = prim::subc_cc(a[i], b[i]);
    uint32_t borrow = prim::subc(0, 0);
    // add back p masked by the borrow
    r[0] = prim::add_cc(d[0], M::P(0) & borrow);
#pragma unroll
    for (int i = 1; i < NLIMB; ++i) r[i] = prim::addc_cc(d[i], M::P(i) & borrow);
}

// modular negation with 0 -> 0 (libff Fp_model::operator-; the reference kernel maps 0 -> p,
// arith.cu:269-274, which is non-canonical and deliberately not inherited)
template <class M>
MSM_DEVICE void fq_neg(fq_t &r, const fq_t &a) {
    uint32_t nz = fq_is_zero(a) ? 0u : 0xffffffffu;
    fq_t t;
    t[0] = prim::sub_cc(M::P(0), a[0]);
#pragma unroll
    for (int i = 1; i < NLIMB; ++i) t[i] = prim::subc_cc(M::P(i), a[i]);
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) r[i] = t[i] & nz;
}

// r = k * a mod p for a small compile-time k (curve coefficients 2, 11, 26, 121): left-to-right
// double-and-add with modular adds (replaces the add-chain multipliers of arith.cu:81-216)
template <class M, unsigned K>
MSM_DEVICE void fq_mul_small(fq_t &r, const fq_t &a) {
    static_assert(K >= 1, "k >= 1");
    int top = 31;
    while (!((K >> top) & 1)) --top;
    fq_t acc;
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) acc[i] = a[i];
#pragma unroll
    for (int bit = top - 1; bit >= 0; --bit) {
        fq_add<M>(acc, acc, acc);
        if ((K >> bit) & 1) fq_add<M>(acc, acc, a);
    }
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) r[i] = acc[i];
}

// r = c * a as a plain integer (NOT reduced); c <= 22 and a < p < 2^753 so the result is < 2^758
// and still fits 24 limbs.  Used to fold the tower non-residue into one operand of a dot product.
MSM_DEVICE void fq_scale_unreduced(fq_t &r, const fq_t &a, uint32_t c) {
    uint32_t hi[NLIMB];
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) hi[i] = prim::mul_hi(a[i], c);
    r[0] = prim::mul_lo(a[0], c);
    r[1] = prim::mad_lo_cc(a[1], c, hi[0]);
#pragma unroll
    for (int i = 2; i < NLIMB; ++i) r[i] = prim::madc_lo_cc(a[i], c, hi[i - 1]);
}

// ---- multiplicand limb sources for fq_dot --------------------------------------------------
// K operands whose limbs sit in "quad-interleaved" memory: quad q (limbs 4q..4q+3) of operand k is
// the uint4 at p[k][q * stride].  Shared-memory slabs use stride 32 (one uint4 per lane), see fe.cuh.
template <int K>
struct BQuads {
    const uint4 *p[K];
    int stride;
    uint4 cur[K];
    MSM_DEVICE uint32_t get(int k, int i) {
        if ((i & 3) == 0) cur[k] = p[k][(i >> 2) * stride];
        return (i & 3) == 0 ? cur[k].x : (i & 3) == 1 ? cur[k].y : (i & 3) == 2 ? cur[k].z : cur[k].w;
    }
    MSM_DEVICE uint4 quad(int k, int q) const { return p[k][q * stride]; }
};
// K operands already in registers
template <int K>
struct BRegs {
    const uint32_t (*b)[NLIMB];
    MSM_DEVICE uint32_t get(int k, int i) const { return b[k][i]; }
    MSM_DEVICE uint4 quad(int k, int q) const {
        uint4 v;
        v.x = b[k][4 * q]; v.y = b[k][4 * q + 1]; v.z = b[k][4 * q + 2]; v.w = b[k][4 * q + 3];
        return v;
    }
};

// r = (sum_{k<K} a[k] * b[k]) * R^-1 mod p, canonical.
// Preconditions: sum_k a[k]*b[k] < p * R (true for canonical b and a[k] <= 13p, K <= 3).
// Accumulator invariant: T = E + (O << 32), E has 25 words (aligned at limb 0), O has 24 words
// (aligned at limb 1); T < 2^800 throughout, so neither chain ever carries out of its top word.
template <class M, int K, class BSrc>
MSM_DEVICE void fq_dot(fq_t &r, const uint32_t (&a)[K][NLIMB], BSrc &bsrc) {
    uint32_t E[NLIMB + 1], O[NLIMB];
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) {
        uint32_t nE[NLIMB + 1], nO[NLIMB];
        const uint32_t b0 = bsrc.get(0, i);
        if (i == 0) {
#pragma unroll
            for (int j = 0; j < NLIMB; j += 2) {
                nE[j] = prim::mul_lo(a[0][j], b0);
                nE[j + 1] = prim::mul_hi(a[0][j], b0);
                nO[j] = prim::mul_lo(a[0][j + 1], b0);
                nO[j + 1] = prim::mul_hi(a[0][j + 1], b0);
            }
            nE[NLIMB] = 0;
        } else {
            // T >>= 32 folded into this row: new E = old O with old E[1] added at word 0 (its
            // carry has the weight of new O[0], where the next chain starts), new O[k] = old E[k+2].
            nE[0] = prim::add_cc(O[0], E[1]);
#pragma unroll
            for (int j = 1; j < NLIMB; j += 2) {
                nO[j - 1] = prim::madc_lo_cc(a[0][j], b0, E[j + 1]);
                nO[j] = prim::madc_hi_cc(a[0][j], b0, (j + 2 <= NLIMB) ? E[(j + 2 <= NLIMB) ? j + 2 : 0] : 0u);
            }
            nE[0] = prim::mad_lo_cc(a[0][0], b0, nE[0]);
            nE[1] = prim::madc_hi_cc(a[0][0], b0, O[1]);
#pragma unroll
            for (int j = 2; j < NLIMB; j += 2) {
                nE[j] = prim::madc_lo_cc(a[0][j], b0, O[j]);
                nE[j + 1] = prim::madc_hi_cc(a[0][j], b0, O[j + 1]);
            }
            nE[NLIMB] = prim::addc(0, 0);
        }
#pragma unroll
        for (int j = 0; j <= NLIMB; ++j) E[j] = nE[j];
#pragma unroll
        for (int j = 0; j < NLIMB; ++j) O[j] = nO[j];

#pragma unroll
        for (int k = 1; k < K; ++k) {
            const uint32_t bk = bsrc.get(k, i);
            E[0] = prim::mad_lo_cc(a[k][0], bk, E[0]);
            E[1] = prim::madc_hi_cc(a[k][0], bk, E[1]);
#pragma unroll
            for (int j = 2; j < NLIMB; j += 2) {
                E[j] = prim::madc_lo_cc(a[k][j], bk, E[j]);
                E[j + 1] = prim::madc_hi_cc(a[k][j], bk, E[j + 1]);
            }
            E[NLIMB] = prim::addc(E[NLIMB], 0);
            O[0] = prim::mad_lo_cc(a[k][1], bk, O[0]);
            O[1] = prim::madc_hi_cc(a[k][1], bk, O[1]);
#pragma unroll
            for (int j = 3; j < NLIMB; j += 2) {
                O[j - 1] = prim::madc_lo_cc(a[k][j], bk, O[j - 1]);
                O[j] = prim::madc_hi_cc(a[k][j], bk, O[j]);
            }
        }

        // Montgomery step: make the low limb vanish
        const uint32_t m = prim::mul_lo(E[0], M::INV);
        E[0] = prim::mad_lo_cc(m, M::P(0), E[0]);
        E[1] = prim::madc_hi_cc(m, M::P(0), E[1]);
#pragma unroll
        for (int j = 2; j < NLIMB; j += 2) {
            E[j] = prim::madc_lo_cc(m, M::P(j), E[j]);
            E[j + 1] = prim::madc_hi_cc(m, M::P(j), E[j + 1]);
        }
        E[NLIMB] = prim::addc(E[NLIMB], 0);
        O[0] = prim::mad_lo_cc(m, M::P(1), O[0]);
        O[1] = prim::madc_hi_cc(m, M::P(1), O[1]);
#pragma unroll
        for (int j = 3; j < NLIMB; j += 2) {
            O[j - 1] = prim::madc_lo_cc(m, M::P(j), O[j - 1]);
            O[j] = prim::madc_hi_cc(m, M::P(j), O[j]);
        }
    }
    // final >>32 and merge: t[k] = O[k] + E[k+1]
    fq_t t;
    t[0] = prim::add_cc(O[0], E[1]);
#pragma unroll
    for (int k = 1; k < NLIMB; ++k) t[k] = prim::addc_cc(O[k], E[k + 1]);
    fq_cond_sub<M>(r, t);
}


// plain Montgomery product of two register operands
template <class M>
MSM_DEVICE void fq_mul(fq_t &r, const fq_t &a, const fq_t &b) {
    uint32_t aa[1][NLIMB];
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) aa[0][i] = a[i];
    BRegs<1> src{reinterpret_cast<const uint32_t(*)[NLIMB]>(&b)};
    fq_dot<M, 1>(r, aa, src);
}

// Montgomery -> plain integer (multiply by the integer 1), reference: Fr::from_monty, arith.cu:356-362
template <class M>
MSM_DEVICE void fq_from_mont(fq_t &r, const fq_t &a) {
    fq_t one;
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) one[i] = (i == 0) ? 1u : 0u;
    fq_mul<M>(r, a, one);
}

#ifdef MNT753_HOST_EMU
#define MSM_COLD inline
#else
#define MSM_COLD __device__ __noinline__
#endif

// ---- modular inversion (binary extended Euclid, one thread) ------------------------------------
// The reference has no inversion on the device (multiexp/arith.cu:347-354 is #if 0); the CPU side
// uses libff's Fp_model::invert (fp.tcc, mpn_gcdext).  Here it serves Montgomery's simultaneous
// inversion in the batched-affine bucket accumulation: ONE lane per team inverts ONE product per tile,
// on the ALU pipe (shifts / add-with-carry), while the other warps keep the multiplier pipe busy.
//
// Right-shift binary gcd on (u, v) with cofactors (x1, x2): u = x1 * a, v = x2 * a (mod p) throughout.
// Zeros are stripped k <= 31 bits at a time; the cofactor is divided by 2^k mod p Montgomery-style:
// x <- (x + m p) / 2^k with m = x * (-p^-1) mod 2^k, which stays below p.
MSM_DEVICE bool fq_is_one(const uint32_t (&a)[NLIMB]) {
    uint32_t o = a[0] ^ 1u;
#pragma unroll
    for (int i = 1; i < NLIMB; ++i) o |= a[i];
    return o == 0;
}

template <class M>
MSM_DEVICE void fq_strip_twos(uint32_t (&u)[NLIMB], uint32_t (&x)[NLIMB]) {
    while (!(u[0] & 1u)) {
        int k = 31;
        if (u[0] != 0u) {
#ifdef MNT753_HOST_EMU
            k = __builtin_ctz(u[0]);
#else
            k = __ffs((int)u[0]) - 1;
#endif
        }
        // u >>= k
#pragma unroll
        for (int i = 0; i < NLIMB; ++i) {
            const uint32_t hi = (i + 1 < NLIMB) ? u[i + 1] : 0u;
            u[i] = (u[i] >> k) | (hi << (32 - k));     // 1 <= k <= 31
        }
        // x <- (x + m p) >> k
        const uint32_t m = (x[0] * M::INV) & ((1u << k) - 1u);
        uint32_t t[NLIMB + 1];
        uint32_t carry = 0;
#pragma unroll
        for (int i = 0; i < NLIMB; ++i) {
            const uint32_t lo = prim::mul_lo(m, M::P(i)), hi = prim::mul_hi(m, M::P(i));
            const uint32_t s1 = prim::add_cc(x[i], lo);
            const uint32_t c1 = prim::addc(hi, 0);       // hi + carry(x + lo) <= 2^32 - 1
            const uint32_t s2 = prim::add_cc(s1, carry);
            carry = prim::addc(c1, 0);
            t[i] = s2;
        }
        t[NLIMB] = carry;
#pragma unroll
        for (int i = 0; i < NLIMB; ++i) x[i] = (t[i] >> k) | (t[i + 1] << (32 - k));
    }
}

// x <- x - y mod p   (x, y < p)
template <class M>
MSM_DEVICE void fq_sub_inplace(uint32_t (&x)[NLIMB], const uint32_t (&y)[NLIMB]) {
    fq_t a, b;
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) { a[i] = x[i]; b[i] = y[i]; }
    fq_sub<M>(a, a, b);
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) x[i] = a[i];
}

// r = a^-1 mod p as plain integers, 0 < a < p.  Returns false (r = 0) for a = 0.
template <class M>
MSM_DEVICE bool fq_inv_plain(fq_t &r, const fq_t &a) {
    uint32_t u[NLIMB], v[NLIMB], x1[NLIMB], x2[NLIMB];
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) { u[i] = a[i]; v[i] = M::P(i); x1[i] = (i == 0) ? 1u : 0u; x2[i] = 0u; r[i] = 0u; }
    if (fq_is_zero(a)) return false;
    fq_strip_twos<M>(u, x1);
    for (int iter = 0; iter < 2 * NLIMB * 32 + 8; ++iter) {
        if (fq_is_one(u)) {
#pragma unroll
            for (int i = 0; i < NLIMB; ++i) r[i] = x1[i];
            return true;
        }
        if (fq_is_one(v)) {
#pragma unroll
            for (int i = 0; i < NLIMB; ++i) r[i] = x2[i];
            return true;
        }
        // d = u - v
        uint32_t d[NLIMB];
        d[0] = prim::sub_cc(u[0], v[0]);
#pragma unroll
        for (int i = 1; i < NLIMB; ++i) d[i] = prim::subc_cc(u[i], v[i]);
        const uint32_t borrow = prim::subc(0, 0);
        if (!borrow) {          // u > v (u == v only when both are 1)
#pragma unroll
            for (int i = 0; i < NLIMB; ++i) u[i] = d[i];
            fq_sub_inplace<M>(x1, x2);
            fq_strip_twos<M>(u, x1);
        } else {                // v = v - u = -d
            v[0] = prim::sub_cc(0u, d[0]);
#pragma unroll
            for (int i = 1; i < NLIMB; ++i) v[i] = prim::subc_cc(0u, d[i]);
            fq_sub_inplace<M>(x2, x1);
            fq_strip_twos<M>(v, x2);
        }
    }
    return false;
}


// ---- fast inversion: binary gcd on 64-bit approximations (after T. Pornin, "Optimized Binary GCD for
// Modular Inversion", 2020) ----------------------------------------------------------------------
// Thirty gcd steps at a time are run on 64-bit approximations of (a, b) -- the exact low 31 bits and the
// top 33 bits at the common bit length -- recording the 2x2 update matrix (f0 g0; f1 g1), |f|+|g| <= 2^30;
// the matrix is then applied once to the 753-bit values (a, b) and, with a Montgomery-style division by
// 2^30, to the cofactors (u, v) kept modulo p:  a = y u, b = y v (mod p).  A wrong comparison on the
// approximations can only make a or b come out negative, which is repaired by negating the row.
// About 1.5k instructions per outer step instead of ~13k for thirty steps of the plain algorithm above;
// that one remains the fallback should the final check b == 1 ever fail.
MSM_DEVICE void fq_mul_small(uint32_t (&r)[NLIMB + 1], const uint32_t (&x)[NLIMB], uint32_t f) {
    uint32_t carry = 0;
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) {
        const uint64_t t = prim::mad_wide(x[i], f, (uint64_t)carry);
        r[i] = (uint32_t)t;
        carry = (uint32_t)(t >> 32);
    }
    r[NLIMB] = carry;
}

// r = |f a + g b| / 2^30 (the division is exact); returns true when f a + g b < 0
MSM_DEVICE bool fq_lin_ab(uint32_t (&r)[NLIMB], const uint32_t (&a)[NLIMB], const uint32_t (&b)[NLIMB], int f, int g) {
    const bool sf = f < 0, sg = g < 0;
    const uint32_t fa = (uint32_t)(sf ? -f : f), ga = (uint32_t)(sg ? -g : g);
    uint32_t P[NLIMB + 1], Q[NLIMB + 1], t[NLIMB + 1];
    fq_mul_small(P, a, fa);
    fq_mul_small(Q, b, ga);
    bool neg;
    if (sf == sg) {
        t[0] = prim::add_cc(P[0], Q[0]);
#pragma unroll
        for (int i = 1; i <= NLIMB; ++i) t[i] = prim::addc_cc(P[i], Q[i]);
        neg = sf;
    } else {
        t[0] = prim::sub_cc(P[0], Q[0]);
#pragma unroll
        for (int i = 1; i <= NLIMB; ++i) t[i] = prim::subc_cc(P[i], Q[i]);
        const uint32_t borrow = prim::subc(0, 0);
        neg = sf;
        if (borrow) {
            neg = sg;
            t[0] = prim::sub_cc(0u, t[0]);
#pragma unroll
            for (int i = 1; i <= NLIMB; ++i) t[i] = prim::subc_cc(0u, t[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) r[i] = (t[i] >> 30) | (t[i + 1] << 2);
    return neg;
}

// r = (f u + g v) / 2^30 mod p for u, v < p and |f| + |g| <= 2^30
template <class M>
MSM_DEVICE void fq_lin_uv(uint32_t (&r)[NLIMB], const uint32_t (&u)[NLIMB], const uint32_t (&v)[NLIMB], int f, int g) {
    const bool sf = f < 0, sg = g < 0;
    const uint32_t fa = (uint32_t)(sf ? -f : f), ga = (uint32_t)(sg ? -g : g);
    fq_t U, V;
    if (sf) fq_neg<M>(U, u); else {
#pragma unroll
        for (int i = 0; i < NLIMB; ++i) U[i] = u[i];
    }
    if (sg) fq_neg<M>(V, v); else {
#pragma unroll
        for (int i = 0; i < NLIMB; ++i) V[i] = v[i];
    }
    uint32_t P[NLIMB + 1], Q[NLIMB + 1], t[NLIMB + 1];
    fq_mul_small(P, U, fa);
    fq_mul_small(Q, V, ga);
    t[0] = prim::add_cc(P[0], Q[0]);
#pragma unroll
    for (int i = 1; i <= NLIMB; ++i) t[i] = prim::addc_cc(P[i], Q[i]);
    const uint32_t q = (t[0] * M::INV) & 0x3fffffffu;
    uint32_t mp[NLIMB];
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) mp[i] = M::P(i);
    fq_mul_small(P, mp, q);
    t[0] = prim::add_cc(t[0], P[0]);
#pragma unroll
    for (int i = 1; i <= NLIMB; ++i) t[i] = prim::addc_cc(t[i], P[i]);
    fq_t w;
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) w[i] = (t[i] >> 30) | (t[i + 1] << 2);
    fq_t out;
    fq_cond_sub<M>(out, w);     // w < 2p
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) r[i] = out[i];
}

template <class M>
MSM_DEVICE bool fq_inv_plain_fast(fq_t &r, const fq_t &y) {
    uint32_t a[NLIMB], b[NLIMB], u[NLIMB], v[NLIMB];
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) { a[i] = y[i]; b[i] = M::P(i); u[i] = (i == 0) ? 1u : 0u; v[i] = 0u; r[i] = 0u; }
    if (fq_is_zero(y)) return false;
    for (int outer = 0; outer < 54; ++outer) {
        if (fq_is_zero(a)) break;
        // 64-bit approximations at the common bit length
        uint32_t ah = 0, am = 0, al = 0, bh = 0, bm = 0, bl = 0;
        bool found = false;
#pragma unroll
        for (int i = NLIMB - 1; i >= 2; --i) {
            if (!found && (a[i] | b[i]) != 0u) { ah = a[i]; am = a[i - 1]; al = a[i - 2]; bh = b[i]; bm = b[i - 1]; bl = b[i - 2]; found = true; }
        }
        uint64_t xa, xb;
        if (!found) {
            xa = ((uint64_t)a[1] << 32) | a[0];
            xb = ((uint64_t)b[1] << 32) | b[0];
        } else {
            const uint32_t top = ah | bh;
#ifdef MNT753_HOST_EMU
            const int s = __builtin_clz(top);
#else
            const int s = __clz((int)top);
#endif
            uint64_t ta = ((uint64_t)ah << 32) | am, tb = ((uint64_t)bh << 32) | bm;
            if (s) { ta = (ta << s) | (al >> (32 - s)); tb = (tb << s) | (bl >> (32 - s)); }
            xa = ((ta >> 31) << 31) | (a[0] & 0x7fffffffu);
            xb = ((tb >> 31) << 31) | (b[0] & 0x7fffffffu);
        }
        int f0 = 1, g0 = 0, f1 = 0, g1 = 1;
        for (int j = 0; j < 30; ++j) {
            if (xa & 1u) {
                if (xa < xb) {
                    const uint64_t tx = xa; xa = xb; xb = tx;
                    int ti = f0; f0 = f1; f1 = ti;
                    ti = g0; g0 = g1; g1 = ti;
                }
                xa -= xb; f0 -= f1; g0 -= g1;
            }
            xa >>= 1;
            f1 <<= 1; g1 <<= 1;
        }
        uint32_t na[NLIMB], nb[NLIMB];
        if (fq_lin_ab(na, a, b, f0, g0)) { f0 = -f0; g0 = -g0; }
        if (fq_lin_ab(nb, a, b, f1, g1)) { f1 = -f1; g1 = -g1; }
        uint32_t nu[NLIMB], nv[NLIMB];
        fq_lin_uv<M>(nu, u, v, f0, g0);
        fq_lin_uv<M>(nv, u, v, f1, g1);
#pragma unroll
        for (int i = 0; i < NLIMB; ++i) { a[i] = na[i]; b[i] = nb[i]; u[i] = nu[i]; v[i] = nv[i]; }
    }
    if (!fq_is_zero(a) || !fq_is_one(b)) return false;
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) r[i] = v[i];
    return true;
}

// Montgomery-form inverse: a = xR -> x^-1 R  (0 -> 0, like the oracle's field inversion).
// A real function on the device, with a single copy of the multiplier and a cold fallback, to keep the
// rarely executed inversion from evicting the hot loops out of the instruction cache.
template <class M>
MSM_COLD bool fq_inv_plain_cold(fq_t &r, const fq_t &a) { return fq_inv_plain<M>(r, a); }
template <class M>
MSM_COLD void fq_mul_by_r2(fq_t &r) {
    fq_t r2, x;
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) { r2[i] = M::R2(i); x[i] = r[i]; }
    fq_mul<M>(r, x, r2);
}
template <class M>
MSM_COLD void fq_inv(fq_t &r, const fq_t &a) {
    fq_t t;
    if (!fq_inv_plain_fast<M>(t, a)) fq_inv_plain_cold<M>(t, a);   // (xR)^-1 = x^-1 R^-1
    fq_mul_by_r2<M>(t);                // x^-1 R^-1 * R^2 / R = x^-1
    fq_mul_by_r2<M>(t);                // x^-1 * R^2 / R = x^-1 R
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) r[i] = t[i];
}

}  // namespace mnt753
