// The same Montgomery dot product as fq_dot (fq.cuh), evaluated on the FP64 pipe.
//
// B200 issues DFMA at full rate (measured 62.8 / clk / SM, profiles/r01_pipe_probe.txt) while the 32x32->64
// integer multiply-add the CIOS is built on runs at 31 / clk / SM, and the two pipes are independent.  A
// double holds a 52-bit limb exactly, and the product of two of them can be split EXACTLY into its high and
// low 52 bits with two fused multiply-adds (round towards zero) and one subtraction:
//
//      h = fma_rz(a, b, 2^104)            = 2^104 + floor(a b / 2^52) * 2^52      (ulp of the binade is 2^52)
//      t = (2^104 + 2^52) - h             = 2^52 - floor(a b / 2^52) * 2^52       (exact)
//      l = fma_rz(a, b, t)                = 2^52 + (a b mod 2^52)                 (exact)
//
// Both results sit in a fixed binade, so their raw bit patterns are "constant + integer" and can be summed
// with plain 64-bit integer additions (one three-input IADD3 pair per product); the constants are removed by
// starting every column accumulator at minus the biases it is going to receive.  (Technique: Emmart, Zheng,
// Weems, "Faster modular exponentiation using double precision floating point arithmetic on the GPU", 2018.)
//
// 753-bit operands = 15 limbs of 52 bits = 780 bits, while the engine's Montgomery radix is R = 2^768
// (bit-exactness with libff).  The second operand is therefore taken as b * 2^12 -- a different bit offset in
// the limb extraction, no arithmetic -- and 15 full CIOS rows divide by 2^780:  a * (b 2^12) / 2^780 = a b / R.
// Per product: 2 * 15 * 15 * (K + 1) / 2 limb products = 3 FP64 + 2 ALU instructions each, against
// 576 (K + 1) + 24 quarter-rate IMAD.WIDE of the integer form.  Results are bit-identical to fq_dot.
#pragma once
#include "../../gpu_groth16_prover_3x_b200/csrc/fq.cuh"

#ifdef MNT753_HOST_EMU
#include <cstring>
#endif

namespace mnt753 {
namespace fp64 {

constexpr int NL = 15;                                   // 52-bit limbs
constexpr int BSHIFT = 52 * NL - 768;                    // 12: b is taken as b * 2^12
constexpr uint64_t MASK52 = (1ull << 52) - 1ull;
constexpr uint64_t BITS_2P52 = 0x4330000000000000ull;    // bit pattern of 2^52
constexpr uint64_t BITS_2P104 = 0x4670000000000000ull;   // bit pattern of 2^104
constexpr uint64_t BITS_C2 = 0x4670000000000001ull;      // 2^104 + 2^52

#ifdef MNT753_HOST_EMU
inline double u2d(uint64_t x) { double d; memcpy(&d, &x, 8); return d; }
inline uint64_t d2u(double d) { uint64_t x; memcpy(&x, &d, 8); return x; }
// exact emulation for integer-valued arguments (a, b >= 0 below 2^53, |c| < 2^106): no dependence on libm or on
// the rounding mode of the host
inline double fma_rz(double a, double b, double c) {
    const unsigned __int128 p = (unsigned __int128)(uint64_t)a * (uint64_t)b;
    __int128 s = (__int128)p + (__int128)c;
    const bool neg = s < 0;
    unsigned __int128 m = neg ? (unsigned __int128)(-s) : (unsigned __int128)s;
    int top = -1;
    for (int i = 127; i >= 0; --i) if ((m >> i) & 1) { top = i; break; }
    if (top > 52) m &= ~(((unsigned __int128)1 << (top - 52)) - 1);
    const double r = (double)m;   // at most 53 significant bits: exact
    return neg ? -r : r;
}
inline double dsub(double a, double b) { return a - b; }
#else
MSM_DEVICE double u2d(uint64_t x) { return __longlong_as_double((long long)x); }
MSM_DEVICE uint64_t d2u(double d) { return (uint64_t)__double_as_longlong(d); }
MSM_DEVICE double fma_rz(double a, double b, double c) { return __fma_rz(a, b, c); }
MSM_DEVICE double dsub(double a, double b) { return __dsub_rn(a, b); }
#endif

// 52-bit limbs of the modulus and -p^-1 mod 2^52, as compile-time constants
template <class M>
struct Mod52 {
    MSM_HD static constexpr uint64_t word(int w) { return (w >= 0 && w < NLIMB) ? (uint64_t)M::P(w < 0 ? 0 : (w < NLIMB ? w : 0)) : 0ull; }
    MSM_HD static constexpr uint64_t P(int j) {
        const int bit = 52 * j, w = bit >> 5, s = bit & 31;
        const uint64_t lo = (word(w) | (word(w + 1) << 32)) >> s;
        const uint64_t hi = s ? (word(w + 2) << (64 - s)) : 0ull;
        return (lo | hi) & MASK52;
    }
    MSM_HD static constexpr uint64_t pinv() {
        const uint64_t p0 = word(0) | (word(1) << 32);
        uint64_t x = 1;
        for (int i = 0; i < 6; ++i) x *= 2ull - p0 * x;    // Newton: p0 * x = 1 mod 2^64
        return (0ull - x) & MASK52;
    }
};

// limb j of (x << SHIFT), x a 768-bit little-endian word array; j is a compile-time constant after unrolling
template <int SHIFT>
MSM_DEVICE uint64_t limb52(const uint32_t (&x)[NLIMB], int j) {
    const int bit = 52 * j - SHIFT;
    if (bit < 0) return ((((uint64_t)x[1] << 32) | x[0]) << (-bit)) & MASK52;
    const int w = bit >> 5, s = bit & 31;
    const uint64_t w0 = w < NLIMB ? x[w < NLIMB ? w : 0] : 0u;
    const uint64_t w1 = w + 1 < NLIMB ? x[w + 1 < NLIMB ? w + 1 : 0] : 0u;
    uint64_t v = (w0 | (w1 << 32)) >> s;
    if (s + 52 > 64 && w + 2 < NLIMB) v |= (uint64_t)x[w + 2 < NLIMB ? w + 2 : 0] << (64 - s);
    return v & MASK52;
}
MSM_DEVICE double int52_to_double(uint64_t v) { return dsub(u2d(v | BITS_2P52), 4503599627370496.0); }

// exact split of a * b (integers below 2^52 held in doubles): adds the raw patterns to the two columns
MSM_DEVICE void split_mul(double a, double b, uint64_t &hbits, uint64_t &lbits) {
    const double h = fma_rz(a, b, u2d(BITS_2P104));
    const double t = dsub(u2d(BITS_C2), h);
    const double l = fma_rz(a, b, t);
    hbits = d2u(h);
    lbits = d2u(l);
}

}  // namespace fp64

// r = (sum_{k<K} a[k] * b[k]) * R^-1 mod p, canonical -- same contract as fq_dot.
template <class M, int K, class BSrc>
MSM_DEVICE void fq_dot_fp(fq_t &r, const uint32_t (&a)[K][NLIMB], BSrc &bsrc) {
    using namespace fp64;
    typedef Mod52<M> P;
    constexpr uint64_t BL = BITS_2P52, BH = BITS_2P104;
    constexpr uint64_t NT = (uint64_t)(K + 1);        // l- (and h-) terms per column position and row

    double A[K][NL];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int j = 0; j < NL; ++j) A[k][j] = int52_to_double(limb52<0>(a[k], j));

    // column window: c[q] is absolute column i + q in row i.  Every accumulator starts at minus the sum of the
    // biases it will receive while it travels from its entry position down to position 0 (or to the end).
    uint64_t c[NL + 1];
#pragma unroll
    for (int q = 0; q <= NL; ++q) {
        const uint64_t nl = q < NL ? (uint64_t)(q + 1) : (uint64_t)(NL - 1), nh = (uint64_t)q;
        c[q] = 0ull - NT * (nl * BL + nh * BH);
    }

    uint32_t bw[K][NLIMB];
#pragma unroll
    for (int i = 0; i < NL; ++i) {
        // words of b that limb i of (b << 12) reaches into and that are not here yet
        const int hi_bit = 52 * i + 51 - BSHIFT;
        const int hi_word = (hi_bit >> 5) < NLIMB - 1 ? (hi_bit >> 5) : NLIMB - 1;
        const int prev_bit = 52 * (i - 1) + 51 - BSHIFT;
        const int prev_hi = i == 0 ? -1 : ((prev_bit >> 5) < NLIMB - 1 ? (prev_bit >> 5) : NLIMB - 1);
#pragma unroll
        for (int q = 0; q < NLIMB / 4; ++q) {
            if (4 * q <= hi_word && 4 * q > prev_hi) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const uint4 v = bsrc.quad(k, q);
                    bw[k][4 * q] = v.x; bw[k][4 * q + 1] = v.y; bw[k][4 * q + 2] = v.z; bw[k][4 * q + 3] = v.w;
                }
            }
        }
        // ---- product row: c += sum_k A[k] * B[k][i]
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const double bd = int52_to_double(limb52<BSHIFT>(bw[k], i));
            uint64_t hprev = 0;
#pragma unroll
            for (int j = 0; j < NL; ++j) {
                uint64_t hb, lb;
                split_mul(A[k][j], bd, hb, lb);
                c[j] += lb + hprev;
                hprev = hb;
            }
            c[NL] += hprev;
        }
        // ---- Montgomery digit of the row: m = c[0] * (-p^-1) mod 2^52 (one l-term of column 0 is still to come)
        const uint64_t v0 = (c[0] + BL) & MASK52;
        uint64_t mh, ml;
        split_mul(int52_to_double(v0), (double)P::pinv(), mh, ml);
        const double md = dsub(u2d(ml), 4503599627370496.0);
        // ---- reduction row: c += m * p
        {
            uint64_t hprev = 0;
#pragma unroll
            for (int j = 0; j < NL; ++j) {
                uint64_t hb, lb;
                split_mul(md, (double)P::P(j), hb, lb);
                c[j] += lb + hprev;
                hprev = hb;
            }
            c[NL] += hprev;
        }
        // ---- column 0 is now a multiple of 2^52: carry it up and slide the window
        c[1] += c[0] >> 52;
#pragma unroll
        for (int q = 0; q < NL; ++q) c[q] = c[q + 1];
        // the column entering in row i + 1 travels positions 15 .. i + 2
        c[NL] = 0ull - NT * ((uint64_t)(NL - 2 - i) * BL + (uint64_t)(NL - 1 - i) * BH);
    }

    // carry propagation, repack to 32-bit words, final conditional subtraction (t < 2p)
    uint64_t L[NL];
    uint64_t carry = 0;
#pragma unroll
    for (int q = 0; q < NL; ++q) {
        const uint64_t v = c[q] + carry;
        L[q] = v & MASK52;
        carry = v >> 52;
    }
    fq_t t;
#pragma unroll
    for (int w = 0; w < NLIMB; ++w) {
        const int bit = 32 * w, q = bit / 52, s = bit % 52;
        uint64_t v = L[q] >> s;
        if (s + 32 > 52 && q + 1 < NL) v |= L[q + 1] << (52 - s);
        t[w] = (uint32_t)v;
    }
    fq_cond_sub<M>(r, t);
}

template <class M>
MSM_DEVICE void fq_mul_fp(fq_t &r, const fq_t &a, const fq_t &b) {
    uint32_t aa[1][NLIMB];
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) aa[0][i] = a[i];
    BRegs<1> src{reinterpret_cast<const uint32_t(*)[NLIMB]>(&b)};
    fq_dot_fp<M, 1>(r, aa, src);
}

}  // namespace mnt753
