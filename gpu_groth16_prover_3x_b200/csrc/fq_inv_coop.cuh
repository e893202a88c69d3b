// Warp-cooperative modular inversion: ONE 753-bit inversion by the 32 lanes of a warp, limb i on lane i.
//
// The reference has no inversion on the device (multiexp/arith.cu:347-354 is #if 0; the CPU prover goes through
// libff's Fp_model::invert / mpn_gcdext, fp.tcc).  The batched-affine accumulation needs one inversion per team
// and round; done by one thread (fq_inv, fq.cuh) it is ~70 k dependent instructions during which the warp's other
// 31 lanes idle -- a tenth of all warp time of a 2^20-point MSM and most of a round's latency in a small one.
//
// Same algorithm as fq_inv_plain_fast (binary gcd on 64-bit approximations after T. Pornin, "Optimized Binary GCD
// for Modular Inversion", 2020): thirty gcd steps on approximations of (a, b) give a 2x2 matrix (f0 g0; f1 g1),
// |f| + |g| <= 2^30, which is then applied to the full-length values.  Here the thirty steps run redundantly on
// every lane (they are uniform), and the four full-length updates are limb-parallel: lane i forms its signed
// 64-bit partial sums f x_i + g y_i, the high words move one lane up with a shuffle, and the few carries that
// remain ripple through a ballot-terminated loop (in practice one or two passes).
//
//   a, b   non-negative, 24 limbs; a row that comes out negative is negated (and so is its (f, g))
//   u, v   cofactors modulo p as 27-limb two's-complement numbers in (-2p, p)  [lane 26 is the sign limb]:
//            u' = (f u + g v + k p) / 2^30,   k = [u<0] f + [v<0] g - ((p^-1 t + [u<0] f + [v<0] g) mod 2^30),
//          t the low word of f u + g v -- the update of Bernstein-Yang / Pornin style inverters that keeps the range
//          without comparing against p (f u+ + g v+ lies in (-2^30 p, 2^30 p) for u+ = u + [u<0] p, minus
//          [0, 2^30) p, over 2^30).
//   invariant  a = y u / R2,  b = y v / R2  (mod p): u starts at R^2 mod p, so for y = x R the result
//          v = R^2 / y = x^-1 R is the inverse in Montgomery form with no multiplication afterwards.
#pragma once
#include "fq.cuh"

namespace mnt753 {

namespace invc {
constexpr unsigned FULL = 0xffffffffu;
constexpr int TOPL = 26;     // sign limb of the two's-complement values

// value = sum_i (lo_i + hi_i 2^32) 2^(32 i)  ->  two's-complement limbs on lanes 0 .. TOPL (carry out of TOPL dropped)
__device__ __forceinline__ uint32_t settle(long long lo, long long hi, int lane) {
    long long cin = __shfl_up_sync(FULL, hi, 1);
    if (lane == 0) cin = 0;
    long long s = lo + cin;
    uint32_t limb = (uint32_t)s;
    int carry = (int)(s >> 32);
    if (lane >= TOPL) carry = 0;
    if (lane > TOPL) limb = 0u;
    while (__any_sync(FULL, carry != 0)) {
        int c = __shfl_up_sync(FULL, carry, 1);
        if (lane == 0 || lane > TOPL) c = 0;
        const long long t = (long long)limb + c;
        limb = (uint32_t)t;
        carry = (int)(t >> 32);
        if (lane >= TOPL) carry = 0;
    }
    return limb;
}
__device__ __forceinline__ bool negative(uint32_t limb) { return (__shfl_sync(FULL, limb, TOPL) >> 31) != 0u; }
// arithmetic shift right by 30 bits across the lanes
__device__ __forceinline__ uint32_t shr30(uint32_t limb, int lane) {
    uint32_t up = __shfl_down_sync(FULL, limb, 1);
    if (lane == TOPL) up = (uint32_t)((int32_t)limb >> 31);
    const uint32_t r = (limb >> 30) | (up << 2);
    return lane > TOPL ? 0u : r;
}
// signed coefficient x unsigned limb, split into a low word and a signed high word
__device__ __forceinline__ void acc(long long &lo, long long &hi, int k, uint32_t x) {
    const long long p = (long long)k * (long long)(unsigned long long)x;
    lo += (long long)(uint32_t)p;
    hi += (p >> 32);
}
}  // namespace invc

// x: limb `lane` of y (Montgomery form x R, canonical; lanes >= 24 ignored) in, limb of y^-1 R^2 = x^-1 R out.
// 0 -> 0.  Returns false (warp-uniform) if the gcd did not end in 1 -- the caller then falls back to fq_inv.
// Must be called by all 32 lanes of a converged warp.
template <class M>
__device__ __noinline__ bool fq_inv_coop(uint32_t &x) {
    using namespace invc;
    const int lane = threadIdx.x & 31;
    uint32_t pl = 0u, r2 = 0u;
#pragma unroll
    for (int j = 0; j < NLIMB; ++j)
        if (lane == j) { pl = M::P(j); r2 = M::R2(j); }
    const uint32_t pinv = (0u - M::INV) & 0x3fffffffu;      // p^-1 mod 2^30 (M::INV = -p^-1 mod 2^32)
    uint32_t a = lane < NLIMB ? x : 0u, b = pl, u = r2, v = 0u;
    if (!__any_sync(FULL, a != 0u)) { x = 0u; return true; }
    for (int outer = 0; outer < 54; ++outer) {
        if (!__any_sync(FULL, a != 0u)) break;
        // 64-bit approximations at the common bit length: exact low 31 bits, top 33 bits
        const unsigned m = __ballot_sync(FULL, (a | b) != 0u);
        const int top = 31 - __clz((int)m);
        const uint32_t a0 = __shfl_sync(FULL, a, 0), b0 = __shfl_sync(FULL, b, 0);
        unsigned long long xa, xb;
        if (top < 2) {
            xa = ((unsigned long long)__shfl_sync(FULL, a, 1) << 32) | a0;
            xb = ((unsigned long long)__shfl_sync(FULL, b, 1) << 32) | b0;
        } else {
            const uint32_t ah = __shfl_sync(FULL, a, top), am = __shfl_sync(FULL, a, top - 1), al = __shfl_sync(FULL, a, top - 2);
            const uint32_t bh = __shfl_sync(FULL, b, top), bm = __shfl_sync(FULL, b, top - 1), bl = __shfl_sync(FULL, b, top - 2);
            const int s = __clz((int)(ah | bh));
            unsigned long long ta = ((unsigned long long)ah << 32) | am, tb = ((unsigned long long)bh << 32) | bm;
            if (s) { ta = (ta << s) | (al >> (32 - s)); tb = (tb << s) | (bl >> (32 - s)); }
            xa = ((ta >> 31) << 31) | (a0 & 0x7fffffffu);
            xb = ((tb >> 31) << 31) | (b0 & 0x7fffffffu);
        }
        int f0 = 1, g0 = 0, f1 = 0, g1 = 1;
        for (int j = 0; j < 30; ++j) {
            if (xa & 1u) {
                if (xa < xb) {
                    const unsigned long long tx = xa; xa = xb; xb = tx;
                    int ti = f0; f0 = f1; f1 = ti;
                    ti = g0; g0 = g1; g1 = ti;
                }
                xa -= xb; f0 -= f1; g0 -= g1;
            }
            xa >>= 1;
            f1 <<= 1; g1 <<= 1;
        }
        // (a, b) <- |(f a + g b) / 2^30| row by row; a negative row flips its coefficients
        uint32_t na, nb;
        {
            long long lo = 0, hi = 0;
            acc(lo, hi, f0, a); acc(lo, hi, g0, b);
            uint32_t t = settle(lo, hi, lane);
            const bool neg = negative(t);
            t = shr30(t, lane);
            if (neg) { t = settle((long long)(uint32_t)~t + (lane == 0 ? 1 : 0), 0, lane); f0 = -f0; g0 = -g0; }
            na = t;
        }
        {
            long long lo = 0, hi = 0;
            acc(lo, hi, f1, a); acc(lo, hi, g1, b);
            uint32_t t = settle(lo, hi, lane);
            const bool neg = negative(t);
            t = shr30(t, lane);
            if (neg) { t = settle((long long)(uint32_t)~t + (lane == 0 ? 1 : 0), 0, lane); f1 = -f1; g1 = -g1; }
            nb = t;
        }
        // (u, v) <- (f u + g v + k p) / 2^30, staying in (-2p, p)
        const bool su = negative(u), sv = negative(v);
        uint32_t nu, nv;
        {
            long long lo = 0, hi = 0;
            acc(lo, hi, f0, u); acc(lo, hi, g0, v);
            const uint32_t t0 = __shfl_sync(FULL, (uint32_t)lo, 0);
            int k = (su ? f0 : 0) + (sv ? g0 : 0);
            k -= (int)((pinv * t0 + (uint32_t)k) & 0x3fffffffu);
            acc(lo, hi, k, pl);
            nu = shr30(settle(lo, hi, lane), lane);
        }
        {
            long long lo = 0, hi = 0;
            acc(lo, hi, f1, u); acc(lo, hi, g1, v);
            const uint32_t t0 = __shfl_sync(FULL, (uint32_t)lo, 0);
            int k = (su ? f1 : 0) + (sv ? g1 : 0);
            k -= (int)((pinv * t0 + (uint32_t)k) & 0x3fffffffu);
            acc(lo, hi, k, pl);
            nv = shr30(settle(lo, hi, lane), lane);
        }
        a = na; b = nb; u = nu; v = nv;
    }
    const bool ok = !__any_sync(FULL, a != 0u) && !__any_sync(FULL, b != (lane == 0 ? 1u : 0u));
    // v in (-2p, p) -> [0, p)
    for (int k = 0; k < 2; ++k)
        if (negative(v)) v = settle((long long)v + (long long)pl, 0, lane);
    x = ok ? v : 0u;
    return ok;
}

}  // namespace mnt753
