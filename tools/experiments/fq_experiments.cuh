// EXPERIMENTS (not part of the engine; nothing under gpu_groth16_prover_3x_b200/ includes this file).
// Two alternative forms of the Montgomery dot product fq_dot (csrc/fq.cuh), both measured slower on B200 and
// kept only for tools/mul_sched_probe.cu and the host-emulation parity test (tests/host_emu):
//   fq_dot_rolled   the row loop kept rolled in blocks of RU rows (smaller instruction footprint)
//   fq_dot_rr       reduced radix 2^28 / 2^29, carry-free IMAD.WIDE into 64-bit columns
// See DESIGN.md section 5 ("Tried and measured, not adopted").
#pragma once
#include "../../gpu_groth16_prover_3x_b200/csrc/fq.cuh"

namespace mnt753 {

// ---- the same dot product with the row loop kept ROLLED in blocks of RU rows ----------------------------
// The fully unrolled fq_dot is ~1.4k instructions (22 KB) of straight-line code; with three warps per
// scheduler at different places in it the instruction fetch cannot keep up (measured: the slab multiplier
// reaches 97 % of the IMAD.WIDE pipe with 8 warps per SM but only 86 % with 12).  Here a block of RU rows
// is unrolled (the per-row ">> 32" stays register renaming inside the block) and the block is iterated
// 24 / RU times; the loop-carried accumulators cost one register move each per block.
// b must come from memory (BQuads): a register-resident b cannot be indexed by the loop counter.
template <class M, int K, int RU, class BSrc>
MSM_DEVICE void fq_dot_rolled(fq_t &r, const uint32_t (&a)[K][NLIMB], BSrc &bsrc) {
    static_assert(RU % 4 == 0 && NLIMB % RU == 0, "rows per block: a multiple of 4 dividing 24");
    uint32_t E[NLIMB + 1], O[NLIMB];
#pragma unroll
    for (int j = 0; j <= NLIMB; ++j) E[j] = 0;
#pragma unroll
    for (int j = 0; j < NLIMB; ++j) O[j] = 0;
#pragma unroll 1
    for (int blk = 0; blk < NLIMB / RU; ++blk) {
        uint4 bq[K];
#pragma unroll
        for (int rr = 0; rr < RU; ++rr) {
            if ((rr & 3) == 0) {
#pragma unroll
                for (int k = 0; k < K; ++k) bq[k] = bsrc.quad(k, blk * (RU / 4) + (rr >> 2));
            }
            uint32_t bw[K];
#pragma unroll
            for (int k = 0; k < K; ++k) bw[k] = (rr & 3) == 0 ? bq[k].x : (rr & 3) == 1 ? bq[k].y : (rr & 3) == 2 ? bq[k].z : bq[k].w;
            uint32_t nE[NLIMB + 1], nO[NLIMB];
            const uint32_t b0 = bw[0];
            // T >>= 32 folded into this row (see fq_dot)
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
#pragma unroll
            for (int j = 0; j <= NLIMB; ++j) E[j] = nE[j];
#pragma unroll
            for (int j = 0; j < NLIMB; ++j) O[j] = nO[j];
#pragma unroll
            for (int k = 1; k < K; ++k) {
                const uint32_t bk = bw[k];
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
    }
    fq_t t;
    t[0] = prim::add_cc(O[0], E[1]);
#pragma unroll
    for (int k = 1; k < NLIMB; ++k) t[k] = prim::addc_cc(O[k], E[k + 1]);
    fq_cond_sub<M>(r, t);
}

// ---- reduced-radix dot product (EXPERIMENT, not used by the engine) ------------------------------
// Same contract as fq_dot, different instruction mix: the product is evaluated in radix 2^W, W < 32,
// so that every partial product is a carry-free IMAD.WIDE.U32 into a 64-bit column accumulator that
// cannot overflow ((K+1) * N * 2^(2W) < 2^64); operands are re-sliced with funnel shifts on the ALU
// pipe, the Montgomery digit of each row is m = T0 * (-p^-1) mod 2^W, columns are renormalised once
// at the end and a last partial step of S = 768 - N*W bits completes the division by R = 2^768, so
// results are bit-identical to the CIOS.  The idea was to escape the carry form IMAD.WIDE.U32.X.
// Measured on B200 (b200msm_microbench kinds 2/3): it does not pay.  IMAD.WIDE.U32 with or
// without carry issues on the half-rate "fmaheavy" sub-pipe (32 lanes/clk/SM), so the 1377 products of
// this form (26^2 * 2 + 25) lose to the 1152 of the 32-bit CIOS: 5.5 vs 7.7 G modmul/s.  Kept for the
// microbenchmark and the host-emulation parity test.
//   K = 1 : W = 29, 26 limbs, S = 14        K = 2, 3 : W = 28, 27 (b, p) / 28 (scaled a) limbs, S = 12
template <class M, int W>
struct Radix {
    static constexpr uint32_t MASK = (1u << W) - 1u;
    MSM_HD static constexpr uint32_t P(int j) {
        const int bit = W * j, w = bit >> 5, s = bit & 31;
        const uint64_t lo = w < NLIMB ? (uint64_t)M::P(w) : 0ull;
        const uint64_t hi = w + 1 < NLIMB ? (uint64_t)M::P(w + 1) : 0ull;
        return (uint32_t)(((hi << 32) | lo) >> s) & MASK;
    }
};

// W-bit limb j of a 768-bit little-endian word array (j is a compile-time constant after unrolling)
template <int W>
MSM_DEVICE uint32_t limb_of(const uint32_t (&x)[NLIMB], int j) {
    const int bit = W * j, w = bit >> 5, s = bit & 31;
    if (w >= NLIMB) return 0u;
    uint32_t v = x[w] >> s;
    if (s + W > 32 && w + 1 < NLIMB) v |= x[w + 1] << (32 - s);
    return v & ((1u << W) - 1u);
}

template <class M, int K, class BSrc>
MSM_DEVICE void fq_dot_rr(fq_t &r, const uint32_t (&a)[K][NLIMB], BSrc &bsrc) {
    constexpr int W = (K == 1) ? 29 : 28;
    constexpr int NB = (MNT753_NUM_BITS + W - 1) / W;      // limbs of a canonical operand and of p
    constexpr int NA = (K == 1) ? NB : (768 + W - 1) / W;  // the a-operands of a tower product may be scaled by NR
    constexpr int S = 768 - W * NB;
    constexpr uint32_t MASK = (1u << W) - 1u;
    typedef Radix<M, W> RX;
    static_assert(S > 0 && S < 32, "leftover step");

    uint32_t A[K][NA];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
        for (int j = 0; j < NA; ++j) A[k][j] = limb_of<W>(a[k], j);

    uint64_t T[NA];
#pragma unroll
    for (int j = 0; j < NA; ++j) T[j] = 0;
    uint32_t bw[K][NLIMB];
#pragma unroll
    for (int i = 0; i < NB; ++i) {
        // fetch the 128-bit quads of b that limb i reaches into and that are not here yet
        const int hi_word = ((W * i + W - 1) >> 5) < NLIMB - 1 ? ((W * i + W - 1) >> 5) : NLIMB - 1;
        const int prev_hi = i == 0 ? -1 : (((W * (i - 1) + W - 1) >> 5) < NLIMB - 1 ? ((W * (i - 1) + W - 1) >> 5) : NLIMB - 1);
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
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const uint32_t bk = limb_of<W>(bw[k], i);
#pragma unroll
            for (int j = 0; j < NA; ++j) T[j] += (uint64_t)A[k][j] * bk;
        }
        const uint32_t m = prim::opaque(((uint32_t)T[0] * (M::INV & MASK)) & MASK);
#pragma unroll
        for (int j = 0; j < NB; ++j) T[j] += (uint64_t)m * RX::P(j);
        const uint64_t c = T[0] >> W;
#pragma unroll
        for (int j = 0; j + 1 < NA; ++j) T[j] = T[j + 1];
        T[NA - 1] = 0;
        T[0] += c;
    }

    // renormalise the columns and repack into 32-bit words: t < 2^S * p + p < 2^768
    uint32_t t[NLIMB + 2];
    {
        uint64_t acc = 0, c = 0;
        int nbits = 0, wi = 0;
#pragma unroll
        for (int j = 0; j < NA; ++j) {
            const uint64_t v = T[j] + c;
            c = v >> W;
            acc |= (uint64_t)((uint32_t)v & MASK) << nbits;
            nbits += W;
            if (nbits >= 32) {
                if (wi < NLIMB) t[wi] = (uint32_t)acc;
                ++wi;
                acc >>= 32;
                nbits -= 32;
            }
        }
        acc |= c << nbits;
#pragma unroll
        for (int x = 0; x < 2; ++x) {
            if (wi < NLIMB) t[wi] = (uint32_t)acc;
            ++wi;
            acc >>= 32;
        }
    }
    // last S bits of the Montgomery division, in the 32-bit domain
    const uint32_t m2 = (t[0] * M::INV) & ((1u << S) - 1u);
    uint32_t u[NLIMB + 1];
    uint64_t cy = 0;
#pragma unroll
    for (int j = 0; j < NLIMB; ++j) {
        const uint64_t v = (uint64_t)m2 * M::P(j) + t[j] + cy;
        u[j] = (uint32_t)v;
        cy = v >> 32;
    }
    u[NLIMB] = (uint32_t)cy;
    fq_t res;
#pragma unroll
    for (int j = 0; j < NLIMB; ++j) res[j] = (u[j] >> S) | (u[j + 1] << (32 - S));
    fq_cond_sub<M>(r, res);
}

template <class M>
MSM_DEVICE void fq_mul_rr(fq_t &r, const fq_t &a, const fq_t &b) {
    uint32_t aa[1][NLIMB];
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) aa[0][i] = a[i];
    BRegs<1> src{reinterpret_cast<const uint32_t(*)[NLIMB]>(&b)};
    fq_dot_rr<M, 1>(r, aa, src);
}


}  // namespace mnt753
