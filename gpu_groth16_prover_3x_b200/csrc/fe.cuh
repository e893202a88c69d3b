// Extension-field elements (Fq, Fq2 = Fq[u]/(u^2-13), Fq3 = Fq[u]/(u^3-11)) living in a
// shared-memory "slab", operated on by a TEAM of DEG warps: warp c of the team owns coefficient c
// of every element, for 32 independent curve operations (one per lane).
//
// Replaces multiexp/arith.cu:370-462 (Fp2, Karatsuba) and :465-613 (Fp3, Karatsuba), which ran one
// element per 16-lane tile.  Here a tower product is DEG dot products of length DEG -- coefficient
// c is sum_{i+j = c mod DEG} a_i * b_j * (NR if i+j >= DEG) -- each with ONE lazy Montgomery
// reduction (fq_dot), the non-residue folded into the a-operand as a small unreduced multiple.
// Cost per tower product: DEG * (DEG+1) * 576 MACs, identical to the reference's Karatsuba with
// per-product reduction (3*1152, 6*1152) but perfectly balanced over the DEG warps and with no
// cross-coefficient additions.
//
// Slab layout: element e, coefficient c, quad q (limbs 4q..4q+3) of lane l is the uint4 at
//     slab[((e*DEG + c)*6 + q)*32 + l]
// so every warp-wide LDS.128/STS.128 touches 512 contiguous bytes (bank-conflict free).
//
// Synchronisation contract (DEG > 1): only mul/sqr/mul_by_a/is_zero read sibling coefficients.
// Each of them does  team_sync(); read+compute; team_sync(); store.  add/sub/neg/copy touch the
// caller's own coefficient only and need no barrier.  With this discipline every operation may be
// used in place.
#pragma once
#include "fq.cuh"
#ifndef MNT753_HOST_EMU
#include "fq_inv_coop.cuh"
#endif

namespace mnt753 {

#ifndef MNT753_QUADS_DEFINED
#define MNT753_QUADS_DEFINED
constexpr int QUADS = 6;    // uint4 per Fq element
constexpr int LANES = 32;
#endif

// Slot operations are real functions on the device (one copy each): the unrolled 24-limb bodies would
// otherwise be replicated at every call site and push the hot loops out of the instruction cache.
#ifdef MNT753_HOST_EMU
#define MSM_OP inline
#else
#define MSM_OP __device__ __noinline__
#endif

#ifdef MNT753_HOST_EMU
#define MSM_FOR_COMP(c) for (int c = 0; c < DEG; ++c)
#define MSM_NCOMP DEG
#define MSM_CI(c) (c)
#else
#define MSM_FOR_COMP(c) for (int c = comp, once_ = 1; once_; once_ = 0)
#define MSM_NCOMP 1
#define MSM_CI(c) 0
#endif

// cube roots of unity for the Frobenius map of Fq3 = Fq[u]/(u^3 - 11) (only MNT6753's base field has one)
template <class M> struct Frob3 {
    MSM_HD static constexpr uint32_t W1(int j) { return j == 0 ? 1u : 0u; }
    MSM_HD static constexpr uint32_t W2(int j) { return j == 0 ? 1u : 0u; }
};
template <> struct Frob3<ModB> {
    MSM_HD static constexpr uint32_t W1(int j) { constexpr uint32_t t[NLIMB] = MNT753_FROB3_W1_B_U32; return t[j]; }
    MSM_HD static constexpr uint32_t W2(int j) { constexpr uint32_t t[NLIMB] = MNT753_FROB3_W2_B_U32; return t[j]; }
};

// Field configuration: M = base modulus, DEG = tower degree, NR = non-residue,
// AKIND selects the curve-coefficient multiplication (mul_by_a):
//   0: G1, a = A0 (small integer)                          coeff_a * x
//   1: MNT4753 G2 twist, a = (A0, 0)                       (A0*c0, A0*c1)        mnt4753_g2.cpp:31-34
//   2: MNT6753 G2 twist, a = (0, 0, A0/NR.. )              (A1*c1, A1*c2, A0*c0) mnt6753_g2.cpp:38-41
template <class M_, int DEG_, unsigned NR_, int AKIND_, unsigned A0_, unsigned A1_>
struct FieldCfg {
    typedef M_ M;
    static constexpr int DEG = DEG_;
    static constexpr unsigned NR = NR_;
    static constexpr int AKIND = AKIND_;
    static constexpr unsigned A0 = A0_, A1 = A1_;
};

template <class F>
struct Team {
    typedef typename F::M M;
    static constexpr int DEG = F::DEG;

    uint4 *slab;      // team slab base + lane
    uint32_t *flags;  // DEG words of team-shared scratch (device, DEG > 1)
    int comp;         // coefficient owned by this warp (device)
    int bar_id;       // named barrier of this team (device, DEG > 1)

    MSM_DEVICE uint4 *elem(int e, int c) const { return slab + ((e * DEG + c) * QUADS) * LANES; }

    MSM_DEVICE void sync() const {
#ifndef MNT753_HOST_EMU
        if (DEG > 1) asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(DEG * 32) : "memory");
        else __syncwarp();
#endif
    }

    MSM_DEVICE void ld(fq_t &r, int e, int c) const {
        const uint4 *p = elem(e, c);
#pragma unroll
        for (int q = 0; q < QUADS; ++q) {
            uint4 v = p[q * LANES];
            r[4 * q] = v.x; r[4 * q + 1] = v.y; r[4 * q + 2] = v.z; r[4 * q + 3] = v.w;
        }
    }
    MSM_DEVICE void st(int e, int c, const fq_t &r, bool pred = true) const {
        uint4 *p = elem(e, c);
        if (pred) {
#pragma unroll
            for (int q = 0; q < QUADS; ++q) {
                uint4 v;
                v.x = r[4 * q]; v.y = r[4 * q + 1]; v.z = r[4 * q + 2]; v.w = r[4 * q + 3];
                p[q * LANES] = v;
            }
        }
    }

    // d = a * b.  Deliberately NOT inlined on the device: one copy of the ~1.3k..2.7k-instruction
    // unrolled dot product per field keeps the hot loop inside the instruction cache.
#ifndef MNT753_HOST_EMU
    __device__ __noinline__
#endif
    void mul(int d, int a, int b, bool pred = true) const {
        fq_t res[MSM_NCOMP];
        sync();
        MSM_FOR_COMP(c) {
            uint32_t aa[DEG][NLIMB];
            BQuads<DEG> src;
            src.stride = LANES;
#pragma unroll
            for (int i = 0; i < DEG; ++i) {
                ld(aa[i], a, i);
                if (i > c) fq_scale_unreduced(aa[i], aa[i], F::NR);
                int j = c - i;
                if (j < 0) j += DEG;
                src.p[i] = elem(b, j);
            }
            fq_dot<M, DEG>(res[MSM_CI(c)], aa, src);
        }
        sync();
        MSM_FOR_COMP(c) st(d, c, res[MSM_CI(c)], pred);
    }
    // d = (a1 - a2) * b, or a1 * b when a2 < 0: the difference is formed in registers on the way into the dot product
    // (batch_affine.cuh: three of the five products of an affine addition take a difference, and a slot operation of
    // its own costs a round trip through the slab).  One function for both forms keeps one copy of the product hot.
#ifndef MNT753_HOST_EMU
    __device__ __noinline__
#endif
    void mulsub(int d, int a1, int a2, int b, bool pred = true) const {
        fq_t res[MSM_NCOMP];
        sync();
        MSM_FOR_COMP(c) {
            uint32_t aa[DEG][NLIMB];
            BQuads<DEG> src;
            src.stride = LANES;
#pragma unroll
            for (int i = 0; i < DEG; ++i) {
                ld(aa[i], a1, i);
                if (a2 >= 0) { fq_t y; ld(y, a2, i); fq_sub<M>(aa[i], aa[i], y); }
                if (i > c) fq_scale_unreduced(aa[i], aa[i], F::NR);
                int j = c - i;
                if (j < 0) j += DEG;
                src.p[i] = elem(b, j);
            }
            fq_dot<M, DEG>(res[MSM_CI(c)], aa, src);
        }
        sync();
        MSM_FOR_COMP(c) st(d, c, res[MSM_CI(c)], pred);
    }
    // d = a^2.  Fq3 = Fq[u]/(u^3 - NR): the symmetric terms pair up, so every coefficient is a dot product of
    // length TWO (one lazy reduction) instead of three -- 1752 instead of 2328 MAC per warp:
    //     c0 = a0 a0 + (2 NR a1) a2      c1 = (2 a0) a1 + (NR a2) a2      c2 = (2 a0) a2 + a1 a1
    // (fp3.tcc:125-163 squares with the Chung-Hasan SQR2 formulas on reduced operands; same result.)
    // Fq and Fq2 square through mul: for Fq2 the two coefficients (a0^2 + NR a1^2, 2 a0 a1) would be unbalanced
    // over the team's two warps, so nothing is gained.
#ifndef MNT753_HOST_EMU
    __device__ __noinline__
#endif
    void sqr3(int d, int a, bool pred) const {
        fq_t res[MSM_NCOMP];
        sync();
        MSM_FOR_COMP(c) {
            uint32_t aa[2][NLIMB];
            BQuads<2> src;
            src.stride = LANES;
            const int r1 = (c == 1) ? 2 : 1;
            const uint32_t k0 = (c == 0) ? 1u : 2u, k1 = (c == 0) ? 2u * F::NR : (c == 1 ? (uint32_t)F::NR : 1u);
            ld(aa[0], a, 0);
            ld(aa[1], a, r1);
            fq_scale_unreduced(aa[0], aa[0], k0);
            fq_scale_unreduced(aa[1], aa[1], k1);
            src.p[0] = elem(a, c);
            src.p[1] = elem(a, c == 2 ? 1 : 2);
            fq_dot<M, 2>(res[MSM_CI(c)], aa, src);
        }
        sync();
        MSM_FOR_COMP(c) st(d, c, res[MSM_CI(c)], pred);
    }
    MSM_DEVICE void sqr(int d, int a, bool pred = true) const {
        if (DEG == 3) sqr3(d, a, pred);
        else mul(d, a, a, pred);
    }

    MSM_OP void add(int d, int a, int b, bool pred = true) const {
        MSM_FOR_COMP(c) { fq_t x, y; ld(x, a, c); ld(y, b, c); fq_add<M>(x, x, y); st(d, c, x, pred); }
    }
    MSM_OP void sub(int d, int a, int b, bool pred = true) const {
        MSM_FOR_COMP(c) { fq_t x, y; ld(x, a, c); ld(y, b, c); fq_sub<M>(x, x, y); st(d, c, x, pred); }
    }
    // d = a - b - c2
    MSM_OP void sub_sub(int d, int a, int b, int c2, bool pred = true) const {
        MSM_FOR_COMP(c) { fq_t x, y; ld(x, a, c); ld(y, b, c); fq_sub<M>(x, x, y); ld(y, c2, c); fq_sub<M>(x, x, y); st(d, c, x, pred); }
    }
    MSM_OP void dbl(int d, int a, bool pred = true) const {
        MSM_FOR_COMP(c) { fq_t x; ld(x, a, c); fq_add<M>(x, x, x); st(d, c, x, pred); }
    }
    // d = neg ? -a : a   (per lane)
    MSM_OP void neg_if(int d, int a, bool neg, bool pred = true) const {
        MSM_FOR_COMP(c) {
            fq_t x, y; ld(x, a, c); fq_neg<M>(y, x);
#pragma unroll
            for (int i = 0; i < NLIMB; ++i) x[i] = neg ? y[i] : x[i];
            st(d, c, x, pred);
        }
    }
    MSM_OP void copy(int d, int a, bool pred = true) const {
        MSM_FOR_COMP(c) { fq_t x; ld(x, a, c); st(d, c, x, pred); }
    }
    MSM_DEVICE void set_zero(int d, bool pred = true) const {
        MSM_FOR_COMP(c) { fq_t x;
#pragma unroll
            for (int i = 0; i < NLIMB; ++i) x[i] = 0;
            st(d, c, x, pred); }
    }
    MSM_DEVICE void set_one(int d, bool pred = true) const {
        MSM_FOR_COMP(c) { fq_t x;
#pragma unroll
            for (int i = 0; i < NLIMB; ++i) x[i] = (c == 0) ? M::R1(i) : 0u;
            st(d, c, x, pred); }
    }

    // d = coeff_a * x  (see FieldCfg::AKIND)
    MSM_OP void mul_by_a(int d, int x, bool pred = true) const {
        fq_t res[MSM_NCOMP];
        sync();
        MSM_FOR_COMP(c) {
            fq_t v;
            if (F::AKIND == 2) {
                ld(v, x, c == 2 ? 0 : c + 1);
                if (c == 2) fq_mul_small<M, F::A0>(res[MSM_CI(c)], v);
                else fq_mul_small<M, F::A1>(res[MSM_CI(c)], v);
            } else {
                ld(v, x, c);
                fq_mul_small<M, F::A0>(res[MSM_CI(c)], v);
            }
        }
        sync();
        MSM_FOR_COMP(c) st(d, c, res[MSM_CI(c)], pred);
    }


    // ---- cross-lane access and inversion (batched-affine accumulation, batch_affine.cuh) ----------
    MSM_DEVICE int lane() const {
#ifdef MNT753_HOST_EMU
        return 0;
#else
        return threadIdx.x & 31;
#endif
    }
    // d (this lane) = a of lane `src`.  Barriers on both sides: other lanes' data must be settled before
    // the read, and every read must be over before a later write to `a`.
    MSM_OP void copy_lane(int d, int a, int src) const {
        fq_t x[MSM_NCOMP];
        sync();
        MSM_FOR_COMP(c) {
            const uint4 *p = elem(a, c) - lane() + src;
#pragma unroll
            for (int q = 0; q < QUADS; ++q) {
                uint4 v = p[q * LANES];
                x[MSM_CI(c)][4 * q] = v.x; x[MSM_CI(c)][4 * q + 1] = v.y; x[MSM_CI(c)][4 * q + 2] = v.z; x[MSM_CI(c)][4 * q + 3] = v.w;
            }
        }
        sync();
        MSM_FOR_COMP(c) st(d, c, x[MSM_CI(c)]);
    }
    // conjugate-like maps whose product with a is the norm to Fq:  Fq2: (a0, -a1);  Fq3: Frobenius^k,
    // (a0, a1 w^k, a2 w^2k) with w = 11^((q-1)/3).  Own coefficient only.
    MSM_OP void frob(int d, int a, int k) const {
        MSM_FOR_COMP(c) {
            fq_t x, y;
            ld(x, a, c);
            if (DEG == 2) {
                if (c == 1) { fq_neg<M>(y, x); st(d, c, y); } else st(d, c, x);
            } else if (DEG == 3) {
                const int e = (c * k) % 3;
                if (e == 0) st(d, c, x);
                else {
                    fq_t w;
#pragma unroll
                    for (int i = 0; i < NLIMB; ++i) w[i] = (e == 1) ? Frob3<M>::W1(i) : Frob3<M>::W2(i);
                    fq_mul<M>(y, x, w);
                    st(d, c, y);
                }
            } else st(d, c, x);
        }
    }
    // d = k * a for k in Fq (Montgomery form, in registers): the caller's own coefficient only
    MSM_OP void scale_fq(int d, int a, const fq_t &k) const {
        MSM_FOR_COMP(c) { fq_t x, y; ld(x, a, c); fq_mul<M>(y, x, k); st(d, c, y); }
    }
    // d = 1 / a on EVERY lane (a != 0): the lanes run the binary gcd side by side (divergent, but the
    // alternative -- Fermat's a^(q-2), 750 squarings through the multiplier -- costs ten times more).  Used by
    // the affine normalisations (util_kernels.cuh), one inversion per lane and run of points.
    MSM_OP void inv_all(int d, int a, int t0, int t1) const {
        if (DEG == 1) {
            MSM_FOR_COMP(c) { fq_t x, r; ld(x, a, c); fq_inv<M>(r, x); st(d, c, r); }
            return;
        }
        frob(t0, a, 1);
        if (DEG == 3) { frob(t1, a, 2); mul(t0, t0, t1); }
        mul(t1, a, t0);                      // norm, lies in Fq: (N, 0[, 0])
#ifdef MNT753_HOST_EMU
        { fq_t x, r; ld(x, t1, 0); fq_inv<M>(r, x); st(t1, 0, r); }
#else
        if (comp == 0) { fq_t x, r; ld(x, t1, 0); fq_inv<M>(r, x); st(t1, 0, r); }
#endif
        mul(d, t0, t1);
    }
    // d = 1 / a for the element of LANE 0 only (other lanes: unspecified); a != 0.  t0, t1 scratch, all
    // four slots distinct.  One inversion in Fq; for the towers the element is first reduced to its norm:
    // a^-1 = conj(a) / N(a),  conj(a) = prod of the non-trivial Frobenius images.  On the device the Fq inversion is
    // done by the whole warp (fq_inv_coop.cuh: limb i of lane 0's element on lane i), with the one-thread binary
    // gcd of fq.cuh as the fallback; the host emulation runs the one-thread version.
    MSM_OP void inv_lane0(int d, int a, int t0, int t1) const {
        int src = a, dst = d;
        if (DEG > 1) {
            frob(t0, a, 1);
            if (DEG == 3) { frob(t1, a, 2); mul(t0, t0, t1); }
            mul(t1, a, t0);                  // norm, lies in Fq: (N, 0[, 0])
            src = dst = t1;
        }
#ifdef MNT753_HOST_EMU
        { fq_t x, r; ld(x, src, 0); fq_inv<M>(r, x); st(dst, 0, r); }
#else
        if (comp == 0) {
            const int l = lane();
            __syncwarp();
            // word k of quad q of lane 0's coefficient 0: 32-bit index (q * LANES) * 4 + k from the column of lane 0
            const int w = (l >> 2) * (LANES * 4) + (l & 3);
            uint32_t x = l < NLIMB ? reinterpret_cast<const uint32_t *>(elem(src, 0) - l)[w] : 0u;
            const bool ok = fq_inv_coop<M>(x);
            __syncwarp();
            if (ok) {
                if (l < NLIMB) reinterpret_cast<uint32_t *>(elem(dst, 0) - l)[w] = x;
            } else if (l == 0) {
                fq_t y, r; ld(y, src, 0); fq_inv<M>(r, y); st(dst, 0, r);
            }
            __syncwarp();
        }
#endif
        if (DEG > 1) mul(d, t0, t1);
    }

    // per-lane test, identical in every warp of the team
    MSM_OP bool is_zero(int e) const {
#ifdef MNT753_HOST_EMU
        bool z = true;
        for (int c = 0; c < DEG; ++c) { fq_t x; ld(x, e, c); z = z && fq_is_zero(x); }
        return z;
#else
        fq_t x;
        ld(x, e, comp);
        bool z = fq_is_zero(x);
        if (DEG == 1) return z;
        const unsigned lane = threadIdx.x & 31;
        unsigned nzmask = __ballot_sync(0xffffffffu, !z);
        sync();
        if (lane == 0) flags[comp] = nzmask;
        sync();
        unsigned all = 0;
#pragma unroll
        for (int c = 0; c < DEG; ++c) all |= flags[c];
        return !((all >> lane) & 1u);
#endif
    }
};

}  // namespace mnt753
