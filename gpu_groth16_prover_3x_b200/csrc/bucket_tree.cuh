// Bucket reduction  S_w = sum_b (b + 1) * B[w][b]  of every bucket set w, as a TREE of independent additions
// (the running-sum reduction of libff's multi_exp_inner, multiexp.tcc:244-278, is a serial chain over the buckets;
// the reference's GPU kernel has no buckets at all, multiexp/reduce.cu:11-76).
//
//   T_0[i] = B[i],  T_l[i] = T_(l-1)[2i] + T_(l-1)[2i + 1]          block sums of 2^l buckets        (NB - 1 additions)
//   O_j    = sum of the ODD entries of T_j = sum of the buckets whose index has bit j set            (~NB additions)
//   S      = T_k[0] + sum_j 2^j O_j                                   NB = 2^k buckets per set
//
// Every level is a list of independent pairwise AFFINE additions, run by the batched-affine tile of
// batch_affine.cuh (6 field multiplications per addition against 27 per bucket for a Jacobian running sum), one
// launch per level while the lists are longer than a warp:
//
//   round r = 1 .. h:  T_(r-1) -> T_r  and one halving step of the r lists that sum O_0 .. O_(r-1)  (O_j starts at
//                  round j + 1 from the odd entries of T_j); (2 + r) NB / 2^(r+1) additions per set
//     k_tree_round   r <= hA: batched-affine additions -- a round costs a tile inversion (~0.17 ms whatever its size)
//                    and 6 multiplications per addition, so it pays while a round has >~ 150 k additions
//     k_tree_jac     r > hA: one Jacobian addition per lane (16 multiplications, ~0.07 ms per level), nodes kept as
//                    Jacobian points (reference kind REF_JAC)
//   k_tree_finish  the k + 1 lists that are left (<= 32 entries each: T_h, the O lists, the masked T_h for j >= h):
//                  one team per list -- butterfly of five Jacobian additions, then the list's j doublings
//   k_sum          the k + 1 terms of a set -> its window sum;  k_horner combines the sets (msm_kernels.cuh)
//
// A node of the tree is (a reference, a scratch slot) at the same index: set * NODES + offset, NODES = 2 NB.
// References are those of batch_affine.cuh; a sum with an empty operand hands the other reference on (no slot is
// ever recycled here, so a reference may outlive its round).
#pragma once
#include "batch_affine.cuh"
#include "tree_plan.cuh"

namespace mnt753 {

// TreeArgs, the node layout and the descriptor of an addition (TreePairs): tree_plan.cuh

template <class G>
__global__ void __launch_bounds__(BaCfg<G>::TS::THREADS, BaCfg<G>::MINB) k_tree_round(BaArgs a, TreeArgs t) {
    typedef typename G::F F;
    typedef BaCfg<G> C;
    extern __shared__ uint4 smem[];
    __shared__ uint32_t s_flags[C::TPB][4];
    int team;
    const Team<F> T = C::TS::make(smem, s_flags, team);
    const int lane = threadIdx.x & 31;
    // the round's additions are dealt in contiguous ranges, round-robin over the blocks
    const uint32_t teams = gridDim.x * C::TPB, me = (uint32_t)team * gridDim.x + blockIdx.x;
    const uint32_t chunk = ((t.P + teams - 1u) / teams + 31u) & ~31u;
    const unsigned long long lo64 = (unsigned long long)me * chunk;
    if (lo64 >= t.P) return;
    const uint32_t p_lo = (uint32_t)lo64, p_hi = (uint32_t)min((unsigned long long)t.P, lo64 + chunk);
    const TreePairs src{t, t.codes, p_hi};
    if (T.comp == 0)
        for (uint32_t p = p_lo + (uint32_t)lane; p < p_hi; p += 32u) src.passthrough(p);
    uint32_t p0 = p_lo;
    while (p0 < p_hi) {
        const uint32_t left = (p_hi - p0 + 31u) / 32u;
        uint32_t B = left;
        if (left > (uint32_t)BA_BMAX) B = left >= 2u * (uint32_t)BA_BMAX ? (uint32_t)BA_BMAX : (left + 1u) / 2u;
        ba_tile(T, a, src, t.R, p0, B);
        p0 += 32u * B;
    }
}

// operand of a Jacobian addition: an affine reference (Z = 1), a Jacobian node, or nothing (Z = 0)
template <class F>
__device__ __forceinline__ void tree_load(const Team<F> &T, const BaArgs &a, const TreeArgs &t, uint32_t set, uint32_t ref, int X, int Y, int Z) {
    constexpr int EW = F::DEG * NLIMB, AFFW = 2 * EW, JACW = 3 * EW;
    const bool none = ref == REF_INF, jac = !none && (ref & REF_JAC) == REF_JAC, aff = !none && !jac;
    const uint32_t *g = jac ? t.J + ((size_t)set * t.jnodes + ((ref & REF_IDX) - (size_t)set * t.nodes - t.jbase)) * JACW : ba_ref_ptr(a, ref, AFFW);
    g2s(T, X, g, !none);
    g2s(T, Y, g + EW, !none);
    g2s(T, Z, g + 2 * EW, jac);
    T.set_zero(X, none);
    T.set_zero(Y, none);
    T.neg_if(Y, Y, aff && (ref & REF_NEG) != 0u, aff);
    T.set_zero(Z, none);
    T.set_one(Z, aff);
}

// round r > hA: one Jacobian addition per lane
template <class G>
__global__ void __launch_bounds__(TailCfg<G>::TS::THREADS) k_tree_jac(BaArgs a, TreeArgs t) {
    typedef typename G::F F;
    typedef TailCfg<G> C;
    constexpr int JACW = 3 * F::DEG * NLIMB;
    extern __shared__ uint4 smem[];
    __shared__ uint32_t s_flags[C::TPB][4];
    int team;
    const Team<F> T = C::TS::make(smem, s_flags, team);
    const uint32_t lane = threadIdx.x & 31;
    const PtSlots s = {0, 1, 2, 6, 7, 8, 9, 10, 11};
    const uint32_t p = (blockIdx.x * C::TPB + team) * 32u + lane;
    const bool valid = p < t.P;
    const TreePairs src{t, t.codes, t.P};
    uint32_t r0 = REF_INF, r1 = REF_INF, node = 0;
    if (valid) {
        const uint32_t *i0, *i1;
        src.locate(p, i0, i1, node);
        r0 = *i0;
        r1 = *i1;
    }
    const uint32_t set = node / t.nodes;
    tree_load(T, a, t, set, r0, s.X1, s.Y1, s.Z1);
    tree_load(T, a, t, set, r1, s.X2, s.Y2, s.Z2);
    T.sync();
    Ec<F>::add(T, s, valid);
    T.sync();
    store_jac(T, t.J + ((size_t)set * t.jnodes + (node - (size_t)set * t.nodes - t.jbase)) * JACW, s.X1, s.Y1, s.Z1, valid);
    if (valid && T.comp == 0) t.R[node] = REF_JAC | node;
}

// fin[set * (k + 1) + l]: l = 0 the sum of all buckets, l = 1 + j the term 2^j O_j  (Jacobian)
template <class G>
__global__ void __launch_bounds__(TailCfg<G>::TS::THREADS) k_tree_finish(BaArgs a, TreeArgs t, uint32_t *fin) {
    typedef typename G::F F;
    typedef TailCfg<G> C;
    constexpr int JACW = 3 * F::DEG * NLIMB;
    extern __shared__ uint4 smem[];
    __shared__ uint32_t s_flags[C::TPB][4];
    int team;
    const Team<F> T = C::TS::make(smem, s_flags, team);
    const uint32_t lane = threadIdx.x & 31;
    const PtSlots s = {0, 1, 2, 6, 7, 8, 9, 10, 11};
    const uint32_t id = blockIdx.x * C::TPB + team;
    const bool tvalid = id < t.W * (t.k + 1u);
    const uint32_t set = tvalid ? id / (t.k + 1u) : 0u, l = tvalid ? id % (t.k + 1u) : 0u;
    uint32_t ref = REF_INF;
    if (tvalid) {
        const uint32_t nt = t.NB >> t.h;                    // entries of T_h (<= 32)
        if (l == 0u) { if (lane < nt) ref = tree_tlist(t, set, t.h)[lane]; }
        else if (l - 1u < t.h) { if (lane < (nt >> 1)) ref = t.R[(size_t)set * t.nodes + tree_ooff(t, l - 1u, t.h - (l - 1u)) + lane]; }
        else if (lane < nt && ((lane >> (l - 1u - t.h)) & 1u)) ref = tree_tlist(t, set, t.h)[lane];
    }
    tree_load(T, a, t, set, ref, s.X1, s.Y1, s.Z1);
    T.sync();
    for (uint32_t m = 1; m < 32u; m <<= 1) {
        T.copy_lane(s.X2, s.X1, (int)(lane ^ m));
        T.copy_lane(s.Y2, s.Y1, (int)(lane ^ m));
        T.copy_lane(s.Z2, s.Z1, (int)(lane ^ m));
        Ec<F>::add(T, s, true);
    }
    for (uint32_t i = 1; i < l; ++i) Ec<F>::dbl(T, s, true);     // l = 1 + j: j doublings
    T.sync();
    store_jac(T, fin + (size_t)id * JACW, s.X1, s.Y1, s.Z1, tvalid && lane == 0u);
}

}  // namespace mnt753
