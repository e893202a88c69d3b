// Index logic of the bucket-reduction tree (bucket_tree.cuh): the node layout, and for addition p of round r its two
// inputs and its node.  No field arithmetic -- in a header of its own so that the CPU test-suite can run exactly this
// code with integers for points and check  S = sum_b (b + 1) B_b  (tests/host_emu/plan_emu.cpp).
#pragma once
#include "ba_plan.cuh"

namespace mnt753 {

struct TreeArgs {
    uint32_t W, NB, k, h, hA;    // sets, buckets per set = 2^k, rounds h = max(0, k - 5), the first hA of them affine
    uint32_t r, q, logq;         // this round; q = NB >> (r + 1) additions per O list (T has 2 q)
    uint32_t nodes;              // 2 NB per set
    uint32_t P;                  // additions of this round, all sets
    uint32_t slot_base;          // scratch slot of node 0
    const uint32_t *bucket_ref;  // level 0: W * NB references left by the accumulation
    uint32_t *R;                 // W * nodes references
    uint8_t *codes;              // P bytes
    uint32_t *J;                 // Jacobian nodes (rounds > hA): node n of set w at (w * jnodes + n - jbase)
    uint32_t jbase, jnodes;      // first node that can be Jacobian (T_(hA+1)); nodes - jbase
};

constexpr uint32_t REF_JAC = REF_NEG | REF_SCRATCH;   // a Jacobian node of the reduction tree (index = node; never all ones)

__device__ __forceinline__ uint32_t tree_toff(const TreeArgs &t, uint32_t l) { return t.NB - (t.NB >> (l - 1u)); }                 // T_l, l >= 1
__device__ __forceinline__ uint32_t tree_ooff(const TreeArgs &t, uint32_t j, uint32_t s) {                                           // O_j after s >= 1 steps
    return t.NB + (t.NB - (t.NB >> j)) + ((t.NB >> (j + 1u)) - (t.NB >> (j + s)));
}
__device__ __forceinline__ const uint32_t *tree_tlist(const TreeArgs &t, uint32_t set, uint32_t l) {
    return l == 0u ? t.bucket_ref + (size_t)set * t.NB : t.R + (size_t)set * t.nodes + tree_toff(t, l);
}

struct TreePairs {
    TreeArgs t;
    uint8_t *codes;
    uint32_t P;      // end of the caller's range of the round's additions
    // addition p of round t.r: its two input references and its node
    __device__ __forceinline__ void locate(uint32_t p, const uint32_t *&i0, const uint32_t *&i1, uint32_t &node) const {
        const uint32_t per = t.q * (2u + t.r);
        const uint32_t set = p / per, rem = p - set * per, unit = rem >> t.logq;
        if (unit < 2u) {
            const uint32_t *in = tree_tlist(t, set, t.r - 1u) + 2u * rem;
            i0 = in; i1 = in + 1;
            node = set * t.nodes + tree_toff(t, t.r) + rem;
        } else {
            const uint32_t j = unit - 2u, s = t.r - j, i = rem & (t.q - 1u);
            if (s == 1u) {
                const uint32_t *in = tree_tlist(t, set, t.r - 1u) + 4u * i + 1u;
                i0 = in; i1 = in + 2;
            } else {
                const uint32_t *in = t.R + (size_t)set * t.nodes + tree_ooff(t, j, s - 1u) + 2u * i;
                i0 = in; i1 = in + 1;
            }
            node = set * t.nodes + tree_ooff(t, j, s) + i;
        }
    }
    __device__ __forceinline__ uint4 get(uint32_t p) const {
        if (p >= P) return make_uint4(REF_INF, REF_INF, 0u, 0u);
        const uint32_t *i0, *i1;
        uint32_t node;
        locate(p, i0, i1, node);
        const uint32_t r0 = *i0, r1 = *i1;
        if (r0 == REF_INF || r1 == REF_INF) return make_uint4(REF_INF, REF_INF, 0u, 0u);
        return make_uint4(r0, r1, t.slot_base + node, node);
    }
    // an empty operand: the node is the other operand (or empty)
    __device__ __forceinline__ void passthrough(uint32_t p) const {
        const uint32_t *i0, *i1;
        uint32_t node;
        locate(p, i0, i1, node);
        const uint32_t r0 = *i0, r1 = *i1;
        if (r0 == REF_INF) t.R[node] = r1;
        else if (r1 == REF_INF) t.R[node] = r0;
    }
};

}  // namespace mnt753
