// Round planning of the batched-affine bucket accumulation (batch_affine.cuh): point references, the arguments of the
// accumulation kernels, the cut of the sorted list into shares and the plan of one round of one share.  Integer and
// index logic only -- no field arithmetic -- in a header of its own so that the CPU test-suite can run exactly this code
// on a simulated warp and model-check the recycling of scratch slots (tests/host_emu/plan_emu.cpp).
#pragma once
#include <cstdint>

namespace mnt753 {

// first index i in [0, n] with offs[i] > x, minus one  (offs non-decreasing, offs[0] = 0 <= x < offs[n])
__device__ __forceinline__ uint32_t bucket_of(const uint32_t *offs, uint32_t n, uint32_t x) {
    uint32_t lo = 0, hi = n;  // invariant: offs[lo] <= x < offs[hi]
    while (hi - lo > 1) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (offs[mid] <= x) lo = mid; else hi = mid;
    }
    return lo;
}

// ---- point references
constexpr uint32_t REF_INF = 0xffffffffu;      // infinity / empty bucket / idle pair
constexpr uint32_t REF_NEG = 0x80000000u;      // table reference: add the NEGATED point
constexpr uint32_t REF_SCRATCH = 0x40000000u;  // index into the scratch point array, else a table row
constexpr uint32_t REF_IDX = 0x3fffffffu;
constexpr int BA_CTL_WORDS = 40;               // [0] rounds (max over teams) [1] largest bucket [2] additions [4 + r] pairs of round r

struct BaArgs {
    uint32_t K;                 // buckets (all sets)
    uint32_t U;                 // teams that take a share of the list
    uint32_t T;                 // entries per share = ceil(E / U), set by the kernels from the list's real length
    const uint32_t *offs;       // K + 1 bucket offsets into the sorted list (offs[K] = E)
    uint32_t *refs[2];          // ping-pong reference lists, E entries each; refs[0] = the sorted entries
    const uint32_t *bases;      // window tables (affine AoS)
    uint32_t *scratch;          // sums: region A (even rounds) at [0, capA), region B (odd rounds) at [capA, capA + capB)
    uint32_t capA;
    uint4 *pairs;               // a team's additions of the current round: (source 0, source 1, scratch slot, list slot)
    uint8_t *codes;             // classification of each addition, parked between the two passes
    uint32_t *cntv;             // K + U: points left in each piece (a bucket, or the part of a split bucket in one share)
    uint32_t *bucket_ref;       // K: what is left of every bucket (preset to REF_INF by the host)
    uint32_t *bnd_ref;          // 2 U: pieces of split buckets, at most two per share (first / last bucket of the share)
    uint32_t *bnd_bucket;       //      their bucket ids (preset to REF_INF by the host)
    uint32_t *ctl;              // BA_CTL_WORDS statistics
    // k_ba_fixup only: the compacted pieces as a list of their own
    uint32_t *fx_refs[2];       // 2 U each
    uint32_t *fx_offs;          // 2 U + 1
    uint32_t *fx_cntv;          // 2 U + 1
    uint32_t *fx_bucket;        // 2 U
    uint32_t fx_scratch_base;   // first scratch slot of the fix-up's own A / B regions
};

// One share of a sorted list as seen by the team that owns it.
struct BaView {
    const uint32_t *offs;       // bucket offsets of the list
    uint32_t E0, E1;            // the share: list entries [E0, E1)
    uint32_t b0, npieces;       // its buckets b0 .. b0 + npieces - 1 (first and last possibly cut by E0 / E1)
    uint32_t id;                // added to the bucket id wherever pieces of one bucket in different shares must not collide
    uint32_t *refs[2];
    uint32_t *cntv;
    uint4 *pairs;               // the share's own region of the pair list
    uint8_t *codes;
    uint32_t a_base, b_base;    // first scratch slot of region A / B
};

__device__ __forceinline__ const uint32_t *ba_ref_ptr(const BaArgs &a, uint32_t ref, int AFFW) {
    return ((ref & REF_SCRATCH) ? a.scratch : a.bases) + (size_t)(ref & REF_IDX) * AFFW;
}

// Start of share t: t * T, moved back to the start of its bucket unless that bucket is a giant (more than half a
// share), which is cut right there.  Monotone in t; share t is [boundary(t), boundary(t + 1)).
__device__ __forceinline__ uint32_t ba_boundary(const BaArgs &a, uint32_t t, uint32_t E) {
    const unsigned long long e = (unsigned long long)t * a.T;
    if (t >= a.U || e >= E) return E;
    const uint32_t b = bucket_of(a.offs, a.K, (uint32_t)e);
    const uint32_t lo = a.offs[b], hi = a.offs[b + 1];
    return (hi - lo > a.T / 2u) ? (uint32_t)e : lo;
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// Plan of round r for one share (executed by ONE warp).  Pieces are taken 32 at a time, a lane per piece for its
// bookkeeping (count halved, odd one out carried over); the PAIRS of those 32 pieces are then dealt over the lanes --
// pair t of the group belongs to the piece whose running pair count first exceeds t, found by a five-step search over
// the lanes -- so that a piece with thousands of points (a small window, the 0 / 1 wires of a witness, the top
// window) is planned by the whole warp and not by one lane.  Every pair of consecutive references of a piece becomes
// an addition appended to the share's pair list; an infinite operand turns it into a copy of the other one, two into
// nothing.  Returns the number of pairs listed (idle ones included).
__device__ __forceinline__ uint32_t ba_plan(const BaView &v, uint32_t r, int lane, uint32_t &maxc) {
    const uint32_t *cur = v.refs[r & 1u];
    uint32_t *nxt = v.refs[(r + 1u) & 1u];
    uint32_t total = 0;
    for (uint32_t base = 0; base < v.npieces; base += 32u) {
        const uint32_t q = base + (uint32_t)lane;
        const bool valid = q < v.npieces;
        const uint32_t b = v.b0 + q;
        uint32_t ps = 0, c = 0;
        if (valid) {
            ps = max(v.offs[b], v.E0);
            c = (r == 0u) ? min(v.offs[b + 1], v.E1) - ps : v.cntv[b + v.id];
        }
        maxc = max(maxc, c);
        const uint32_t np = c >> 1;
        const uint32_t incl = warp_incl_scan(np, lane);
        const uint32_t excl = incl - np, T = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t out0 = (r & 1u) ? v.b_base + ((ps + b + v.id) >> 2) : v.a_base + (ps >> 1);
        if (c & 1u) nxt[ps + np] = cur[ps + c - 1u];
        if (valid) v.cntv[b + v.id] = (c + 1u) >> 1;
        for (uint32_t t0 = 0; t0 < T; t0 += 32u) {
            const uint32_t t = t0 + (uint32_t)lane;
            int l = 0;                                   // lanes whose running count is <= t = the pair's piece
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) {
                const uint32_t probe = __shfl_sync(0xffffffffu, incl, l + step - 1);
                if (probe <= t) l += step;
            }
            const int src = l < 32 ? l : 31;
            const uint32_t pps = __shfl_sync(0xffffffffu, ps, src), pex = __shfl_sync(0xffffffffu, excl, src);
            const uint32_t pout = __shfl_sync(0xffffffffu, out0, src);
            if (t < T) {
                const uint32_t j = t - pex;
                const uint32_t r1 = cur[pps + 2u * j], r2 = cur[pps + 2u * j + 1u];
                // An infinite operand: the other one is COPIED to the pair's output slot (d.y = REF_INF), not passed on
                // by reference -- a reference handed through would outlive the round its slot is reserved for (the slot
                // scheme recycles a piece's slots every second round) and be overwritten under the reader.
                uint4 d;
                if (r1 == REF_INF && r2 == REF_INF) {
                    nxt[pps + j] = REF_INF;
                    d = make_uint4(REF_INF, REF_INF, 0u, 0u);
                } else if (r1 == REF_INF) d = make_uint4(r2, REF_INF, pout + j, pps + j);
                else d = make_uint4(r1, r2, pout + j, pps + j);
                v.pairs[total + t] = d;
            }
        }
        total += T;
    }
    return total;
}

}  // namespace mnt753
