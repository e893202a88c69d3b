// Batched-affine bucket accumulation (the reference's counterpart is the Straus kernel + pairwise tree,
// multiexp/reduce.cu:11-127).
//
// The sorted list of (bucket, point) entries is reduced by ROUNDS of pairwise AFFINE additions:
// a bucket holding k points holds ceil(k/2) after a round, so every bucket is down to one point after
// ceil(log2 k) rounds and the total number of additions is the same sum_b (k_b - 1) as a serial chain.
// An affine addition costs one field inversion; Montgomery's simultaneous inversion shares ONE
// inversion among all the additions of a tile (32 lanes x B additions):
//
//   forward   d_i = x2 - x1 (or 2 y1 when doubling, 1 when there is nothing to add);  prefix products
//   tile      product over the 32 lanes (xor butterfly), one binary-gcd inversion by one thread,
//             each lane's own inverse = inverse of the tile product x product of the other lanes
//   backward  1/d_i = inv * prefix_i;  inv *= d_i;  lambda = (y2 - y1)/d_i;  x3, y3
//
// = 6 field multiplications per addition (1 forward, 5 backward) + 10 per lane and tile, against 11
// (8M + 3S) for the Jacobian mixed addition.  Prefix products are parked in the x-half of the output
// slot of the same addition, so the round needs no scratch of its own.
//
// TEAM-LOCAL ROUNDS (round 2 of the build; round 1 ran every round as five launches over the whole list):
// the sorted list is cut into U contiguous ranges of about E / U entries, one per team (a team = DEG warps
// = 32 lanes), at bucket boundaries -- only a bucket larger than half a team's share is split, its pieces
// being summed afterwards by k_ba_fixup.  Each team then takes ITS buckets through all their rounds inside
// ONE persistent launch (k_batch_add): per round it plans its pairs (lane per bucket: pair descriptors,
// odd-one-out carried over), runs forward / inversion / backward over them, and rewrites its part of a
// ping-pong list of point REFERENCES (table row | sign, or scratch slot, or infinity).  No grid-wide
// synchronisation, no per-round scan / plan launches, and a round's fixed cost (the tile inversion) is hidden
// behind the other teams of the SM instead of idling the GPU.  When all teams are done every bucket is one
// reference in bucket_ref[], which the bucket reduction reads.
#pragma once
#ifdef MNT753_HOST_EMU
#include "curves.cuh"
#else
#include "msm_kernels.cuh"
#include "ba_plan.cuh"
#endif

namespace mnt753 {

constexpr uint32_t BA_NONE = 0xffffffffu;
#ifndef B200_BA_BMAX
#define B200_BA_BMAX 2048
#endif
constexpr int BA_BMAX = B200_BA_BMAX;  // additions per lane and tile
enum : uint32_t { BA_NORMAL = 0, BA_DBL = 1, BA_CANCEL = 2, BA_COPY1 = 3, BA_IDLE = 4 };

// Six slab slots per team: the operands, the running inverse and one prefix/scratch.  The denominator
// overwrites X2 (x2 is recovered as x1 + d), lambda overwrites Y2.
struct BaSlots { int X1, Y1, X2, Y2, INV, PRE; };

// Phase 1 of one addition: slots X1, X2 hold the abscissae (X2 only where has2).  Leaves the denominator
// d in X2 and classifies the pair.  load_y(pred) must bring (signed) Y1, Y2 into their slots for the
// lanes with pred; it is called only when some lane has x1 == x2.  PRE is scratch.
template <class F, class LoadY>
MSM_DEVICE uint32_t pair_forward(const Team<F> &T, const BaSlots &s, bool valid, bool has2, LoadY load_y) {
    T.sub(s.X2, s.X2, s.X1);
    const bool dz = T.is_zero(s.X2);
    const bool need_y = has2 && dz;
    uint32_t code = !valid ? BA_IDLE : (!has2 ? BA_COPY1 : BA_NORMAL);
    if (team_any(need_y)) {
        load_y(need_y);
        T.sub(s.PRE, s.Y2, s.Y1);
        const bool eq = T.is_zero(s.PRE);
        const bool y0 = T.is_zero(s.Y1);   // a point of order two doubles to infinity
        const bool dbl = need_y && eq && !y0;
        if (need_y) code = dbl ? BA_DBL : BA_CANCEL;
        T.dbl(s.X2, s.Y1, dbl);
    }
    T.set_one(s.X2, code >= BA_CANCEL);
    return code;
}

// Phase 2: slots X1, Y1, X2, Y2 hold the (signed) operands, PRE the prefix product, INV the running
// inverse.  Leaves the sum in (X2, Y2) and advances INV.  Returns true when the result is infinity.
// Split in two so that a kernel can let the ordinates arrive (cp.async) while the head runs: the head
// touches X1, X2, PRE, INV -- and Y1 only on lanes that double.
template <class F>
MSM_DEVICE void pair_backward_head(const Team<F> &T, const BaSlots &s, uint32_t code) {
    const bool dbl = code == BA_DBL;
    T.sub(s.X2, s.X2, s.X1);             // d = x2 - x1
    if (team_any(dbl)) T.dbl(s.X2, s.Y1, dbl); //   or 2 y1
    T.set_one(s.X2, code >= BA_CANCEL);
    T.mul(s.PRE, s.INV, s.PRE);          // 1 / d
    T.mul(s.INV, s.INV, s.X2);           // inverse of the shorter prefix
}
template <class F>
MSM_DEVICE bool pair_backward_tail(const Team<F> &T, const BaSlots &s, uint32_t code) {
    const bool dbl = code == BA_DBL;
    T.sub(s.Y2, s.Y2, s.Y1);             // numerator y2 - y1
    if (team_any(dbl)) {                 //   or 3 x1^2 + a; d is dead on those lanes, X2 <- 0 so that x1 + x2 = 2 x1 below
        T.sqr(s.X2, s.X1, dbl);
        T.add(s.Y2, s.X2, s.X2, dbl);
        T.add(s.Y2, s.Y2, s.X2, dbl);
        T.set_one(s.X2, dbl);
        T.mul_by_a(s.X2, s.X2, dbl);
        T.add(s.Y2, s.Y2, s.X2, dbl);
        T.set_zero(s.X2, dbl);
    }
    T.mul(s.Y2, s.Y2, s.PRE);            // lambda
    T.sqr(s.PRE, s.Y2);
    T.add(s.X2, s.X2, s.X1);
    T.add(s.X2, s.X2, s.X1);             // x1 + x2 = 2 x1 + d
    T.sub(s.X2, s.PRE, s.X2);            // x3 = lambda^2 - x1 - x2
    T.sub(s.PRE, s.X1, s.X2);
    T.mul(s.PRE, s.Y2, s.PRE);
    T.sub(s.Y2, s.PRE, s.Y1);            // y3 = lambda (x1 - x3) - y1
    const bool copy1 = code == BA_COPY1, cancel = code == BA_CANCEL;
    if (team_any(copy1)) { T.copy(s.X2, s.X1, copy1); T.copy(s.Y2, s.Y1, copy1); }
    if (team_any(cancel)) { T.set_zero(s.X2, cancel); T.set_zero(s.Y2, cancel); }
    return cancel;
}
template <class F>
MSM_DEVICE bool pair_backward(const Team<F> &T, const BaSlots &s, uint32_t code) {
    pair_backward_head(T, s, code);
    return pair_backward_tail(T, s, code);
}

// ---- the UNCLASSIFIED passes of a tile (ba_tile): every pair is taken for a generic addition, x1 != x2.
// d = (a1 - a2) * b.  Fq: the difference is formed inside the product (Team::mulsub).  Towers: every warp of the
// team reads all DEG coefficients of the first factor, so the fused form would subtract DEG times over; there the
// difference goes through the slab once (it replaces a1, which is dead or wanted as the difference in every use).
template <class F>
MSM_DEVICE void ba_mul_diff(const Team<F> &T, int d, int a1, int a2, int b, bool pred) {
    if (F::DEG == 1) T.mulsub(d, a1, a2, b, pred);
    else { T.sub(a1, a1, a2); T.mul(d, a1, b, pred); }
}
template <class F>
MSM_DEVICE void ba_mul_plain(const Team<F> &T, int d, int a1, int b) {
    if (F::DEG == 1) T.mulsub(d, a1, -1, b); else T.mul(d, a1, b);
}
// forward step: X1, X2 hold the abscissae; the running product INV takes the denominator x2 - x1 (pred: a real addition)
template <class F>
MSM_DEVICE void pair_generic_forward(const Team<F> &T, const BaSlots &s, bool pred) {
    ba_mul_diff(T, s.INV, s.X2, s.X1, s.INV, pred);
}
// backward step: X1, Y1, X2, Y2 hold the (signed) operands, PRE the exclusive prefix product, INV the inverse of the
// inclusive one.  Leaves the sum in (X2, Y2) and the inverse of the shorter prefix in INV (valid: a real addition).
template <class F>
MSM_DEVICE void pair_generic_backward(const Team<F> &T, const BaSlots &s, bool valid) {
    ba_mul_plain(T, s.PRE, s.INV, s.PRE);               // 1 / d
    ba_mul_diff(T, s.INV, s.X2, s.X1, s.INV, valid);    // inverse of the shorter prefix (towers: X2 <- d)
    ba_mul_diff(T, s.Y2, s.Y2, s.Y1, s.PRE, true);      // lambda = (y2 - y1) / d
    if (F::DEG == 3) T.sqr(s.PRE, s.Y2); else ba_mul_plain(T, s.PRE, s.Y2, s.Y2);
    if (F::DEG == 1) T.sub_sub(s.X2, s.PRE, s.X1, s.X2);                       // x3 = lambda^2 - x1 - x2
    else { T.sub_sub(s.X2, s.PRE, s.X2, s.X1); T.sub(s.X2, s.X2, s.X1); }     //    = lambda^2 - d - 2 x1
    ba_mul_diff(T, s.PRE, s.X1, s.X2, s.Y2, true);      // lambda (x1 - x3)
    T.sub(s.Y2, s.PRE, s.Y1);                           // y3
}

// After the forward pass INV holds each lane's product of denominators.  Replace it by the lane's own
// inverse: xor-butterfly product over the 32 lanes, one inversion (lane 0), own = tile^-1 * others.
// Scratch: S, A, B, C (four distinct slots, all free between the passes).
template <class F>
MSM_DEVICE void tile_inverse(const Team<F> &T, int INV, int S, int A, int B, int C) {
#ifdef MNT753_HOST_EMU
    T.inv_lane0(A, INV, B, C);
    T.copy(INV, A);
#else
    const int lane = T.lane();
    for (int m = 1; m < 32; m <<= 1) {
        T.copy_lane(A, INV, lane ^ m);
        if (m == 1) T.copy(S, A); else T.mul(S, S, A);
        T.mul(INV, INV, A);
    }
    T.inv_lane0(A, INV, B, C);
    T.copy_lane(B, A, 0);
    T.mul(INV, B, S);
#endif
}

#ifndef MNT753_HOST_EMU
// point references, the arguments of the accumulation and the round planning: ba_plan.cuh

template <class G>
struct BaCfg {
    static constexpr int DEG = G::F::DEG;
    static constexpr int NSLOT = 6;
    // The twelve warps of an SM form ONE block of 384 threads (not three blocks of four warps): same occupancy,
    // same code, but the multiplier measures 7.5-7.7 G modmul/s with twelve warps in one block against 6.65 as
    // 3 x 128, 4 x 96, 6 x 64 or 12 x 32 threads (tools/mul_sched_probe.cu, profiles/r01_mul_sched_probe.txt).
    static constexpr int TPB = 12 / DEG;
    static constexpr int MINB = 1;
    typedef TeamSetup<G, NSLOT, TPB> TS;
};

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// one coordinate (96 * DEG bytes) of the point whose first word is g: the caller's coefficient only
template <class F>
__device__ __forceinline__ void prefetch_coord(const Team<F> &T, const uint32_t *g) {
    const char *p = reinterpret_cast<const char *>(g) + T.comp * (NLIMB * 4);
    prefetch_l2(p);
    prefetch_l2(p + NLIMB * 4 - 4);
}

#if B200_COOP_MOVES
#define BA_G2S g2s_coop
#define BA_S2G s2g_coop
#define BA_G2S_BEGIN() g2s_coop_begin()
#define BA_G2S_WAIT() g2s_coop_wait()
#else
#define BA_G2S g2s
#define BA_S2G s2g
#define BA_G2S_BEGIN()
#define BA_G2S_WAIT()
#endif

// Where a tile's additions come from: get(p) is addition p of the caller's list -- (source 0, source 1, scratch slot
// of the sum, index of its reference in the output list) -- or an idle descriptor (x == REF_INF) beyond the list;
// codes[p] parks the classification between the two passes.
struct ListPairs {          // a share's planned pair list (k_batch_add)
    const uint4 *pairs;
    uint8_t *codes;
    uint32_t P;
    __device__ __forceinline__ uint4 get(uint32_t p) const { return p < P ? pairs[p] : make_uint4(REF_INF, REF_INF, 0u, 0u); }
};

// The additions [p0, p0 + 32 B) of the list: forward pass, one inversion, backward pass.
template <class F, class Src>
__device__ __forceinline__ void ba_tile(const Team<F> &T, const BaArgs &a, const Src &src, uint32_t *nxt_refs, uint32_t p0, uint32_t B) {
    constexpr int DEG = F::DEG;
    constexpr int EW = DEG * NLIMB, AFFW = 2 * EW;
    const int lane = threadIdx.x & 31;
    const BaSlots s = {0, 1, 2, 3, 4, 5};
    const uint4 idle = make_uint4(REF_INF, REF_INF, 0u, 0u);
    // ---- forward, unclassified: every pair is taken for a generic addition (x1 != x2), the denominators go straight
    // into the running product.  A pair that is NOT generic -- equal abscissae (doubling, cancellation) or a copy --
    // shows at the end: a zero product, or a descriptor without a second operand; the tile then runs the classified
    // passes below instead.  With independent points that never happens, and the generic passes need no per-pair zero
    // test (two team barriers in the towers), no classification parked in memory and 4 instead of 14 slot operations.
    uint4 nxt, nxt2;
    bool generic = true;
    {
        T.set_one(s.INV);
        nxt = src.get(p0 + lane);
        nxt2 = B > 1u ? src.get(p0 + 32u + lane) : idle;
        for (uint32_t i = 0; i < B; ++i) {
            const uint32_t p = p0 + i * 32u + lane;
            const uint4 d = nxt;
            const bool valid = d.x != REF_INF, has2 = d.y != REF_INF;
            if (team_any(valid && !has2)) generic = false;
            const uint32_t *g1 = ba_ref_ptr(a, d.x, AFFW), *g2 = ba_ref_ptr(a, d.y, AFFW);
            nxt = nxt2;
            if (nxt.x != REF_INF) {
                prefetch_coord(T, ba_ref_ptr(a, nxt.x, AFFW));
                if (nxt.y != REF_INF) prefetch_coord(T, ba_ref_ptr(a, nxt.y, AFFW));
            }
            nxt2 = i + 2 < B ? src.get(p + 64u) : idle;
            BA_G2S_BEGIN();
            BA_G2S(T, s.X1, g1, valid);
            BA_G2S(T, s.X2, g2, has2);
            BA_G2S_WAIT();
            BA_S2G(T, a.scratch + (size_t)d.z * AFFW, s.INV, valid);                // exclusive prefix, parked in the output slot
            pair_generic_forward(T, s, valid && has2);
        }
        if (team_any(T.is_zero(s.INV))) generic = false;
    }
    if (generic) {
        tile_inverse(T, s.INV, s.X1, s.Y1, s.X2, s.Y2);
        const uint32_t pl = p0 + (B - 1u) * 32u + lane;
        nxt = src.get(pl);
        nxt2 = B > 1u ? src.get(pl - 32u) : idle;
        for (int i = (int)B - 1; i >= 0; --i) {
            const uint32_t p = p0 + (uint32_t)i * 32u + lane;
            const uint4 d = nxt;
            const bool valid = d.x != REF_INF;
            const uint32_t *g1 = ba_ref_ptr(a, d.x, AFFW), *g2 = ba_ref_ptr(a, d.y, AFFW);
            uint32_t *out = a.scratch + (size_t)d.z * AFFW;
            nxt = nxt2;
            if (nxt.x != REF_INF) {
                const uint32_t *n1 = ba_ref_ptr(a, nxt.x, AFFW), *n2 = ba_ref_ptr(a, nxt.y, AFFW);
                prefetch_coord(T, n1); prefetch_coord(T, n1 + EW);
                prefetch_coord(T, n2); prefetch_coord(T, n2 + EW);
                prefetch_coord(T, a.scratch + (size_t)nxt.z * AFFW);
            }
            nxt2 = i > 1 ? src.get(p - 64u) : idle;
            BA_G2S_BEGIN();
            BA_G2S(T, s.X1, g1, valid);
            BA_G2S(T, s.Y1, g1 + EW, valid);
            BA_G2S(T, s.X2, g2, valid);
            BA_G2S(T, s.Y2, g2 + EW, valid);
            BA_G2S(T, s.PRE, out, valid);
            BA_G2S_WAIT();
            const bool n1 = valid && (d.x & REF_NEG) != 0u, n2 = valid && (d.y & REF_NEG) != 0u;
            if (team_any(n1)) T.neg_if(s.Y1, s.Y1, n1, valid);
            if (team_any(n2)) T.neg_if(s.Y2, s.Y2, n2, valid);
            pair_generic_backward(T, s, valid);
            BA_S2G(T, out, s.X2, valid);
            BA_S2G(T, out + EW, s.Y2, valid);
            if (valid && T.comp == 0) nxt_refs[d.w] = REF_SCRATCH | d.z;
        }
        return;
    }
    // ---- forward, classified: denominators and prefix products
    T.set_one(s.INV);
    // descriptors are read two steps ahead and the coordinates of the next step are pulled into L2, so
    // that neither the list nor the gathers are waited for at DRAM latency
    nxt = src.get(p0 + lane);
    nxt2 = B > 1u ? src.get(p0 + 32u + lane) : idle;
    for (uint32_t i = 0; i < B; ++i) {
        const uint32_t p = p0 + i * 32u + lane;
        const uint4 d = nxt;
        const bool valid = d.x != REF_INF, has2 = d.y != REF_INF;      // !has2: copy of operand 1
        const uint32_t *g1 = ba_ref_ptr(a, d.x, AFFW), *g2 = ba_ref_ptr(a, d.y, AFFW);
        nxt = nxt2;
        if (nxt.x != REF_INF) {
            prefetch_coord(T, ba_ref_ptr(a, nxt.x, AFFW));
            if (nxt.y != REF_INF) prefetch_coord(T, ba_ref_ptr(a, nxt.y, AFFW));
        }
        nxt2 = i + 2 < B ? src.get(p + 64u) : idle;
        BA_G2S_BEGIN();
        BA_G2S(T, s.X1, g1, valid);
        BA_G2S(T, s.X2, g2, has2);
        BA_G2S_WAIT();
        const uint32_t code = pair_forward(T, s, valid, has2, [&](bool pred) {
            BA_G2S_BEGIN();
            BA_G2S(T, s.Y1, g1 + EW, pred);
            BA_G2S(T, s.Y2, g2 + EW, pred);
            BA_G2S_WAIT();
            T.neg_if(s.Y1, s.Y1, (d.x & REF_NEG) != 0u, pred);
            T.neg_if(s.Y2, s.Y2, (d.y & REF_NEG) != 0u, pred);
        });
        if (valid && T.comp == 0) src.codes[p] = (uint8_t)code;                    // parked until the backward pass
        BA_S2G(T, a.scratch + (size_t)d.z * AFFW, s.INV, valid);                    // exclusive prefix, parked in the output slot
        T.mul(s.INV, s.INV, s.X2);
    }
    // ---- one inversion for the whole tile
    tile_inverse(T, s.INV, s.X1, s.Y1, s.X2, s.Y2);
    // ---- backward: the additions
    {
        const uint32_t pl = p0 + (B - 1u) * 32u + lane;
        nxt = src.get(pl);
        nxt2 = B > 1u ? src.get(pl - 32u) : idle;
    }
    for (int i = (int)B - 1; i >= 0; --i) {
        const uint32_t p = p0 + (uint32_t)i * 32u + lane;
        const uint4 d = nxt;
        const bool valid = d.x != REF_INF, has2 = d.y != REF_INF;
        const uint32_t *g1 = ba_ref_ptr(a, d.x, AFFW), *g2 = ba_ref_ptr(a, d.y, AFFW);
        uint32_t *out = a.scratch + (size_t)d.z * AFFW;
        nxt = nxt2;
        if (nxt.x != REF_INF) {
            const uint32_t *n1 = ba_ref_ptr(a, nxt.x, AFFW), *n2 = ba_ref_ptr(a, nxt.y, AFFW);
            prefetch_coord(T, n1); prefetch_coord(T, n1 + EW);
            if (nxt.y != REF_INF) { prefetch_coord(T, n2); prefetch_coord(T, n2 + EW); }
            prefetch_coord(T, a.scratch + (size_t)nxt.z * AFFW);
        }
        nxt2 = i > 1 ? src.get(p - 64u) : idle;
        const uint32_t code = valid ? src.codes[p] : (uint32_t)BA_IDLE;
        BA_G2S_BEGIN();
        BA_G2S(T, s.X1, g1, valid);
        BA_G2S(T, s.Y1, g1 + EW, valid);
        BA_G2S(T, s.X2, g2, has2);
        BA_G2S(T, s.Y2, g2 + EW, has2);
        BA_G2S(T, s.PRE, out, valid);
        BA_G2S_WAIT();
        T.neg_if(s.Y1, s.Y1, (d.x & REF_NEG) != 0u, valid);
        T.neg_if(s.Y2, s.Y2, (d.y & REF_NEG) != 0u, has2);
        const bool res_inf = pair_backward(T, s, code);
        BA_S2G(T, out, s.X2, valid);
        BA_S2G(T, out + EW, s.Y2, valid);
        if (valid && T.comp == 0) nxt_refs[d.w] = res_inf ? REF_INF : (REF_SCRATCH | d.z);
    }
}

// All rounds of one share.  Returns the number of rounds executed; the survivors are in v.refs[rounds & 1].
template <class F>
__device__ __forceinline__ uint32_t ba_share(const Team<F> &T, const BaArgs &a, const BaView &v, uint32_t *s_pairs /* team-shared word */) {
    const int lane = threadIdx.x & 31;
    uint32_t r = 0, adds = 0;
    for (;; ++r) {
        __threadfence();      // the sums and references of the previous round, written by the team's other lanes and warps
        T.sync();
        if (T.comp == 0) {
            uint32_t maxc = 0;
            const uint32_t P = ba_plan(v, r, lane, maxc);
            if (lane == 0) *s_pairs = P;
            if (r == 0) {
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) maxc = max(maxc, __shfl_xor_sync(0xffffffffu, maxc, d));
                if (lane == 0) atomicMax(a.ctl + 1, maxc);
            }
            if (lane == 0 && P && r < (uint32_t)(BA_CTL_WORDS - 4)) atomicAdd(a.ctl + 4 + r, P);
        }
        __threadfence();
        T.sync();
        const uint32_t P = *s_pairs;
        if (P == 0u) break;
        adds += P;
        // tiles of at most BA_BMAX additions per lane, the last two of a long list evened out
        uint32_t p0 = 0;
        while (p0 < P) {
            const uint32_t left = (P - p0 + 31u) / 32u;
            uint32_t B = left;
            if (left > (uint32_t)BA_BMAX) B = left >= 2u * (uint32_t)BA_BMAX ? (uint32_t)BA_BMAX : (left + 1u) / 2u;
            ba_tile(T, a, ListPairs{v.pairs, v.codes, P}, v.refs[(r + 1u) & 1u], p0, B);
            p0 += 32u * B;
        }
    }
    if (T.comp == 0 && lane == 0) {
        atomicMax(a.ctl, r);
        if (adds) atomicAdd(a.ctl + 2, adds);
    }
    return r;
}

template <class G>
__global__ void __launch_bounds__(BaCfg<G>::TS::THREADS, BaCfg<G>::MINB) k_batch_add(BaArgs a) {
    typedef typename G::F F;
    typedef BaCfg<G> C;
    extern __shared__ uint4 smem[];
    __shared__ uint32_t s_flags[C::TPB][4];
    __shared__ uint32_t s_w[C::TPB][4];
    int team;
    const Team<F> T = C::TS::make(smem, s_flags, team);
    const int lane = threadIdx.x & 31;
    const uint32_t E = a.offs[a.K];
    a.T = max(1u, (E + a.U - 1u) / a.U);
    // shares are dealt round-robin over the blocks so that a short list still spreads over all SMs
    for (uint32_t t = (uint32_t)team * gridDim.x + blockIdx.x; t < a.U; t += gridDim.x * C::TPB) {
        T.sync();
        if (T.comp == 0 && lane == 0) {
            const uint32_t E0 = ba_boundary(a, t, E), E1 = ba_boundary(a, t + 1u, E);
            s_w[team][0] = E0;
            s_w[team][1] = E1;
            s_w[team][2] = E0 < E1 ? bucket_of(a.offs, a.K, E0) : 0u;
            s_w[team][3] = E0 < E1 ? bucket_of(a.offs, a.K, E1 - 1u) : 0u;
        }
        T.sync();
        BaView v;
        v.offs = a.offs;
        v.E0 = s_w[team][0];
        v.E1 = s_w[team][1];
        v.b0 = s_w[team][2];
        const uint32_t b1 = s_w[team][3];
        T.sync();
        if (v.E0 >= v.E1) continue;
        v.npieces = b1 - v.b0 + 1u;
        v.id = t;
        v.refs[0] = a.refs[0];
        v.refs[1] = a.refs[1];
        v.cntv = a.cntv;
        v.pairs = a.pairs + (v.E0 >> 1) + t;
        v.codes = a.codes + (v.E0 >> 1) + t;
        v.a_base = 0u;
        v.b_base = a.capA;
        const uint32_t rounds = ba_share(T, a, v, &s_w[team][0]);
        // what is left of every piece: whole buckets are final, cut ones go to the boundary list
        if (T.comp == 0) {
            const uint32_t *cur = a.refs[rounds & 1u];
            for (uint32_t q = (uint32_t)lane; q < v.npieces; q += 32u) {
                const uint32_t b = v.b0 + q;
                const uint32_t lo = a.offs[b], hi = a.offs[b + 1];
                const uint32_t ps = max(lo, v.E0), pe = min(hi, v.E1);
                const uint32_t c = rounds ? a.cntv[b + t] : pe - ps;
                const uint32_t ref = c ? cur[ps] : REF_INF;
                if (lo >= v.E0 && hi <= v.E1) a.bucket_ref[b] = ref;
                else {
                    const uint32_t slot = 2u * t + (lo < v.E0 ? 0u : 1u);
                    a.bnd_ref[slot] = ref;
                    a.bnd_bucket[slot] = b;
                }
            }
        }
    }
}

// Pieces of split buckets (k_batch_add's boundary list) -> one reference per split bucket.  One team: the pieces are
// compacted into a list of their own, runs of equal bucket id are its buckets, and the same rounds reduce it.
// With no giant bucket in the MSM the list is empty and the kernel returns at once.
template <class G>
__global__ void __launch_bounds__(BaCfg<G>::TS::THREADS, BaCfg<G>::MINB) k_ba_fixup(BaArgs a) {
    typedef typename G::F F;
    typedef BaCfg<G> C;
    extern __shared__ uint4 smem[];
    __shared__ uint32_t s_flags[C::TPB][4];
    __shared__ uint32_t s_w[4];
    int team;
    const Team<F> T = C::TS::make(smem, s_flags, team);
    if (team != 0) return;
    const int lane = threadIdx.x & 31;
    if (T.comp == 0) {
        // compaction of the occupied slots (slot order = list order, so pieces of one bucket are consecutive)
        uint32_t n = 0, nb = 0, prev = REF_INF;
        for (uint32_t base = 0; base < 2u * a.U; base += 32u) {
            const uint32_t i = base + (uint32_t)lane;
            const uint32_t bk = i < 2u * a.U ? a.bnd_bucket[i] : REF_INF;
            const bool occ = bk != REF_INF;
            const unsigned m = __ballot_sync(0xffffffffu, occ);
            const uint32_t pos = n + (uint32_t)__popc(m & ((1u << lane) - 1u));
            // bucket id of the previous occupied slot: the nearest lower occupied lane, else the carry
            uint32_t before = prev;
            {
                const unsigned lower = m & ((1u << lane) - 1u);
                const int src = lower ? 31 - __clz((int)lower) : 0;
                const uint32_t vsrc = __shfl_sync(0xffffffffu, bk, src);
                if (lower) before = vsrc;
            }
            const bool head = occ && bk != before;
            const unsigned hm = __ballot_sync(0xffffffffu, head);
            if (occ) a.fx_refs[0][pos] = a.bnd_ref[i];
            if (head) {
                const uint32_t hpos = nb + (uint32_t)__popc(hm & ((1u << lane) - 1u));
                a.fx_offs[hpos] = pos;
                a.fx_bucket[hpos] = bk;
            }
            if (m) prev = __shfl_sync(0xffffffffu, bk, 31 - __clz((int)m));
            n += (uint32_t)__popc(m);
            nb += (uint32_t)__popc(hm);
        }
        if (lane == 0) { a.fx_offs[nb] = n; s_w[0] = n; s_w[1] = nb; }
    }
    __threadfence_block();
    T.sync();
    const uint32_t n = s_w[0], nb = s_w[1];
    if (n == 0u) return;
    BaView v;
    v.offs = a.fx_offs;
    v.E0 = 0u;
    v.E1 = n;
    v.b0 = 0u;
    v.npieces = nb;
    v.id = 0u;
    v.refs[0] = a.fx_refs[0];
    v.refs[1] = a.fx_refs[1];
    v.cntv = a.fx_cntv;
    v.pairs = a.pairs;          // k_batch_add is over: its pair list is free
    v.codes = a.codes;
    v.a_base = a.fx_scratch_base;
    v.b_base = a.fx_scratch_base + a.U + 1u;
    T.sync();
    const uint32_t rounds = ba_share(T, a, v, &s_w[2]);
    if (T.comp == 0) {
        const uint32_t *cur = a.fx_refs[rounds & 1u];
        for (uint32_t q = (uint32_t)lane; q < nb; q += 32u) {
            const uint32_t ps = a.fx_offs[q];
            const uint32_t c = rounds ? a.fx_cntv[q] : a.fx_offs[q + 1] - ps;
            a.bucket_ref[a.fx_bucket[q]] = c ? cur[ps] : REF_INF;
        }
    }
}
#endif  // !MNT753_HOST_EMU

}  // namespace mnt753
