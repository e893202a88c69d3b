// Batched-affine bucket accumulation (replaces the per-lane Jacobian chains of k_accumulate and the
// edge fold; the reference's counterpart is the Straus kernel + pairwise tree, multiexp/reduce.cu:11-127).
//
// The sorted list of (bucket, point) entries is reduced by ROUNDS of pairwise AFFINE additions:
// a bucket holding k points holds ceil(k/2) after a round, so every bucket is down to one point after
// ceil(log2(max occupancy)) rounds and the total number of additions is the same sum_b (k_b - 1) as a
// serial chain.  An affine addition costs one field inversion; Montgomery's simultaneous inversion
// shares ONE inversion among all the additions of a tile (32 lanes x B additions):
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
// Per round r:   k_scan_halve* (next offsets = scan of ceil(k/2))  ->  k_plan (source rows of every
// output point; binary search of its bucket)  ->  k_batch_add (the additions, persistent tiles).
// Round 0 reads the window tables through the sorted entry list (row | sign << 31); later rounds read
// the previous round's output.  Rounds after the last useful one exit on a device-side flag, so the host
// enqueues a fixed worst-case number of rounds without ever synchronising.
#pragma once
#ifdef MNT753_HOST_EMU
#include "curves.cuh"
#else
#include "msm_kernels.cuh"
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
// ------------------------------------------------------------------------------------------------
struct BaArgs {
    uint32_t K;                 // buckets (all sets)
    uint32_t round;
    const uint32_t *off_cur;    // K + 1 offsets of the round's input list
    uint32_t *off_next;         // K + 1 offsets of its output list
    const uint32_t *entries;    // round 0: sorted (row | sign << 31)
    const uint32_t *bases;      // round 0: window tables
    const uint32_t *in_pts;     // round > 0: previous output (affine AoS)
    const uint8_t *in_inf;
    uint32_t *out_pts;
    uint8_t *out_inf;
    uint4 *pairs;               // the round's additions: (source 0, source 1, output index, -)
    uint32_t *npairs;           // [rounds]
    uint32_t *maxcnt;           // [rounds + 1] largest bucket occupancy entering each round
    uint32_t *tile_counter;     // [rounds]
    uint32_t *nrounds;          // rounds actually executed
    uint32_t *bsum;             // scan scratch
    uint32_t nscan;
};

__device__ __forceinline__ bool ba_round_active(const BaArgs &a) { return a.round == 0 || a.maxcnt[a.round] > 1u; }

// offsets of the next round: exclusive scan of ceil(k / 2); also records the largest k of this round.
// (Three launches, same structure as k_scan_local / k_scan_bsum / k_scan_add.)
static __global__ void __launch_bounds__(SCAN_T) k_ba_scan_local(BaArgs a) {
    __shared__ uint32_t sh[SCAN_T];
    __shared__ uint32_t smax;
    if (a.round > 0 && a.maxcnt[a.round - 1] <= 1u) return;   // previous round did not run: nothing left
    if (threadIdx.x == 0) smax = 0;
    uint32_t base = blockIdx.x * SCAN_B + threadIdx.x * SCAN_E;
    uint32_t v[SCAN_E], s = 0, mx = 0;
    for (int e = 0; e < SCAN_E; ++e) {
        uint32_t k = (base + e < a.K) ? a.off_cur[base + e + 1] - a.off_cur[base + e] : 0u;
        mx = max(mx, k);
        v[e] = (k + 1u) >> 1;
        s += v[e];
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    atomicMax(&smax, mx);
    for (int d = 1; d < SCAN_T; d <<= 1) {
        uint32_t t = (threadIdx.x >= (unsigned)d) ? sh[threadIdx.x - d] : 0u;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    uint32_t excl = sh[threadIdx.x] - s;
    for (int e = 0; e < SCAN_E; ++e) { if (base + e < a.K) a.off_next[base + e] = excl; excl += v[e]; }
    if (threadIdx.x == SCAN_T - 1) a.bsum[blockIdx.x] = sh[SCAN_T - 1];
    if (threadIdx.x == 0) atomicMax(&a.maxcnt[a.round], smax);
}
static __global__ void __launch_bounds__(SCAN_T) k_ba_scan_bsum(BaArgs a) {
    if (!ba_round_active(a)) return;
    __shared__ uint32_t sh[SCAN_T];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < a.nscan; base += SCAN_T) {
        uint32_t i = base + threadIdx.x;
        uint32_t s = (i < a.nscan) ? a.bsum[i] : 0u;
        sh[threadIdx.x] = s;
        __syncthreads();
        for (int d = 1; d < SCAN_T; d <<= 1) {
            uint32_t t = (threadIdx.x >= (unsigned)d) ? sh[threadIdx.x - d] : 0u;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < a.nscan) a.bsum[i] = carry + sh[threadIdx.x] - s;
        __syncthreads();
        if (threadIdx.x == 0) carry += sh[SCAN_T - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) a.off_next[a.K] = carry;
}
static __global__ void __launch_bounds__(SCAN_T) k_ba_scan_add(BaArgs a) {
    if (!ba_round_active(a)) return;
    uint32_t base = blockIdx.x * SCAN_B + threadIdx.x * SCAN_E;
    uint32_t add = a.bsum[blockIdx.x];
    for (int e = 0; e < SCAN_E; ++e)
        if (base + e < a.K) a.off_next[base + e] += add;
}

// Plan of a round.  For every output point j: its bucket b (binary search in off_next), local index l, and
// its inputs off_cur[b] + 2l (+ 1 when the bucket still has a partner for it).  Real additions are appended
// to the round's pair list (warp-aggregated atomic; order is irrelevant, each pair carries its output
// index); an input without a partner -- or whose partner is infinity -- is copied to its output slot right
// here, so that the arithmetic kernel sees additions only.
//   round 0 : sources are sorted entries (table row | sign << 31), never infinity (filtered by the sort)
//   later   : sources are indices into the previous round's output, with its infinity flags
template <class G, bool FIRST>
__global__ void __launch_bounds__(256) k_ba_plan(BaArgs a) {
    typedef typename G::F F;
    typedef typename F::M M;
    constexpr int DEG = F::DEG, EW = DEG * NLIMB, AFFW = 2 * EW;
    if (!ba_round_active(a)) return;
    const uint32_t E = a.off_next[a.K];
    const uint32_t *in = FIRST ? a.bases : a.in_pts;
    const int lane = threadIdx.x & 31;
    const uint32_t stride = gridDim.x * blockDim.x;
    const uint32_t iters = (E + stride - 1) / stride;
    for (uint32_t it = 0; it < iters; ++it) {
        const uint32_t j = it * stride + blockIdx.x * blockDim.x + threadIdx.x;
        const bool valid = j < E;
        uint32_t s0 = 0, s1 = BA_NONE;
        bool s0_inf = false;
        // the 32 outputs of a warp are consecutive: two full binary searches (its first and last output) bound
        // the few-step search of every lane in between
        const uint32_t jw0 = j - (uint32_t)lane;
        uint32_t bb = 0;
        if ((lane == 0 || lane == 31) && jw0 < E) bb = bucket_of(a.off_next, a.K, lane == 0 ? jw0 : min(jw0 + 31u, E - 1u));
        const uint32_t b_first = __shfl_sync(0xffffffffu, bb, 0), b_last = __shfl_sync(0xffffffffu, bb, 31);
        if (valid) {
            uint32_t blo = b_first, bhi = b_last + 1u;   // invariant: off_next[blo] <= j < off_next[bhi]
            while (bhi - blo > 1u) {
                const uint32_t mid = blo + ((bhi - blo) >> 1);
                if (a.off_next[mid] <= j) blo = mid; else bhi = mid;
            }
            const uint32_t b = blo;
            const uint32_t l = j - a.off_next[b];
            const uint32_t lo = a.off_cur[b], cnt = a.off_cur[b + 1] - lo;
            const uint32_t i0 = lo + 2u * l;
            const bool has2 = 2u * l + 1u < cnt;
            if (FIRST) {
                s0 = a.entries[i0];
                if (has2) s1 = a.entries[i0 + 1];
            } else {
                const bool inf0 = a.in_inf[i0] != 0;
                const bool inf1 = has2 ? a.in_inf[i0 + 1] != 0 : true;
                if (has2 && !inf0 && !inf1) { s0 = i0; s1 = i0 + 1; }
                else if (has2 && inf0 && !inf1) s0 = i0 + 1;
                else { s0 = i0; s0_inf = inf0; }
            }
        }
        const bool is_pair = valid && s1 != BA_NONE;
        const unsigned m = __ballot_sync(0xffffffffu, is_pair);
        uint32_t base = 0;
        if (lane == 0 && m) base = atomicAdd(a.npairs + a.round, (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (is_pair) a.pairs[base + __popc(m & ((1u << lane) - 1u))] = make_uint4(s0, s1, j, 0u);
        if (valid && !is_pair) {
            const uint4 *src = reinterpret_cast<const uint4 *>(in + (size_t)(s0 & 0x7fffffffu) * AFFW);
            uint4 *dst = reinterpret_cast<uint4 *>(a.out_pts + (size_t)j * AFFW);
#pragma unroll
            for (int q = 0; q < DEG * QUADS; ++q) dst[q] = src[q];
            const bool neg = FIRST && (s0 >> 31);
#pragma unroll
            for (int c = 0; c < DEG; ++c) {
                fq_t y, ny;
#pragma unroll
                for (int q = 0; q < QUADS; ++q) { uint4 v = src[(DEG + c) * QUADS + q]; y[4 * q] = v.x; y[4 * q + 1] = v.y; y[4 * q + 2] = v.z; y[4 * q + 3] = v.w; }
                if (neg) { fq_neg<M>(ny, y);
#pragma unroll
                    for (int i = 0; i < NLIMB; ++i) y[i] = ny[i]; }
#pragma unroll
                for (int q = 0; q < QUADS; ++q) { uint4 v; v.x = y[4 * q]; v.y = y[4 * q + 1]; v.z = y[4 * q + 2]; v.w = y[4 * q + 3]; dst[(DEG + c) * QUADS + q] = v; }
            }
            a.out_inf[j] = s0_inf ? 1 : 0;
        }
    }
}

template <class G>
struct BaCfg {
    static constexpr int DEG = G::F::DEG;
    // PREFETCH (experiment, off): three more slots double-buffer the gathers (cp.async straight into the slab
    // while the previous addition is being computed); 27 KB of slab per warp then allow 8 warps per SM instead
    // of 12.  Measured on B200 at 2^20: G1 56.3 ms against 54.3 ms without it (Fq2 150.9 vs 146.4) -- with the
    // sources already pulled into L2 one step ahead (prefetch_coord) the gathers are not what the warps wait
    // for, and the extra commit/wait traffic costs more than it hides.  8 warps per SM without PREFETCH
    // (-DB200_BA_MINB=2) is within noise of 12 for G1 / Fq2 and 14 % slower for Fq3.
#ifndef B200_BA_PREFETCH
#define B200_BA_PREFETCH 0
#endif
    static constexpr bool PREFETCH = B200_BA_PREFETCH && DEG < 3;
    static constexpr int NSLOT = PREFETCH ? 9 : 6;
    // ONEBLOCK (default): the twelve warps of an SM form ONE block of 384 threads instead of three blocks of four
    // warps (four of three for Fq3).  Same occupancy, same code -- but the multiplier measures 7.5-7.7 G modmul/s
    // with twelve warps in one block against 6.65 as 3 x 128, 4 x 96, 6 x 64 or 12 x 32 threads
    // (tools/mul_sched_probe.cu, profiles/r01_mul_sched_probe.txt), and the accumulation follows: 2^20 G1
    // 47.2 -> 43.9 ms, MNT6753 G2 2^18 83.2 -> 69.4 ms.
#ifndef B200_BA_ONEBLOCK
#define B200_BA_ONEBLOCK 1
#endif
#if B200_BA_ONEBLOCK
    static constexpr int TPB = 12 / DEG;
    static constexpr int MINB = 1;
#else
    static constexpr int TPB = DEG == 1 ? 4 : (DEG == 2 ? 2 : 1);
#ifdef B200_BA_MINB
    static constexpr int MINB = B200_BA_MINB;
#else
    static constexpr int MINB = PREFETCH ? 2 : (DEG == 1 ? 3 : (DEG == 2 ? 3 : 4));
#endif
#endif
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

template <class G, bool FIRST>
__global__ void __launch_bounds__(BaCfg<G>::TS::THREADS, BaCfg<G>::MINB) k_batch_add(BaArgs a) {
    typedef typename G::F F;
    typedef BaCfg<G> C;
    constexpr int DEG = F::DEG;
    constexpr int EW = DEG * NLIMB, AFFW = 2 * EW;
    if (!ba_round_active(a)) return;
    extern __shared__ uint4 smem[];
    __shared__ uint32_t s_flags[C::TPB][4];
    __shared__ uint32_t s_tile[C::TPB];
    int team;
    const Team<F> T = C::TS::make(smem, s_flags, team);
    const int lane = threadIdx.x & 31;
    BaSlots s = {0, 1, 2, 3, 4, 5};
    int NX1 = 6, NX2 = 7, NPRE = 8;   // spare slots of the double buffer (PREFETCH only)
    const uint32_t *in = FIRST ? a.bases : a.in_pts;

    if (blockIdx.x == 0 && threadIdx.x == 0) *a.nrounds = a.round + 1;
    const uint32_t E = a.npairs[a.round];
    // Tile = 32 lanes x B consecutive pairs of the list, claimed from an atomic cursor.  Small rounds: one
    // tile per team (a round then costs one inversion latency, not several).  Large rounds: full tiles of
    // BA_BMAX while more than one full tile per team is left, then the remainder in equal shares, so that
    // the teams finish together without paying for many small tiles (each tile costs one inversion).
    const uint32_t teams = gridDim.x * C::TPB;
    const uint32_t B0 = (E + teams * 32u - 1u) / (teams * 32u);
    const uint4 idle = make_uint4(0u, 0u, 0u, 0u);
    __shared__ uint32_t s_B[C::TPB];

    bool first_tile = true;
    for (;;) {
        T.sync();
        if (T.comp == 0 && lane == 0) {
            uint32_t Bt;
            if (B0 <= (uint32_t)BA_BMAX) Bt = B0 < 1u ? 1u : B0;
            else {
                const uint32_t cur = *reinterpret_cast<volatile uint32_t *>(a.tile_counter + a.round);
                const uint32_t rem = E > cur ? E - cur : 0u;
                Bt = (rem + teams * 32u - 1u) / (teams * 32u);
                Bt = Bt < 16u ? 16u : (Bt > (uint32_t)BA_BMAX ? (uint32_t)BA_BMAX : Bt);
#ifdef B200_BA_STAGGER
                // (experiment, off: measured 48.3 vs 47.3 ms) all teams start a round together; first tiles of
                // different lengths keep their forward / inversion / backward phases from lining up across the SM
                if (first_tile) Bt = max(16u, Bt * (((blockIdx.x + (uint32_t)team) & 3u) + 1u) / 4u);
#endif
            }
            s_tile[team] = atomicAdd(a.tile_counter + a.round, 32u * Bt);
            s_B[team] = Bt;
        }
        T.sync();
        const uint32_t base = s_tile[team];
        const uint32_t B = s_B[team];
        first_tile = false;
        if (base >= E) break;

        // ---- forward: denominators and prefix products
        T.set_one(s.INV);
        // descriptors are read two steps ahead and the coordinates of the next step are pulled into L2, so
        // that neither the list nor the gathers are waited for at DRAM latency
        uint4 nxt = (base + lane < E) ? a.pairs[base + lane] : idle;
        uint4 nxt2 = (B > 1u && base + 32u + lane < E) ? a.pairs[base + 32u + lane] : idle;
        for (uint32_t i = 0; i < B; ++i) {
            const uint32_t p = base + i * 32u + lane;
            const bool valid = p < E;
            const uint4 d = nxt;
            const uint32_t r1 = d.x & 0x7fffffffu, r2 = d.y & 0x7fffffffu, j = d.z;
            nxt = nxt2;
            if (i + 1 < B && p + 32u < E) {
                prefetch_coord(T, in + (size_t)(nxt.x & 0x7fffffffu) * AFFW);
                prefetch_coord(T, in + (size_t)(nxt.y & 0x7fffffffu) * AFFW);
            }
            nxt2 = (i + 2 < B && p + 64u < E) ? a.pairs[p + 64u] : idle;
            if (C::PREFETCH) {
                if (i == 0) {
                    g2s_async(T, s.X1, in + (size_t)r1 * AFFW, valid);
                    g2s_async(T, s.X2, in + (size_t)r2 * AFFW, valid);
                    async_commit();
                }
                const bool more = i + 1 < B;
                if (more) {      // next step's abscissae land in the spare slots while this step computes
                    const bool nv = p + 32u < E;
                    g2s_async(T, NX1, in + (size_t)(nxt.x & 0x7fffffffu) * AFFW, nv);
                    g2s_async(T, NX2, in + (size_t)(nxt.y & 0x7fffffffu) * AFFW, nv);
                    async_commit();
                    async_wait<1>();
                } else async_wait<0>();
            } else {
                BA_G2S_BEGIN();
                BA_G2S(T, s.X1, in + (size_t)r1 * AFFW, valid);
                BA_G2S(T, s.X2, in + (size_t)r2 * AFFW, valid);
                BA_G2S_WAIT();
            }
            const uint32_t code = pair_forward(T, s, valid, valid, [&](bool pred) {
                BA_G2S_BEGIN();
                BA_G2S(T, s.Y1, in + (size_t)r1 * AFFW + EW, pred);
                BA_G2S(T, s.Y2, in + (size_t)r2 * AFFW + EW, pred);
                BA_G2S_WAIT();
                if (FIRST) { T.neg_if(s.Y1, s.Y1, d.x >> 31, pred); T.neg_if(s.Y2, s.Y2, d.y >> 31, pred); }
            });
            if (valid && T.comp == 0) a.out_inf[j] = (uint8_t)code;   // parked until the backward pass
            BA_S2G(T, a.out_pts + (size_t)j * AFFW, s.INV, valid);       // exclusive prefix, parked in the output slot
            T.mul(s.INV, s.INV, s.X2);
            if (C::PREFETCH) { int t = s.X1; s.X1 = NX1; NX1 = t; t = s.X2; s.X2 = NX2; NX2 = t; }
        }
        // ---- one inversion for the whole tile
        tile_inverse(T, s.INV, s.X1, s.Y1, s.X2, s.Y2);
        // ---- backward: the additions
        {
            const uint32_t pl = base + (B - 1u) * 32u + lane;
            nxt = (pl < E) ? a.pairs[pl] : idle;
            nxt2 = (B > 1u && pl - 32u < E) ? a.pairs[pl - 32u] : idle;
        }
        for (int i = (int)B - 1; i >= 0; --i) {
            const uint32_t p = base + (uint32_t)i * 32u + lane;
            const bool valid = p < E;
            const uint4 d = nxt;
            const uint32_t r1 = d.x & 0x7fffffffu, r2 = d.y & 0x7fffffffu, j = d.z;
            nxt = nxt2;
            if (i > 0 && p - 32u < E) {
                const uint32_t *n1 = in + (size_t)(nxt.x & 0x7fffffffu) * AFFW, *n2 = in + (size_t)(nxt.y & 0x7fffffffu) * AFFW;
                prefetch_coord(T, n1); prefetch_coord(T, n1 + EW);
                prefetch_coord(T, n2); prefetch_coord(T, n2 + EW);
                prefetch_coord(T, a.out_pts + (size_t)nxt.z * AFFW);
                prefetch_l2(a.out_inf + nxt.z);
            }
            nxt2 = (i > 1 && p - 64u < E) ? a.pairs[p - 64u] : idle;
            const uint32_t code = valid ? a.out_inf[j] : (uint32_t)BA_IDLE;
            bool res_inf;
            if (C::PREFETCH) {
                if (i == (int)B - 1) {   // first step of the pass: nothing was prefetched for it
                    g2s_async(T, s.X1, in + (size_t)r1 * AFFW, valid);
                    g2s_async(T, s.X2, in + (size_t)r2 * AFFW, valid);
                    g2s_async(T, s.PRE, a.out_pts + (size_t)j * AFFW, valid);
                    async_commit();
                }
                g2s_async(T, s.Y1, in + (size_t)r1 * AFFW + EW, valid);      // needed after two products
                g2s_async(T, s.Y2, in + (size_t)r2 * AFFW + EW, valid);
                async_commit();
                const bool more = i > 0;
                if (more) {              // next step: abscissae and prefix into the spare slots
                    const bool nv = p - 32u < E;
                    g2s_async(T, NX1, in + (size_t)(nxt.x & 0x7fffffffu) * AFFW, nv);
                    g2s_async(T, NX2, in + (size_t)(nxt.y & 0x7fffffffu) * AFFW, nv);
                    g2s_async(T, NPRE, a.out_pts + (size_t)nxt.z * AFFW, nv);
                    async_commit();
                    async_wait<2>();
                } else async_wait<1>();
                if (team_any(code == BA_DBL)) { if (more) async_wait<1>(); else async_wait<0>(); }
                if (FIRST && team_any(code == BA_DBL)) T.neg_if(s.Y1, s.Y1, d.x >> 31, valid);
                pair_backward_head(T, s, code);
                if (more) async_wait<1>(); else async_wait<0>();
                if (FIRST) {
                    if (!team_any(code == BA_DBL)) T.neg_if(s.Y1, s.Y1, d.x >> 31, valid);
                    T.neg_if(s.Y2, s.Y2, d.y >> 31, valid);
                }
                res_inf = pair_backward_tail(T, s, code);
            } else {
                BA_G2S_BEGIN();
                BA_G2S(T, s.X1, in + (size_t)r1 * AFFW, valid);
                BA_G2S(T, s.Y1, in + (size_t)r1 * AFFW + EW, valid);
                BA_G2S(T, s.X2, in + (size_t)r2 * AFFW, valid);
                BA_G2S(T, s.Y2, in + (size_t)r2 * AFFW + EW, valid);
                BA_G2S(T, s.PRE, a.out_pts + (size_t)j * AFFW, valid);
                BA_G2S_WAIT();
                if (FIRST) { T.neg_if(s.Y1, s.Y1, d.x >> 31, valid); T.neg_if(s.Y2, s.Y2, d.y >> 31, valid); }
                res_inf = pair_backward(T, s, code);
            }
            BA_S2G(T, a.out_pts + (size_t)j * AFFW, s.X2, valid);
            BA_S2G(T, a.out_pts + (size_t)j * AFFW + EW, s.Y2, valid);
            if (valid && T.comp == 0) a.out_inf[j] = res_inf ? 1 : 0;
            if (C::PREFETCH) { int t = s.X1; s.X1 = NX1; NX1 = t; t = s.X2; s.X2 = NX2; NX2 = t; t = s.PRE; s.PRE = NPRE; NPRE = t; }
        }
    }
}
#endif  // !MNT753_HOST_EMU

}  // namespace mnt753
