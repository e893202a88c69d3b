// Pippenger MSM pipeline kernels (replaces multiexp/reduce.cu:11-152 of the reference: the
// precomputed-multiples Straus kernel + pairwise tree is gone, and so is the 31x table).
//
//   k_flag_inf        once per base set: mark bases encoded as infinity (y == 0)
//   k_from_mont       scalars: Montgomery -> integer (reference: reduce.cu:35-36)
//   k_count           signed-digit recode (recode.cuh), histogram of (window, bucket) keys      } sort_kernels.cuh
//   k_scan_*          exclusive prefix sum of the histogram                                      }
//   k_scatter         single-pass radix (counting) sort: point indices grouped by (window, bucket) }
//   k_batch_add       bucket accumulation (batch_affine.cuh, round planning in ba_plan.cuh): rounds of pairwise affine
//   k_ba_fixup        additions, one persistent launch, every team takes its share of the sorted list through all rounds
//   k_tree_round      bucket reduction  sum_b (b + 1) B_b  as a tree of pairwise affine additions (bucket_tree.cuh, index
//   k_tree_finish     logic in tree_plan.cuh): one launch per level, then the k + 1 short lists of every set
//   k_sum             plain segmented sums (the terms of a set -> its window sum; partial results of shards)
//   k_horner          window combine: result = sum_w 2^(c*w) * S_w
//
// Work decomposition: a TEAM = DEG warps handles 32 lanes (see fe.cuh); blocks hold TPB teams.
#pragma once
#include <cuda_runtime.h>

#include "curves.cuh"
#include "ba_plan.cuh"
#include "sort_kernels.cuh"

namespace mnt753 {

// MsmArgs and the recode / counting-sort kernels: sort_kernels.cuh

// ------------------------------------------------------------------------------------------
// kernel-side team construction
template <class G, int NSLOT, int TPB>
struct TeamSetup {
    typedef typename G::F F;
    static constexpr int DEG = F::DEG;
    static constexpr int THREADS = TPB * DEG * 32;
    static constexpr size_t SMEM = (size_t)TPB * NSLOT * DEG * QUADS * LANES * sizeof(uint4);
    __device__ static Team<F> make(uint4 *smem, uint32_t (*flags)[4], int &team) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        team = warp / DEG;
        Team<F> T;
        T.slab = smem + (size_t)team * (NSLOT * DEG * QUADS * LANES) + lane;
        T.flags = flags[team];
        T.comp = warp % DEG;
        T.bar_id = 1 + team;
        return T;
    }
};

// global <-> slab movement of one coefficient (96 B contiguous in global memory)
template <class F>
__device__ __forceinline__ void g2s(const Team<F> &T, int slot, const uint32_t *g, bool pred) {
    if (pred) {
        const uint4 *src = reinterpret_cast<const uint4 *>(g) + T.comp * QUADS;
        uint4 *dst = T.elem(slot, T.comp);
#pragma unroll
        for (int q = 0; q < QUADS; ++q) dst[q * LANES] = src[q];
    }
}
template <class F>
__device__ __forceinline__ void s2g(const Team<F> &T, uint32_t *g, int slot, bool pred) {
    if (pred) {
        uint4 *dst = reinterpret_cast<uint4 *>(g) + T.comp * QUADS;
        const uint4 *src = T.elem(slot, T.comp);
#pragma unroll
        for (int q = 0; q < QUADS; ++q) dst[q] = src[q * LANES];
    }
}
// The same two movements done by the WARP for its 32 lanes together: lane l moves quad (k * 32 + l) % 6 of the
// element of lane (k * 32 + l) / 6, k = 0..5, so that one LDG.128 / STG.128 covers the contiguous 96 bytes of five
// or six elements (8-11 cache lines) instead of one 16-byte piece of each of 32 elements (32 lines, and every
// 32-byte sector requested twice).  g is the lane's own element as for g2s / s2g.  Must be called by all 32 lanes.
#ifndef B200_COOP_MOVES
#define B200_COOP_MOVES 1
#endif
template <class F>
__device__ __forceinline__ void g2s_coop(const Team<F> &T, int slot, const uint32_t *g, bool pred) {
    // asynchronous (LDGSTS): issue any number of these, then g2s_coop_wait() once -- one round trip for all of them
    const int lane = threadIdx.x & 31;
    const unsigned long long mine = pred ? reinterpret_cast<unsigned long long>(g + T.comp * NLIMB) : 0ull;
    uint4 *col0 = T.elem(slot, T.comp) - lane;
#pragma unroll
    for (int k = 0; k < QUADS; ++k) {
        const int idx = k * 32 + lane, p = idx / QUADS, q = idx - p * QUADS;
        const unsigned long long gp = __shfl_sync(0xffffffffu, mine, p);
        if (gp != 0ull) {
            const unsigned saddr = (unsigned)__cvta_generic_to_shared(col0 + q * LANES + p);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(reinterpret_cast<const uint4 *>(gp) + q) : "memory");
        }
    }
}
__device__ __forceinline__ void g2s_coop_begin() { __syncwarp(); }   // the lanes' earlier slab accesses are over
__device__ __forceinline__ void g2s_coop_wait() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
}
template <class F>
__device__ __forceinline__ void s2g_coop(const Team<F> &T, uint32_t *g, int slot, bool pred) {
    const int lane = threadIdx.x & 31;
    const unsigned long long mine = pred ? reinterpret_cast<unsigned long long>(g + T.comp * NLIMB) : 0ull;
    const uint4 *col0 = T.elem(slot, T.comp) - lane;
    __syncwarp();
#pragma unroll
    for (int k = 0; k < QUADS; ++k) {
        const int idx = k * 32 + lane, p = idx / QUADS, q = idx - p * QUADS;
        const unsigned long long gp = __shfl_sync(0xffffffffu, mine, p);
        if (gp != 0ull) reinterpret_cast<uint4 *>(gp)[q] = col0[q * LANES + p];
    }
    __syncwarp();
}
// asynchronous 128-bit global -> shared copies (LDGSTS)
template <class F>
__device__ __forceinline__ void g2s_async(const Team<F> &T, int slot, const uint32_t *g, bool pred) {
    if (pred) {
        const uint4 *src = reinterpret_cast<const uint4 *>(g) + T.comp * QUADS;
        uint4 *dst = T.elem(slot, T.comp);
#pragma unroll
        for (int q = 0; q < QUADS; ++q) {
            unsigned saddr = (unsigned)__cvta_generic_to_shared(dst + q * LANES);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(saddr), "l"(src + q) : "memory");
        }
    }
}
__device__ __forceinline__ void async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ void async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <class F>
__device__ __forceinline__ void load_jac(const Team<F> &T, int X, int Y, int Z, const uint32_t *g, bool pred) {
    constexpr int EW = F::DEG * NLIMB;
    g2s(T, X, g, pred);
    g2s(T, Y, g + EW, pred);
    g2s(T, Z, g + 2 * EW, pred);
}
template <class F>
__device__ __forceinline__ void store_jac(const Team<F> &T, uint32_t *g, int X, int Y, int Z, bool pred) {
    constexpr int EW = F::DEG * NLIMB;
    s2g(T, g, X, pred);
    s2g(T, g + EW, Y, pred);
    s2g(T, g + 2 * EW, Z, pred);
}

// ------------------------------------------------------------------------------------------
template <int DEG>
__global__ void k_flag_inf(const uint32_t *bases, uint32_t n, uint8_t *flag) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 *y = reinterpret_cast<const uint4 *>(bases + (size_t)i * 2 * DEG * NLIMB + DEG * NLIMB);
    uint32_t o = 0;
    for (int q = 0; q < DEG * QUADS; ++q) { uint4 v = y[q]; o |= v.x | v.y | v.z | v.w; }
    flag[i] = (o == 0);
}

template <class Fr>
__global__ void __launch_bounds__(128) k_from_mont(uint32_t *scalars, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint4 *p = reinterpret_cast<uint4 *>(scalars + (size_t)i * NLIMB);
    fq_t x, r;
#pragma unroll
    for (int q = 0; q < QUADS; ++q) { uint4 v = p[q]; x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w; }
    fq_from_mont<Fr>(r, x);
#pragma unroll
    for (int q = 0; q < QUADS; ++q) { uint4 v; v.x = r[4 * q]; v.y = r[4 * q + 1]; v.z = r[4 * q + 2]; v.w = r[4 * q + 3]; p[q] = v; }
}

// ------------------------------------------------------------------------------------------
// generic tail kernels: 12 slots  TOT(0-2) RUN(3-5) Q(6-8) T(9-11)
template <class G>
struct TailCfg {
    static constexpr int DEG = G::F::DEG;
    static constexpr int NSLOT = 12;
    static constexpr int TPB = DEG == 1 ? 2 : 1;
    typedef TeamSetup<G, NSLOT, TPB> TS;
};

// out[w*nout + o] = sum_{j<32} in[w*nin + o*32 + j]   (nout = ceil(nin/32)): one TEAM per output, one lane
// per input, xor-butterfly of five full additions (every lane ends up with the sum; lane 0 stores it).
// A serial 32-term sum per lane took 32 dependent additions (1.2 ms at ~38 us each); this takes five.
template <class G>
__global__ void __launch_bounds__(TailCfg<G>::TS::THREADS) k_sum(const uint32_t *in, uint32_t *out, uint32_t W, uint32_t nin, uint32_t m) {
    typedef typename G::F F;
    typedef TailCfg<G> C;
    constexpr int JACW = 3 * F::DEG * NLIMB;
    extern __shared__ uint4 smem[];
    __shared__ uint32_t s_flags[C::TPB][4];
    int team;
    const Team<F> T = C::TS::make(smem, s_flags, team);
    const int lane = threadIdx.x & 31;
    const PtSlots s = {0, 1, 2, 6, 7, 8, 9, 10, 11};
    const uint32_t nout = (nin + 31) / 32;
    const uint32_t id = blockIdx.x * C::TPB + team;          // output index
    const bool tvalid = id < W * nout;
    const uint32_t w = tvalid ? id / nout : 0u, o = tvalid ? id % nout : 0u;
    const uint32_t idx = o * 32u + (uint32_t)lane;
    const bool act = tvalid && idx < nin;
    (void)m;
    load_jac(T, s.X1, s.Y1, s.Z1, in + ((size_t)w * nin + idx) * JACW, act);
    T.set_zero(s.X1, !act); T.set_zero(s.Y1, !act); T.set_zero(s.Z1, !act);
    for (int k = 1; k < 32; k <<= 1) {
        T.copy_lane(s.X2, s.X1, lane ^ k);
        T.copy_lane(s.Y2, s.Y1, lane ^ k);
        T.copy_lane(s.Z2, s.Z1, lane ^ k);
        Ec<F>::add(T, s, true);
    }
    T.sync();
    store_jac(T, out + (size_t)id * JACW, s.X1, s.Y1, s.Z1, tvalid && lane == 0);
}

// result = sum_w 2^(c*w) * winsum[w]   (one lane; the other 31 idle)
template <class G>
__global__ void __launch_bounds__(TailCfg<G>::TS::THREADS) k_horner(MsmArgs a) {
    typedef typename G::F F;
    typedef TailCfg<G> C;
    constexpr int JACW = 3 * F::DEG * NLIMB;
    extern __shared__ uint4 smem[];
    __shared__ uint32_t s_flags[C::TPB][4];
    int team;
    const Team<F> T = C::TS::make(smem, s_flags, team);
    if (team != 0) return;
    const int lane = threadIdx.x & 31;
    const PtSlots s = {0, 1, 2, 6, 7, 8, 9, 10, 11};
    const bool me = lane == 0;
    T.set_zero(s.Z1);
    for (int w = a.W - 1; w >= 0; --w) {
        if (w != a.W - 1)
            for (int i = 0; i < a.c; ++i) Ec<F>::dbl(T, s, true);
        load_jac(T, s.X2, s.Y2, s.Z2, a.winsum + (size_t)w * JACW, me);
        T.set_zero(s.Z2, !me);
        Ec<F>::add(T, s, me);
    }
    // infinity is reported as (1, 1, 0) like the reference's ec_jac::set_zero (curves.cu:104-114)
    const bool inf = T.is_zero(s.Z1);
    T.set_one(s.X1, inf);
    T.set_one(s.Y1, inf);
    store_jac(T, a.result, s.X1, s.Y1, s.Z1, me);
}

}  // namespace mnt753
