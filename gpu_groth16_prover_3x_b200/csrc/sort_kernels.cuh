// The arguments of an MSM and its first phase: signed-digit recoding of the scalars and the counting sort of the
// (point, digit) entries by bucket -- histogram, exclusive scan, scatter.  Integer work only; in a header of its own so
// that the CPU test-suite can run exactly these kernels on a simulated thread block (tests/host_emu/sort_emu.cpp).
#pragma once
#include "fq.cuh"
#include "recode.cuh"

namespace mnt753 {

#ifndef MNT753_QUADS_DEFINED
#define MNT753_QUADS_DEFINED
constexpr int QUADS = 6;    // uint4 per Fq element
constexpr int LANES = 32;
#endif

struct MsmArgs {
    // problem
    uint32_t n;        // number of points
    uint32_t i0, i1;   // k_count: range of points handled by this launch (scalar upload is chunked)
    int c;             // window width in bits
    int Wd;            // number of signed digits (windows) per scalar = ceil(754 / c)
    int W;             // number of bucket sets = ceil(Wd / NT): digit w lands in set w % W, using table w / W
    uint32_t tab_stride;  // points per precomputed table (table t holds 2^(c*W*t) * P_i), see BaseSet
    int glv;           // G2 with window tables: the scalar is split k = k0 + k1 lam (glv.cuh), two halves of Wh digits each;
    int Wh;            //   digit w of half h is digit h * Wh + w of the MSM, Wd = 2 Wh
    uint32_t NB;       // buckets per set = 2^(c-1)
    uint32_t K;        // W * NB
    // buffers
    const uint32_t *bases;      // affine AoS, 2*DEG*24 words per point; row t * tab_stride + i = 2^(c*W*t) * P_i
    const uint8_t *base_inf;    // 1 if base is infinity
    uint32_t *scalars;          // n * 24 words; Montgomery in, plain integer after k_from_mont
    uint32_t *count;            // K
    uint32_t *offs;             // K + 1
    uint32_t *cursor;           // K
    uint32_t *entries;          // n * Wd : table row | sign << 31
    uint32_t *winsum;           // W Jacobian points
    uint32_t *result;           // 1 Jacobian point
};

// every non-zero signed digit of scalar i: f(digit index, digit).  A split scalar (a.glv) is two half scalars of
// 12 limbs with their signs in bit 31 of the top limb.
template <class Fn>
__device__ __forceinline__ void scalar_digits(const MsmArgs &a, uint32_t i, Fn f) {
    uint32_t k[NLIMB];
    const uint4 *p = reinterpret_cast<const uint4 *>(a.scalars + (size_t)i * NLIMB);
    for (int q = 0; q < QUADS; ++q) { uint4 v = p[q]; k[4 * q] = v.x; k[4 * q + 1] = v.y; k[4 * q + 2] = v.z; k[4 * q + 3] = v.w; }
    if (!a.glv) { for_each_digit(k, NLIMB, a.c, a.Wd, f); return; }
    for (int h = 0; h < 2; ++h) {
        uint32_t *kh = k + 12 * h;
        const bool neg = (kh[11] >> 31) != 0u;
        kh[11] &= 0x7fffffffu;
        for_each_digit(kh, 12, a.c, a.Wh, [&](int w, int d) { f(h * a.Wh + w, neg ? -d : d); });
    }
}

static __global__ void k_count(MsmArgs a) {
    uint32_t i = a.i0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.i1 || a.base_inf[i]) return;
    scalar_digits(a, i, [&](int w, int d) {
        uint32_t b = (uint32_t)(d < 0 ? -d : d) - 1u;
        atomicAdd(&a.count[(uint32_t)(w % a.W) * a.NB + b], 1u);
    });
}

static __global__ void k_scatter(MsmArgs a) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n || a.base_inf[i]) return;
    scalar_digits(a, i, [&](int w, int d) {
        uint32_t b = (uint32_t)(d < 0 ? -d : d) - 1u;
        uint32_t pos = atomicAdd(&a.cursor[(uint32_t)(w % a.W) * a.NB + b], 1u);
        a.entries[pos] = ((uint32_t)(w / a.W) * a.tab_stride + i) | (d < 0 ? 0x80000000u : 0u);
    });
}

// ---- exclusive scan of count[0..K) -> offs[0..K], offs[K] = total -----------------------------
constexpr int SCAN_T = 256, SCAN_E = 4, SCAN_B = SCAN_T * SCAN_E;
static __global__ void __launch_bounds__(SCAN_T) k_scan_local(const uint32_t *in, uint32_t *out, uint32_t *bsum, uint32_t K) {
    __shared__ uint32_t sh[SCAN_T];
    uint32_t base = blockIdx.x * SCAN_B + threadIdx.x * SCAN_E;
    uint32_t v[SCAN_E], s = 0;
    for (int e = 0; e < SCAN_E; ++e) { v[e] = (base + e < K) ? in[base + e] : 0u; s += v[e]; }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int d = 1; d < SCAN_T; d <<= 1) {
        uint32_t t = (threadIdx.x >= (unsigned)d) ? sh[threadIdx.x - d] : 0u;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    uint32_t excl = sh[threadIdx.x] - s;
    for (int e = 0; e < SCAN_E; ++e) { if (base + e < K) out[base + e] = excl; excl += v[e]; }
    if (threadIdx.x == SCAN_T - 1) bsum[blockIdx.x] = sh[SCAN_T - 1];
}
static __global__ void __launch_bounds__(SCAN_T) k_scan_bsum(uint32_t *bsum, uint32_t nb, uint32_t *total) {
    __shared__ uint32_t sh[SCAN_T];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += SCAN_T) {
        uint32_t i = base + threadIdx.x;
        uint32_t s = (i < nb) ? bsum[i] : 0u;
        sh[threadIdx.x] = s;
        __syncthreads();
        for (int d = 1; d < SCAN_T; d <<= 1) {
            uint32_t t = (threadIdx.x >= (unsigned)d) ? sh[threadIdx.x - d] : 0u;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        if (i < nb) bsum[i] = carry + sh[threadIdx.x] - s;
        __syncthreads();
        if (threadIdx.x == 0) carry += sh[SCAN_T - 1];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}
static __global__ void __launch_bounds__(SCAN_T) k_scan_add(uint32_t *out, uint32_t *cursor, const uint32_t *bsum, uint32_t K) {
    uint32_t base = blockIdx.x * SCAN_B + threadIdx.x * SCAN_E;
    uint32_t add = bsum[blockIdx.x];
    for (int e = 0; e < SCAN_E; ++e)
        if (base + e < K) { uint32_t v = out[base + e] + add; out[base + e] = v; cursor[base + e] = v; }
}

}  // namespace mnt753
