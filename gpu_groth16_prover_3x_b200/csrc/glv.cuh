// The G2 endomorphism the reference's README points at (README.md:73, eprint 2008/117; CPU mirror: libff
// G2::mul_by_q, mnt4753_g2.cpp:364-368, mnt6753_g2.cpp:370-375): on G2 the twist Frobenius
//     psi(x, y) = (cX Frob(x), cY Frob(y))
// is multiplication by lam = q mod r, a 377-bit number.  Every G2 scalar is split as k = k0 + k1 lam (mod r) with
// |k0|, |k1| < 2^379, so that an MSM over n points with 753-bit scalars becomes one over the 2 n points
// {P_i, psi(P_i)} with 379-bit scalars.  With window tables that leaves the number of additions as it was, but the
// upper half of the tables -- 2^(c t) psi(P_i) = psi(2^(c t) P_i) -- is derived from the lower half by two
// multiplications with Fq constants per point instead of c doublings and an affine normalisation: half the table
// build time of every G2 base set (3.2 instead of 6.3 s for 2^20 MNT6753 G2 points).
//
//   k_glv_split   k -> (|k0|, sign) || (|k1|, sign) in place, 12 limbs each (sign in bit 31 of the half's top limb)
//   k_psi_many    out[i] = psi(in[i]) for affine points (infinity = all zero stays all zero)
// Constants: mnt753_constants.h, derived and checked by tools/gen_constants.py (glv_params).
#pragma once
#include "msm_kernels.cuh"
#include "glv_split.cuh"

namespace mnt753 {

// one thread per scalar (plain integer, already out of Montgomery form)
template <int CURVE>
__global__ void __launch_bounds__(128) k_glv_split(uint32_t *scalars, uint32_t n) {
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n) return;
    uint4 *p = reinterpret_cast<uint4 *>(scalars + (size_t)idx * NLIMB);
    uint32_t k[NLIMB];
#pragma unroll
    for (int q = 0; q < QUADS; ++q) { const uint4 v = p[q]; k[4 * q] = v.x; k[4 * q + 1] = v.y; k[4 * q + 2] = v.z; k[4 * q + 3] = v.w; }
    uint32_t r[2][GLV_L];
    glv_split<CURVE>(k, r);
#pragma unroll
    for (int q = 0; q < QUADS; ++q) {
        const int h = q / 3, o = (q % 3) * 4;
        p[q] = make_uint4(r[h][o], r[h][o + 1], r[h][o + 2], r[h][o + 3]);
    }
}

// out[i] = psi(in[i]), one lane per point
template <class G>
__global__ void __launch_bounds__(TailCfg<G>::TS::THREADS) k_psi_many(uint32_t n, const uint32_t *in, uint32_t *out) {
    typedef typename G::F F;
    typedef TailCfg<G> C;
    typedef Glv<G::CURVE> K;
    constexpr int EW = F::DEG * NLIMB;
    extern __shared__ uint4 smem[];
    __shared__ uint32_t s_flags[C::TPB][4];
    int team;
    const Team<F> T = C::TS::make(smem, s_flags, team);
    const uint32_t id = (blockIdx.x * C::TPB + team) * 32 + (threadIdx.x & 31);
    const bool act = id < n;
    g2s(T, 0, in + (size_t)id * 2 * EW, act);
    g2s(T, 1, in + (size_t)id * 2 * EW + EW, act);
    T.set_zero(0, !act);
    T.set_zero(1, !act);
    T.sync();
    T.frob(2, 0, 1);
    T.frob(3, 1, 1);
    fq_t cx, cy;
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) { cx[i] = K::TWX(i); cy[i] = K::TWY(i); }
    T.scale_fq(2, 2, cx);
    T.scale_fq(3, 3, cy);
    T.sync();
    s2g(T, out + (size_t)id * 2 * EW, 2, act);
    s2g(T, out + (size_t)id * 2 * EW + EW, 3, act);
}

}  // namespace mnt753
