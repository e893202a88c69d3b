// Device-side utilities around the MSM: affine normalisation (one binary-gcd inversion per lane, shared by
// a whole batch through Montgomery's trick), scalar multiplication of a single point, and the
// synthetic base generator  P_i = P0 + i*Q  of SURVEY.md 8(d) (the structured bases whose MSM has a
// closed form), produced directly in HBM so that benchmark inputs never cross PCIe.
//
// The reference has no field inversion on the device (multiexp/arith.cu:347-354 is #if 0) and
// normalises on the host with libff (to_affine_coordinates, mnt4753_g1.cpp:68-83); these kernels
// are the device counterpart used by b200msm_to_affine / b200msm_bases_synthetic.
#pragma once
#include "msm_kernels.cuh"

namespace mnt753 {

struct UtilSlots {
    static constexpr int X1 = 0, Y1 = 1, Z1 = 2, X2 = 3, Y2 = 4, Z2 = 5, T0 = 6, T1 = 7, T2 = 8, PRE = 9, INV = 10, TMP = 11;
};

// acc (slots 0..2) = k * (X2, Y2) for a per-lane 32-bit k; `top` = highest bit index to visit (uniform).
template <class F>
__device__ void small_scalar_mul(const Team<F> &T, const PtSlots &s, uint32_t k, int top, bool &acc_inf) {
    T.set_zero(s.X1); T.set_zero(s.Y1); T.set_zero(s.Z1);
    acc_inf = true;
    for (int bit = top; bit >= 0; --bit) {
        if (bit != top) Ec<F>::dbl(T, s, true);
        Ec<F>::madd(T, s, false, (k >> bit) & 1u, acc_inf);
    }
}

// Second half of Montgomery's simultaneous inversion for one lane's run [start, start + B) of Jacobian
// points: on entry slot PRE holds the product of the run's (non-zero-substituted) Z coordinates and
// prefix[idx] the inclusive prefix products; writes affine(jac[idx]) to out (infinity -> all zero).
template <class F>
__device__ void normalise_run(const Team<F> &T, uint32_t start, uint32_t B, uint32_t n, bool valid, const uint32_t *jac,
                              const uint32_t *prefix, uint32_t *out) {
    typedef UtilSlots U;
    constexpr int EW = F::DEG * NLIMB, AFFW = 2 * EW, JACW = 3 * EW;
    T.inv_all(U::INV, U::PRE, U::T0, U::T1);
    for (int j = (int)B - 1; j >= 0; --j) {
        const uint32_t idx = start + (uint32_t)j;
        const bool act = valid && idx < n;
        if (!team_any(act)) continue;
        const uint32_t *pj = jac + (size_t)idx * JACW;
        g2s(T, U::X1, pj, act);
        g2s(T, U::Y1, pj + EW, act);
        g2s(T, U::TMP, pj + 2 * EW, act);
        if (j > 0) g2s(T, U::Z2, prefix + (size_t)(idx - 1) * EW, act);
        T.sync();
        if (j == 0) T.set_one(U::Z2);
        const bool inf = T.is_zero(U::TMP);
        T.set_one(U::TMP, inf);
        T.mul(U::T0, U::INV, U::Z2);           // 1 / Z_j
        T.mul(U::INV, U::INV, U::TMP, act);    // inverse of the shorter prefix
        T.sqr(U::T1, U::T0);
        T.mul(U::X1, U::X1, U::T1);
        T.mul(U::T1, U::T1, U::T0);
        T.mul(U::Y1, U::Y1, U::T1);
        T.set_zero(U::X1, inf);
        T.set_zero(U::Y1, inf);
        T.sync();
        s2g(T, out + (size_t)idx * AFFW, U::X1, act);
        s2g(T, out + (size_t)idx * AFFW + EW, U::Y1, act);
    }
}

// out[i] = affine(P0 + i*Q), i < n.  One lane per run of B consecutive indices.
template <class G>
__global__ void __launch_bounds__(TailCfg<G>::TS::THREADS) k_synth_bases(uint32_t n, uint32_t B, const uint32_t *p0, const uint32_t *q,
                                                                         uint32_t *out, uint32_t *jac, uint32_t *prefix) {
    typedef typename G::F F;
    typedef TailCfg<G> C;
    typedef UtilSlots U;
    constexpr int EW = F::DEG * NLIMB, AFFW = 2 * EW, JACW = 3 * EW;
    extern __shared__ uint4 smem[];
    __shared__ uint32_t s_flags[C::TPB][4];
    int team;
    const Team<F> T = C::TS::make(smem, s_flags, team);
    const PtSlots s = {U::X1, U::Y1, U::Z1, U::X2, U::Y2, U::Z2, U::T0, U::T1, U::T2};
    const uint32_t id = (blockIdx.x * C::TPB + team) * 32 + (threadIdx.x & 31);
    const uint64_t start64 = (uint64_t)id * B;
    const bool valid = start64 < n;
    const uint32_t start = valid ? (uint32_t)start64 : 0u;
    bool acc_inf;
    g2s(T, U::X2, q, true);
    g2s(T, U::Y2, q + EW, true);
    small_scalar_mul(T, s, start, 31 - __clz(max(n, 2u) - 1u), acc_inf);
    g2s(T, U::X2, p0, true);
    g2s(T, U::Y2, p0 + EW, true);
    T.sync();
    Ec<F>::madd(T, s, false, true, acc_inf);
    T.sync();
    g2s(T, U::X2, q, true);
    g2s(T, U::Y2, q + EW, true);
    T.sync();
    T.set_one(U::PRE);
    for (uint32_t j = 0; j < B; ++j) {
        const uint32_t idx = start + j;
        const bool act = valid && idx < n;
        if (!team_any(act)) break;
        T.set_zero(U::Z1, acc_inf);
        T.sync();
        store_jac(T, jac + (size_t)idx * JACW, U::X1, U::Y1, U::Z1, act);
        T.copy(U::TMP, U::Z1);
        T.set_one(U::TMP, acc_inf);
        T.mul(U::PRE, U::PRE, U::TMP, act);
        T.sync();
        s2g(T, prefix + (size_t)idx * EW, U::PRE, act);
        Ec<F>::madd(T, s, false, act, acc_inf);
    }
    normalise_run<F>(T, start, B, n, valid, jac, prefix, out);
}

// ---- precomputed window tables -------------------------------------------------------------------
// jac[i] = 2^s * in[i] for n affine points (infinity, encoded y == 0, stays infinity: Z = 2*Y*Z = 0).
// A doubling needs six slots only, so this kernel runs with the 12-warps-per-SM team layout of k_batch_add.
template <class G>
struct DblCfg {
    static constexpr int DEG = G::F::DEG;
    static constexpr int TPB = 12 / DEG;     // one block of twelve warps per SM (see BaCfg, batch_affine.cuh)
    static constexpr int MINB = 1;
    typedef TeamSetup<G, 6, TPB> TS;
};
template <class G>
__global__ void __launch_bounds__(DblCfg<G>::TS::THREADS, DblCfg<G>::MINB) k_dbl_many(uint32_t n, int s_dbl, const uint32_t *in, uint32_t *jac) {
    typedef typename G::F F;
    typedef DblCfg<G> C;
    constexpr int EW = F::DEG * NLIMB;
    extern __shared__ uint4 smem[];
    __shared__ uint32_t s_flags[C::TPB][4];
    int team;
    const Team<F> T = C::TS::make(smem, s_flags, team);
    const PtSlots s = {0, 1, 2, 0, 0, 0, 3, 4, 5};   // no second operand
    const uint32_t id = (blockIdx.x * C::TPB + team) * 32 + (threadIdx.x & 31);
    const bool act = id < n;
    g2s(T, s.X1, in + (size_t)id * 2 * EW, act);
    g2s(T, s.Y1, in + (size_t)id * 2 * EW + EW, act);
    T.set_zero(s.X1, !act); T.set_zero(s.Y1, !act);
    T.set_one(s.Z1);
    T.sync();
    for (int i = 0; i < s_dbl; ++i) Ec<F>::dbl(T, s, true);
    T.sync();
    store_jac(T, jac + (size_t)id * 3 * EW, s.X1, s.Y1, s.Z1, act);
}

// out[i] = affine(jac[i]), i < n: one lane per run of B consecutive points, one inversion per lane.
template <class G>
__global__ void __launch_bounds__(TailCfg<G>::TS::THREADS) k_batch_normalise(uint32_t n, uint32_t B, const uint32_t *jac, uint32_t *prefix,
                                                                             uint32_t *out) {
    typedef typename G::F F;
    typedef TailCfg<G> C;
    typedef UtilSlots U;
    constexpr int EW = F::DEG * NLIMB, JACW = 3 * EW;
    extern __shared__ uint4 smem[];
    __shared__ uint32_t s_flags[C::TPB][4];
    int team;
    const Team<F> T = C::TS::make(smem, s_flags, team);
    const uint32_t id = (blockIdx.x * C::TPB + team) * 32 + (threadIdx.x & 31);
    const uint64_t start64 = (uint64_t)id * B;
    const bool valid = start64 < n;
    const uint32_t start = valid ? (uint32_t)start64 : 0u;
    T.set_one(U::PRE);
    for (uint32_t j = 0; j < B; ++j) {
        const uint32_t idx = start + j;
        const bool act = valid && idx < n;
        if (!team_any(act)) break;
        g2s(T, U::TMP, jac + (size_t)idx * JACW + 2 * EW, act);
        T.sync();
        const bool inf = T.is_zero(U::TMP);
        T.set_one(U::TMP, inf || !act);
        T.mul(U::PRE, U::PRE, U::TMP, act);
        T.sync();
        s2g(T, prefix + (size_t)idx * EW, U::PRE, act);
    }
    T.sync();
    normalise_run<F>(T, start, B, n, valid, jac, prefix, out);
}

// out[i] = affine(jac[i]) in the wire format (infinity -> all zero), one lane per point.
template <class G>
__global__ void __launch_bounds__(TailCfg<G>::TS::THREADS) k_to_affine(uint32_t n, const uint32_t *jac, uint32_t *out) {
    typedef typename G::F F;
    typedef TailCfg<G> C;
    typedef UtilSlots U;
    constexpr int EW = F::DEG * NLIMB;
    extern __shared__ uint4 smem[];
    __shared__ uint32_t s_flags[C::TPB][4];
    int team;
    const Team<F> T = C::TS::make(smem, s_flags, team);
    const uint32_t id = (blockIdx.x * C::TPB + team) * 32 + (threadIdx.x & 31);
    const bool act = id < n;
    load_jac(T, U::X1, U::Y1, U::Z1, jac + (size_t)id * 3 * EW, act);
    T.set_zero(U::X1, !act); T.set_zero(U::Y1, !act); T.set_zero(U::Z1, !act);
    T.sync();
    const bool inf = T.is_zero(U::Z1);
    T.set_one(U::Z1, inf);
    T.inv_all(U::INV, U::Z1, U::T0, U::T2);
    T.sqr(U::T1, U::INV);
    T.mul(U::X1, U::X1, U::T1);
    T.mul(U::T1, U::T1, U::INV);
    T.mul(U::Y1, U::Y1, U::T1);
    T.set_zero(U::X1, inf);
    T.set_zero(U::Y1, inf);
    T.sync();
    s2g(T, out + (size_t)id * 2 * EW, U::X1, act);
    s2g(T, out + (size_t)id * 2 * EW + EW, U::Y1, act);
}

// out (Jacobian) = k * P for one affine point and one Montgomery-form scalar (lane 0 of one team).
template <class G>
__global__ void __launch_bounds__(TailCfg<G>::TS::THREADS) k_scalar_mul(const uint32_t *p, const uint32_t *k_mont, uint32_t *out) {
    typedef typename G::F F;
    typedef TailCfg<G> C;
    typedef UtilSlots U;
    constexpr int EW = F::DEG * NLIMB;
    extern __shared__ uint4 smem[];
    __shared__ uint32_t s_flags[C::TPB][4];
    int team;
    const Team<F> T = C::TS::make(smem, s_flags, team);
    if (team != 0) return;
    const PtSlots s = {U::X1, U::Y1, U::Z1, U::X2, U::Y2, U::Z2, U::T0, U::T1, U::T2};
    fq_t km, k;
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) km[i] = k_mont[i];
    fq_from_mont<typename G::Fr>(k, km);
    g2s(T, U::X2, p, true);
    g2s(T, U::Y2, p + EW, true);
    T.set_zero(U::X1); T.set_zero(U::Y1); T.set_zero(U::Z1);
    T.sync();
    const bool p_inf = T.is_zero(U::Y2);
    bool acc_inf = true;
    for (int bit = MNT753_NUM_BITS - 1; bit >= 0; --bit) {
        Ec<F>::dbl(T, s, true);
        uint32_t w = 0;
#pragma unroll
        for (int i = 0; i < NLIMB; ++i) if (i == (bit >> 5)) w = k[i];
        Ec<F>::madd(T, s, false, !p_inf && ((w >> (bit & 31)) & 1u), acc_inf);
    }
    T.set_zero(U::Z1, acc_inf);
    T.set_one(U::X1, acc_inf);
    T.set_one(U::Y1, acc_inf);
    T.sync();
    store_jac(T, out, U::X1, U::Y1, U::Z1, (threadIdx.x & 31) == 0);
}

}  // namespace mnt753
