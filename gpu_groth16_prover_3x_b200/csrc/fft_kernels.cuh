// The H-polynomial of the Groth16 prover on the device: the seven radix-2 FFTs and the pointwise steps of
// compute_H (cuda_prover_piecewise.cu:14-49), which the reference runs on the CPU through libfqfft
// (basic_radix2_domain.tcc:63-126, basic_radix2_domain_aux.tcc).  SURVEY.md 8(f) rank 1: once the MSMs take
// tens of milliseconds these FFTs *are* the proof latency.
//
//   ca' = FFT( g^i * iFFT(ca) ),  cb', cc' likewise;   h = (ca' * cb' - cc') / Z(g);   H = g^-i * iFFT(h)
//
// over Fr, with the domain's omega = root_of_unity^(2^(s - log m)) (libff get_root_of_unity), the coset
// generator g = 17 and Z(g) = g^m - 1.  Arithmetic is exact, so any correct DFT gives the same words as
// libfqfft; the schedule here avoids every bit-reversal pass:
//   iFFT  = decimation-in-frequency, natural order in, bit-reversed order out
//   * g^i / m  taken from a table stored in bit-reversed order
//   FFT   = decimation-in-time, bit-reversed in, natural out
//   the last iFFT scatters through the bit reversal while it multiplies by g^-i / m.
// One thread per butterfly, one Montgomery product in registers each (fq_mul: 97 % of the IMAD.WIDE pipe).
#pragma once
#ifndef MNT753_HOST_EMU
#include <cuda_runtime.h>
#endif

#include "fq.cuh"

namespace mnt753 {

__device__ __forceinline__ void ld_fq(fq_t &x, const uint32_t *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
#pragma unroll
    for (int i = 0; i < 6; ++i) { uint4 v = q[i]; x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w; }
}
__device__ __forceinline__ void st_fq(uint32_t *p, const fq_t &x) {
    uint4 *q = reinterpret_cast<uint4 *>(p);
#pragma unroll
    for (int i = 0; i < 6; ++i) { uint4 v; v.x = x[4 * i]; v.y = x[4 * i + 1]; v.z = x[4 * i + 2]; v.w = x[4 * i + 3]; q[i] = v; }
}

enum { FC_OMEGA = 0, FC_OMEGA_INV = 1, FC_G = 2, FC_G_INV = 3, FC_M_INV = 4, FC_Z_INV = 5, FC_ONE = 6, FC_COUNT = 7 };

// the domain constants of size m = 2^logm (one thread)
template <class M>
__global__ void k_fft_consts(uint32_t *out, int logm) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    fq_t w, t, g, one, r2;
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) { w[i] = M::ROOT(i); g[i] = M::G17(i); one[i] = M::R1(i); r2[i] = M::R2(i); }
    for (int i = 0; i < M::TWO_ADICITY - logm; ++i) fq_mul<M>(w, w, w);
    st_fq(out + FC_OMEGA * NLIMB, w);
    fq_inv<M>(t, w);
    st_fq(out + FC_OMEGA_INV * NLIMB, t);
    st_fq(out + FC_G * NLIMB, g);
    fq_inv<M>(t, g);
    st_fq(out + FC_G_INV * NLIMB, t);
    fq_t mm;
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) mm[i] = (i == (logm >> 5)) ? (1u << (logm & 31)) : 0u;
    fq_mul<M>(mm, mm, r2);              // m in Montgomery form
    fq_inv<M>(t, mm);
    st_fq(out + FC_M_INV * NLIMB, t);
    fq_t z;
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) z[i] = g[i];
    for (int i = 0; i < logm; ++i) fq_mul<M>(z, z, z);   // g^m
    fq_sub<M>(z, z, one);
    fq_inv<M>(t, z);
    st_fq(out + FC_Z_INV * NLIMB, t);
    st_fq(out + FC_ONE * NLIMB, one);
}

__device__ __forceinline__ uint32_t bitrev(uint32_t i, int logn) { return logn ? (__brev(i) >> (32 - logn)) : 0u; }

// out[i] = scale * x^e,  e = i, or e = bitrev(i) when rev_logn > 0
template <class M>
__global__ void __launch_bounds__(128) k_powers(uint32_t *out, uint32_t n, const uint32_t *x, const uint32_t *scale, int rev_logn) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t e = rev_logn ? bitrev(i, rev_logn) : i;
    fq_t b, acc;
    ld_fq(b, x);
    ld_fq(acc, scale);
    for (uint32_t k = e; k; k >>= 1) {
        if (k & 1u) fq_mul<M>(acc, acc, b);
        fq_mul<M>(b, b, b);
    }
    st_fq(out + (size_t)i * NLIMB, acc);
}

// one decimation-in-frequency stage of span len:  (u, v) -> (u + v, (u - v) * w^(j * n/len))
template <class M>
__global__ void __launch_bounds__(128) k_ntt_dif(uint32_t *a, const uint32_t *tw, uint32_t n, uint32_t len) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n / 2) return;
    const uint32_t half = len >> 1, j = t % half, k = (t / half) * len;
    uint32_t *pu = a + (size_t)(k + j) * NLIMB, *pv = pu + (size_t)half * NLIMB;
    fq_t u, v, s, d;
    ld_fq(u, pu);
    ld_fq(v, pv);
    fq_add<M>(s, u, v);
    fq_sub<M>(d, u, v);
    st_fq(pu, s);
    if (j) { fq_t w; ld_fq(w, tw + (size_t)j * (n / len) * NLIMB); fq_mul<M>(d, d, w); }
    st_fq(pv, d);
}

// one decimation-in-time stage of span len:  (u, v) -> (u + w v, u - w v)
template <class M>
__global__ void __launch_bounds__(128) k_ntt_dit(uint32_t *a, const uint32_t *tw, uint32_t n, uint32_t len) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n / 2) return;
    const uint32_t half = len >> 1, j = t % half, k = (t / half) * len;
    uint32_t *pu = a + (size_t)(k + j) * NLIMB, *pv = pu + (size_t)half * NLIMB;
    fq_t u, v, s, d;
    ld_fq(u, pu);
    ld_fq(v, pv);
    if (j) { fq_t w; ld_fq(w, tw + (size_t)j * (n / len) * NLIMB); fq_mul<M>(v, v, w); }
    fq_add<M>(s, u, v);
    fq_sub<M>(d, u, v);
    st_fq(pu, s);
    st_fq(pv, d);
}


// ---- several stages per pass ---------------------------------------------------------------------------
// A block owns a TILE of T = 2^nbits elements whose indices differ only in bits [sbit, sbit + nbits) and runs
// all butterflies on those bits in shared memory: a 2^20-point transform is two passes over HBM instead of
// twenty.  Tile element u lives at global index  base + (u << sbit); in shared memory its 128-bit quad q sits
// at sm[q * T + u], so that neighbouring threads touch neighbouring 16-byte words (no bank conflicts).
// DIF runs the tile's bits from the highest down, DIT from the lowest up; the twiddle of the butterfly on
// global bit b with lower index bits j is w^(j << (logn - b - 1)), looked up in the full table (L2 resident).
constexpr int NTT_TILE_THREADS = 256;
constexpr int NTT_TILE_BITS = 10;   // at most 1024 elements = 96 KB of shared memory per block

template <class M, bool DIF>
__global__ void __launch_bounds__(NTT_TILE_THREADS) k_ntt_tile(uint32_t *a, const uint32_t *tw, int logn, int sbit, int nbits) {
    extern __shared__ uint4 sm[];
    const uint32_t T = 1u << nbits;
    const uint32_t o = blockIdx.x;
    const uint32_t lo = o & ((1u << sbit) - 1u), hi = o >> sbit;
    const size_t base = ((size_t)hi << (sbit + nbits)) | lo;
    for (uint32_t u = threadIdx.x; u < T; u += blockDim.x) {
        const uint4 *g = reinterpret_cast<const uint4 *>(a + (base + ((size_t)u << sbit)) * NLIMB);
#pragma unroll
        for (int q = 0; q < 6; ++q) sm[q * T + u] = g[q];
    }
    __syncthreads();
    for (int step = 0; step < nbits; ++step) {
        const int lb = DIF ? nbits - 1 - step : step;
        const uint32_t h = 1u << lb;
        const int b = sbit + lb;
        for (uint32_t t = threadIdx.x; t < T / 2; t += blockDim.x) {
            const uint32_t jl = t & (h - 1u), u0 = ((t >> lb) << (lb + 1)) | jl, u1 = u0 + h;
            fq_t x, y, sres, dres;
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                uint4 v = sm[q * T + u0]; x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
                v = sm[q * T + u1]; y[4 * q] = v.x; y[4 * q + 1] = v.y; y[4 * q + 2] = v.z; y[4 * q + 3] = v.w;
            }
            const size_t j = ((size_t)jl << sbit) | lo;
            const size_t e = j << (logn - b - 1);
            if (DIF) {
                fq_add<M>(sres, x, y);
                fq_sub<M>(dres, x, y);
                if (e) { fq_t w; ld_fq(w, tw + e * NLIMB); fq_mul<M>(dres, dres, w); }
            } else {
                if (e) { fq_t w; ld_fq(w, tw + e * NLIMB); fq_mul<M>(y, y, w); }
                fq_add<M>(sres, x, y);
                fq_sub<M>(dres, x, y);
            }
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                uint4 v; v.x = sres[4 * q]; v.y = sres[4 * q + 1]; v.z = sres[4 * q + 2]; v.w = sres[4 * q + 3]; sm[q * T + u0] = v;
                v.x = dres[4 * q]; v.y = dres[4 * q + 1]; v.z = dres[4 * q + 2]; v.w = dres[4 * q + 3]; sm[q * T + u1] = v;
            }
        }
        __syncthreads();
    }
    for (uint32_t u = threadIdx.x; u < T; u += blockDim.x) {
        uint4 *g = reinterpret_cast<uint4 *>(a + (base + ((size_t)u << sbit)) * NLIMB);
#pragma unroll
        for (int q = 0; q < 6; ++q) g[q] = sm[q * T + u];
    }
}

// a[i] *= tab[i]
template <class M>
__global__ void __launch_bounds__(128) k_pointwise_mul(uint32_t *a, const uint32_t *tab, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fq_t x, y;
    ld_fq(x, a + (size_t)i * NLIMB);
    ld_fq(y, tab + (size_t)i * NLIMB);
    fq_mul<M>(x, x, y);
    st_fq(a + (size_t)i * NLIMB, x);
}

// out[i] = in[i] * k      (Fr, Montgomery: used to fold the proof's r into the B1-query scalars)
template <class M>
__global__ void __launch_bounds__(128) k_scale(uint32_t *out, const uint32_t *in, const uint32_t *k, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fq_t x, y;
    ld_fq(x, in + (size_t)i * NLIMB);
    ld_fq(y, k);
    fq_mul<M>(x, x, y);
    st_fq(out + (size_t)i * NLIMB, x);
}

// a[i] = (a[i] * b[i] - c[i]) * zinv        (cuda_prover_piecewise.cu:29-39)
template <class M>
__global__ void __launch_bounds__(128) k_h_pointwise(uint32_t *a, const uint32_t *b, const uint32_t *c, const uint32_t *zinv, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fq_t x, y, z;
    ld_fq(x, a + (size_t)i * NLIMB);
    ld_fq(y, b + (size_t)i * NLIMB);
    fq_mul<M>(x, x, y);
    ld_fq(y, c + (size_t)i * NLIMB);
    fq_sub<M>(x, x, y);
    ld_fq(z, zinv);
    fq_mul<M>(x, x, z);
    st_fq(a + (size_t)i * NLIMB, x);
}

// out[bitrev(p)] = a[p] * tab[p]; out[n] = 0        (icosetFFT tail + vector_Fr_zeros(m + 1), :41-48)
template <class M>
__global__ void __launch_bounds__(128) k_h_final(uint32_t *out, const uint32_t *a, const uint32_t *tab, uint32_t n, int logn) {
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > n) return;
    fq_t x, y;
    if (p == n) {
#pragma unroll
        for (int i = 0; i < NLIMB; ++i) x[i] = 0;
        st_fq(out + (size_t)n * NLIMB, x);
        return;
    }
    ld_fq(x, a + (size_t)p * NLIMB);
    ld_fq(y, tab + (size_t)p * NLIMB);
    fq_mul<M>(x, x, y);
    st_fq(out + (size_t)bitrev(p, logn) * NLIMB, x);
}

}  // namespace mnt753
