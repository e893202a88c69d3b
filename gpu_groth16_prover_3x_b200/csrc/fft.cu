// C ABI of the H-polynomial path (include/b200_msm.h: b200msm_compute_h): host side of fft_kernels.cuh.
// Replaces compute_H<B> (cuda_prover_piecewise.cu:14-49), i.e. B::domain_iFFT / domain_cosetFFT /
// vector_Fr_muleq / vector_Fr_subeq / domain_divide_by_Z_on_coset / domain_icosetFFT
// (prover_reference_functions.cpp) on libfqfft's basic_radix2_domain.
#include <cstdlib>

#include "host_ctx.cuh"
#include "fft_kernels.cuh"

namespace {

void fft_free(FftState &f) {
    for (uint32_t **p : {&f.consts, &f.tw, &f.twi, &f.cg_br, &f.cgi_br, &f.a, &f.b, &f.c, &f.out}) {
        if (*p) cudaFree(*p);
        *p = nullptr;
    }
    f.logm = -1;
}

template <class M>
int fft_prepare(b200msm_ctx *ctx, int logm) {
    FftState &f = ctx->fft;
    if (!f.stream) {
        CU(cudaStreamCreateWithFlags(&f.stream, cudaStreamNonBlocking));
        CU(cudaEventCreate(&f.ev[0]));
        CU(cudaEventCreate(&f.ev[1]));
    }
    if (f.logm == logm) return B200MSM_OK;
    fft_free(f);
    f.staged = 0;
    CU(cudaFuncSetAttribute(k_ntt_tile<M, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 << NTT_TILE_BITS));
    CU(cudaFuncSetAttribute(k_ntt_tile<M, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 << NTT_TILE_BITS));
    const size_t m = size_t(1) << logm, EB = NLIMB * 4;
    CU(cudaMalloc(&f.consts, FC_COUNT * EB));
    CU(cudaMalloc(&f.tw, std::max<size_t>(m / 2, 1) * EB));
    CU(cudaMalloc(&f.twi, std::max<size_t>(m / 2, 1) * EB));
    CU(cudaMalloc(&f.cg_br, m * EB));
    CU(cudaMalloc(&f.cgi_br, m * EB));
    CU(cudaMalloc(&f.a, m * EB));
    CU(cudaMalloc(&f.b, m * EB));
    CU(cudaMalloc(&f.c, m * EB));
    CU(cudaMalloc(&f.out, (m + 1) * EB));
    cudaStream_t st = f.stream;
    CU(cudaEventRecord(f.ev[0], st));
    k_fft_consts<M><<<1, 32, 0, st>>>(f.consts, logm);
    const uint32_t *C = f.consts;
    const unsigned gh = (unsigned)((m / 2 + 127) / 128), gm = (unsigned)((m + 127) / 128);
    if (m >= 2) {
        k_powers<M><<<gh, 128, 0, st>>>(f.tw, (uint32_t)(m / 2), C + FC_OMEGA * NLIMB, C + FC_ONE * NLIMB, 0);
        k_powers<M><<<gh, 128, 0, st>>>(f.twi, (uint32_t)(m / 2), C + FC_OMEGA_INV * NLIMB, C + FC_ONE * NLIMB, 0);
    }
    k_powers<M><<<gm, 128, 0, st>>>(f.cg_br, (uint32_t)m, C + FC_G * NLIMB, C + FC_M_INV * NLIMB, logm);
    k_powers<M><<<gm, 128, 0, st>>>(f.cgi_br, (uint32_t)m, C + FC_G_INV * NLIMB, C + FC_M_INV * NLIMB, logm);
    CU(cudaEventRecord(f.ev[1], st));
    CU(cudaGetLastError());
    CU(cudaEventSynchronize(f.ev[1]));
    CU(cudaEventElapsedTime(&f.tables_ms, f.ev[0], f.ev[1]));
    f.logm = logm;
    return B200MSM_OK;
}



// bits of the index are processed in groups of up to NTT_TILE_BITS by k_ntt_tile; transforms of fewer than
// 2^6 points (and the B200MSM_FFT_SIMPLE=1 debugging switch) use the one-launch-per-stage kernels
bool fft_simple() {
    static const bool s = getenv("B200MSM_FFT_SIMPLE") != nullptr;
    return s;
}
template <class M, bool DIF>
void ntt_tiled(const FftState &f, uint32_t *x, int logm, const uint32_t *tw) {
    // DIF: highest bits first; DIT: lowest bits first.  Groups of equal size, at most NTT_TILE_BITS each.
    const int ngroups = (logm + NTT_TILE_BITS - 1) / NTT_TILE_BITS;
    int bits[8], start[8];
    for (int g = 0, s = 0; g < ngroups; ++g) { bits[g] = logm / ngroups + (g < logm % ngroups ? 1 : 0); start[g] = s; s += bits[g]; }
    for (int k = 0; k < ngroups; ++k) {
        const int g = DIF ? ngroups - 1 - k : k;
        const size_t smem = (size_t(96) << bits[g]);
        k_ntt_tile<M, DIF><<<1u << (logm - bits[g]), NTT_TILE_THREADS, smem, f.stream>>>(x, tw, logm, start[g], bits[g]);
    }
}
template <class M>
void ifft_dif(const FftState &f, uint32_t *x, uint32_t m, int logm) {   // natural in, bit-reversed out, NOT yet scaled by 1/m
    if (logm >= 6 && !fft_simple()) { ntt_tiled<M, true>(f, x, logm, f.twi); return; }
    for (uint32_t len = m; len >= 2; len >>= 1) k_ntt_dif<M><<<(m / 2 + 127) / 128, 128, 0, f.stream>>>(x, f.twi, m, len);
}
template <class M>
void fft_dit(const FftState &f, uint32_t *x, uint32_t m, int logm) {    // bit-reversed in, natural out
    if (logm >= 6 && !fft_simple()) { ntt_tiled<M, false>(f, x, logm, f.tw); return; }
    for (uint32_t len = 2; len <= m && len; len <<= 1) k_ntt_dit<M><<<(m / 2 + 127) / 128, 128, 0, f.stream>>>(x, f.tw, m, len);
}

// compute_H in two kinds of steps on the FFT stream, so that a caller whose three input vectors arrive one after the
// other (b200msm_prove_sharded_file reads them from the input file) can have each one uploaded and transformed while
// the next is still on its way:
//   stage(which, src)   vector `which` (0 = ca, 1 = cb, 2 = cc) -> the context's work vector, then its coset evaluation:
//                       inverse FFT, multiplication by the coset powers, forward FFT.  Asynchronous.
//   finish()            (a * b - c) / Z pointwise, inverse FFT, coset scaling; waits for the result.
template <class M>
int compute_h_check(b200msm_ctx *ctx, size_t d, int &logm) {
    const size_t m = d + 1;
    logm = 0;
    while ((size_t(1) << logm) < m) ++logm;
    if ((size_t(1) << logm) != m || logm > M::TWO_ADICITY || logm > 30)
        return fail(ctx, B200MSM_ERR_ARG, "d + 1 = %zu is not a power of two within the field's 2-adicity (2^%d)", m, M::TWO_ADICITY);
    return fft_prepare<M>(ctx, logm);
}

template <class M>
int compute_h_stage(b200msm_ctx *ctx, size_t d, int which, const uint64_t *src) {
    int logm;
    int rc = compute_h_check<M>(ctx, d, logm);
    if (rc) return rc;
    FftState &f = ctx->fft;
    if (which < 0 || which > 2 || (f.staged & (1u << which))) return fail(ctx, B200MSM_ERR_ARG, "compute_H: vector %d out of order", which);
    cudaStream_t st = f.stream;
    const size_t m = d + 1;
    const unsigned gm = (unsigned)((m + 127) / 128);
    if (f.staged == 0) CU(cudaEventRecord(f.ev[0], st));
    uint32_t *x = which == 0 ? f.a : (which == 1 ? f.b : f.c);
    CU(cudaMemcpyAsync(x, src, m * NLIMB * 4, cudaMemcpyDefault, st));
    ifft_dif<M>(f, x, (uint32_t)m, logm);                         // coset evaluation of the vector
    k_pointwise_mul<M><<<gm, 128, 0, st>>>(x, f.cg_br, (uint32_t)m);
    fft_dit<M>(f, x, (uint32_t)m, logm);
    CU(cudaGetLastError());
    f.staged |= 1u << which;
    return B200MSM_OK;
}

template <class M>
int compute_h_finish(b200msm_ctx *ctx, size_t d, uint64_t *out_host, const uint64_t **out_dev) {
    int logm;
    int rc = compute_h_check<M>(ctx, d, logm);
    if (rc) return rc;
    FftState &f = ctx->fft;
    if (f.staged != 7u) { f.staged = 0; return fail(ctx, B200MSM_ERR_ARG, "compute_H: not all three vectors were staged"); }
    f.staged = 0;
    cudaStream_t st = f.stream;
    const size_t m = d + 1;
    const unsigned gm = (unsigned)((m + 127) / 128), gm1 = (unsigned)((m + 1 + 127) / 128);
    k_h_pointwise<M><<<gm, 128, 0, st>>>(f.a, f.b, f.c, f.consts + FC_Z_INV * NLIMB, (uint32_t)m);
    ifft_dif<M>(f, f.a, (uint32_t)m, logm);
    k_h_final<M><<<gm1, 128, 0, st>>>(f.out, f.a, f.cgi_br, (uint32_t)m, logm);
    if (out_host) CU(cudaMemcpyAsync(out_host, f.out, (m + 1) * NLIMB * 4, cudaMemcpyDefault, st));
    CU(cudaEventRecord(f.ev[1], st));
    CU(cudaGetLastError());
    CU(cudaEventSynchronize(f.ev[1]));
    CU(cudaEventElapsedTime(&f.last_ms, f.ev[0], f.ev[1]));
    if (out_dev) *out_dev = reinterpret_cast<const uint64_t *>(f.out);
    return B200MSM_OK;
}

template <class M>
int compute_h_impl(b200msm_ctx *ctx, size_t d, const uint64_t *ca, const uint64_t *cb, const uint64_t *cc, uint64_t *out_host,
                   const uint64_t **out_dev) {
    ctx->fft.staged = 0;
    int rc = compute_h_stage<M>(ctx, d, 0, ca);
    if (!rc) rc = compute_h_stage<M>(ctx, d, 1, cb);
    if (!rc) rc = compute_h_stage<M>(ctx, d, 2, cc);
    if (rc) { ctx->fft.staged = 0; return rc; }
    return compute_h_finish<M>(ctx, d, out_host, out_dev);
}

}  // namespace

// out_dev[i] = in_dev[i] * k over Fr of the context's curve (device pointers; k: 24 words on the device), on `st`
int b200msm_internal_fr_scale(b200msm_ctx *ctx, size_t n, const uint32_t *in_dev, const uint32_t *k_dev, uint32_t *out_dev, cudaStream_t st) {
    const unsigned g = (unsigned)((n + 127) / 128);
    if (ctx->curve == B200MSM_MNT4753) k_scale<ModB><<<g, 128, 0, st>>>(out_dev, in_dev, k_dev, (uint32_t)n);
    else k_scale<ModA><<<g, 128, 0, st>>>(out_dev, in_dev, k_dev, (uint32_t)n);
    CU(cudaGetLastError());
    return B200MSM_OK;
}

// internal (prover.cu): build the domain tables for d + 1 points ahead of the first proof
int b200msm_internal_fft_prepare(b200msm_ctx *ctx, size_t d) {
    int logm = 0;
    while ((size_t(1) << logm) < d + 1) ++logm;
    if ((size_t(1) << logm) != d + 1) return fail(ctx, B200MSM_ERR_ARG, "d + 1 = %zu is not a power of two", d + 1);
    CU(cudaSetDevice(ctx->device));
    if (ctx->curve == B200MSM_MNT4753) return logm <= ModB::TWO_ADICITY ? fft_prepare<ModB>(ctx, logm) : fail(ctx, B200MSM_ERR_ARG, "domain too large");
    return logm <= ModA::TWO_ADICITY ? fft_prepare<ModA>(ctx, logm) : fail(ctx, B200MSM_ERR_ARG, "domain too large");
}

// internal (prover.cu): compute_H fed one vector at a time, see compute_h_stage above.  The scalar field of a curve is the
// base field of the other one.
int b200msm_internal_compute_h_stage(b200msm_ctx *ctx, size_t d, int which, const uint64_t *src) {
    if (!ctx || !src) return B200MSM_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    return ctx->curve == B200MSM_MNT4753 ? compute_h_stage<ModB>(ctx, d, which, src) : compute_h_stage<ModA>(ctx, d, which, src);
}
int b200msm_internal_compute_h_finish(b200msm_ctx *ctx, size_t d, const uint64_t **out_dev) {
    if (!ctx || !out_dev) return B200MSM_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    return ctx->curve == B200MSM_MNT4753 ? compute_h_finish<ModB>(ctx, d, nullptr, out_dev) : compute_h_finish<ModA>(ctx, d, nullptr, out_dev);
}
// forget staged vectors and wait for whatever the FFT stream still reads from host memory (error paths)
void b200msm_internal_compute_h_abort(b200msm_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    ctx->fft.staged = 0;
    if (ctx->fft.stream) cudaStreamSynchronize(ctx->fft.stream);
}

extern "C" {

int b200msm_compute_h(b200msm_ctx *ctx, size_t d, const uint64_t *ca, const uint64_t *cb, const uint64_t *cc, uint64_t *out_host,
                      const uint64_t **out_dev) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (!ca || !cb || !cc || (!out_host && !out_dev)) return fail(ctx, B200MSM_ERR_ARG, "null pointer");
    CU(cudaSetDevice(ctx->device));
    // the scalar field of a curve is the base field of the other one
    return ctx->curve == B200MSM_MNT4753 ? compute_h_impl<ModB>(ctx, d, ca, cb, cc, out_host, out_dev)
                                         : compute_h_impl<ModA>(ctx, d, ca, cb, cc, out_host, out_dev);
}

int b200msm_compute_h_timings(b200msm_ctx *ctx, float ms[2]) {
    if (!ctx || !ms) return B200MSM_ERR_ARG;
    ms[0] = ctx->fft.last_ms;
    ms[1] = ctx->fft.tables_ms;
    return B200MSM_OK;
}

void b200msm_fft_release(b200msm_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->fft.stream) cudaStreamSynchronize(ctx->fft.stream);
    fft_free(ctx->fft);
    if (ctx->fft.stream) {
        cudaEventDestroy(ctx->fft.ev[0]);
        cudaEventDestroy(ctx->fft.ev[1]);
        cudaStreamDestroy(ctx->fft.stream);
        ctx->fft.stream = nullptr;
    }
}

}  // extern "C"
