// The whole `compute` step of the reference's GPU prover behind the C ABI: proving key resident in HBM, one
// call per proof.  b200msm_prove is run_prover of cuda_prover_piecewise.cu:96-230 with
//   * the five multi-exponentiations A, B1, B2, L, H on the engine (A and H ran on the CPU in the reference),
//   * compute_H on the device (b200msm_compute_h),
//   * the proof assembly  C = Ht + Lt + r * Bt1  and the three affine normalisations on the device
//     (the reference: libff on the host, lines 198-204),
// writing the proof bytes of groth16_output_write (A || B || C, affine, infinity as zeros;
// libsnark/serialization.hpp:43-67).  It only composes the other entry points of include/b200_msm.h.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <cerrno>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include "host_ctx.cuh"

struct b200msm_key {
    size_t d = 0, m = 0;
    int shard = 0, nshards = 1;          // this key holds points [lo[q], lo[q] + cnt[q]) of every query q
    size_t lo[5] = {0, 0, 0, 0, 0}, cnt[5] = {0, 0, 0, 0, 0};
    size_t w_lo = 0, w_cnt = 0;          // witness elements the shard needs: w[w_lo, w_lo + w_cnt)
    int slot[5] = {-1, -1, -1, -1, -1};  // A, B1, B2, L, H
    uint32_t *w_dev = nullptr, *rw_dev = nullptr, *r_dev = nullptr;   // witness, r * witness, r (device, reused per proof)
    uint32_t *h_dev = nullptr;           // shards > 0: their slice of the H coefficients (copied from shard 0's GPU)
    cudaStream_t stream = nullptr;
    cudaEvent_t ready = nullptr;
    cudaEvent_t t0 = nullptr;            // start of the proof on this shard (B200MSM_TRACE=1: per-lane timeline)
};

int b200msm_internal_fr_scale(b200msm_ctx *ctx, size_t n, const uint32_t *in_dev, const uint32_t *k_dev, uint32_t *out_dev, cudaStream_t st);
extern "C" int b200msm_internal_reserve(b200msm_ctx *ctx, int lane, int slot, size_t n);
int b200msm_internal_fft_prepare(b200msm_ctx *ctx, size_t d);
int b200msm_internal_compute_h_stage(b200msm_ctx *ctx, size_t d, int which, const uint64_t *src);
int b200msm_internal_compute_h_finish(b200msm_ctx *ctx, size_t d, const uint64_t **out_dev);
void b200msm_internal_compute_h_abort(b200msm_ctx *ctx);
const GroupOps &b200msm_internal_ops(int curve, int group);

namespace {
inline int g2_deg(const b200msm_ctx *ctx) { return ctx->curve == B200MSM_MNT4753 ? 2 : 3; }

// Bytes [off, off + len) of an open file into dst.  A witness of 2^20 elements is 100 MB and the H coefficients three
// times that: one thread copies out of the page cache at 6-7 GB/s (64 ms for the default MNT4753 input, more than
// the five MSMs take on eight GPUs: profiles/r02_proof_8gpu_lane_trace.txt), so large ranges are cut into slices of
// at least 8 MB read by up to eight threads with pread.
constexpr double READ_BYTES_PER_NS = 10.0;     // what the planner of the lane split expects of it
bool read_slice(int fd, char *dst, size_t off, size_t len) {
    while (len) {
        const ssize_t got = pread(fd, dst, len, (off_t)off);
        if (got < 0 && errno == EINTR) continue;
        if (got <= 0) return false;                       // error, or the file ends before the range does
        dst += got; off += (size_t)got; len -= (size_t)got;
    }
    return true;
}
bool read_range(int fd, char *dst, size_t off, size_t len) {
    const size_t min_slice = size_t(8) << 20;
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const size_t nt = std::min<size_t>(std::min<size_t>(8, hw), std::max<size_t>(1, len / min_slice));
    if (nt <= 1) return read_slice(fd, dst, off, len);
    std::vector<char> okv(nt, 0);
    std::vector<std::thread> th;
    const size_t per = ((len + nt - 1) / nt + 4095) & ~size_t(4095);
    for (size_t t = 0; t < nt; ++t) {
        const size_t lo = std::min(len, t * per), hi = std::min(len, lo + per);
        th.emplace_back([=, &okv] { okv[t] = read_slice(fd, dst + lo, off + lo, hi - lo) ? 1 : 0; });
    }
    for (auto &x : th) x.join();
    for (char o : okv) if (!o) return false;
    return true;
}


// One pass over every kernel a proof launches, on zeros: a small MSM per query on its lane and, on shard 0, compute_H.
// The CUDA runtime loads a kernel at its FIRST launch and that load waits for the kernels already running; with the
// queries of a proof side by side on persistent kernels (lane_split_plan below) the first proof would serialise on
// those loads -- measured: 47 ms for the first MNT6753 default proof against 31 ms for the following ones.
int key_warm_up(b200msm_ctx *ctx, const b200msm_key *key) {
    const size_t nw = 1024, fft_n = key->shard == 0 ? key->d + 1 : 0;
    void *zeros = nullptr;
    const size_t bytes = std::max<size_t>(nw, 3 * fft_n) * 96;
    CU(cudaMalloc(&zeros, bytes));
    int rc = B200MSM_OK;
    uint64_t out[5][108];
    if (cudaMemset(zeros, 0, bytes) != cudaSuccess) rc = fail(ctx, B200MSM_ERR_CUDA, "cudaMemset failed");
    for (int q = 0; q < 5 && !rc; ++q)
        rc = b200msm_msm_async(ctx, q, key->slot[q], 0, static_cast<const uint64_t *>(zeros), std::min(nw, key->cnt[q]), out[q]);
    if (!rc && fft_n) {
        const uint64_t *z = static_cast<const uint64_t *>(zeros), *h = nullptr;
        rc = b200msm_compute_h(ctx, key->d, z, z + fft_n * 12, z + 2 * fft_n * 12, nullptr, &h);
    }
    for (int q = 0; q < 5; ++q)
        if (ctx->lanes[q].pending) { const int rcw = b200msm_wait(ctx, q); if (!rc) rc = rcw; }
    cudaFree(zeros);
    return rc;
}
}

extern "C" {

void b200msm_key_free(b200msm_ctx *ctx, b200msm_key *key) {
    if (!ctx || !key) return;
    for (int s : key->slot)
        if (s >= 0) b200msm_bases_free(ctx, s);
    cudaSetDevice(ctx->device);
    if (key->w_dev) cudaFree(key->w_dev);
    if (key->rw_dev) cudaFree(key->rw_dev);
    if (key->r_dev) cudaFree(key->r_dev);
    if (key->h_dev) cudaFree(key->h_dev);
    if (key->ready) cudaEventDestroy(key->ready);
    if (key->t0) cudaEventDestroy(key->t0);
    if (key->stream) cudaStreamDestroy(key->stream);
    delete key;
}

int b200msm_key_load_shard(b200msm_ctx *ctx, const void *params_image, size_t bytes, int shard, int nshards, b200msm_key **out) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (!params_image || !out || bytes < 16 || nshards < 1 || shard < 0 || shard >= nshards) return fail(ctx, B200MSM_ERR_ARG, "bad argument");
    *out = nullptr;
    // layout of <curve>-parameters (generate_parameters.cpp:59-108): u64 d, u64 m, A[m+1] G1, B1[m+1] G1,
    // B2[m+1] G2, L[m-1] G1, H[d] G1, every point affine x || y in Montgomery limbs
    const uint64_t *p = static_cast<const uint64_t *>(params_image);
    const size_t d = p[0], m = p[1], g1 = 24, g2 = 24 * (size_t)g2_deg(ctx);
    if (m < 1 || d > (size_t(1) << 30) || m > (size_t(1) << 30)) return fail(ctx, B200MSM_ERR_ARG, "implausible d = %zu, m = %zu", d, m);
    const size_t want = 16 + 8 * (2 * (m + 1) * g1 + (m + 1) * g2 + (m - 1) * g1 + d * g1);
    if (bytes != want) return fail(ctx, B200MSM_ERR_ARG, "parameter image of %zu bytes, expected %zu for d = %zu, m = %zu", bytes, want, d, m);
    b200msm_key *key = new b200msm_key;
    key->d = d;
    key->m = m;
    key->shard = shard;
    key->nshards = nshards;
    p += 2;
    const int group[5] = {B200MSM_G1, B200MSM_G1, B200MSM_G2, B200MSM_G1, B200MSM_G1};
    const size_t count[5] = {m + 1, m + 1, m + 1, m - 1, d}, words[5] = {g1, g1, g2, g1, g1};
    for (int q = 0; q < 5; ++q) {
        // point-range sharding: shard g of G owns [N g / G, N (g + 1) / G) of every query
        b200msm_shard_range(count[q], shard, nshards, &key->lo[q], &key->cnt[q]);
        int rc = b200msm_bases_upload(ctx, group[q], p + key->lo[q] * words[q], key->cnt[q], &key->slot[q]);
        if (rc) { b200msm_key_free(ctx, key); return rc; }
        p += count[q] * words[q];
    }
    // the witness slice the shard reads: A, B1, B2 take w[lo, lo + cnt), L takes w[2 + lo_L, 2 + lo_L + cnt_L)  (main.cpp:214-217)
    const size_t a_lo = key->lo[0], a_hi = a_lo + key->cnt[0], l_lo = 2 + key->lo[3], l_hi = l_lo + key->cnt[3];
    key->w_lo = a_lo < l_lo ? a_lo : l_lo;
    key->w_cnt = (a_hi > l_hi ? a_hi : l_hi) - key->w_lo;
    CU(cudaSetDevice(ctx->device));
    bool ok = cudaMalloc(&key->w_dev, (key->w_cnt + 1) * 96) == cudaSuccess && cudaMalloc(&key->rw_dev, (key->cnt[0] + 1) * 96) == cudaSuccess &&
              cudaMalloc(&key->r_dev, 96) == cudaSuccess &&
              cudaStreamCreateWithFlags(&key->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&key->ready, cudaEventDisableTiming) == cudaSuccess && cudaEventCreate(&key->t0) == cudaSuccess;
    if (ok && shard > 0) ok = cudaMalloc(&key->h_dev, (key->cnt[4] + 1) * 96) == cudaSuccess;
    if (!ok) { b200msm_key_free(ctx, key); return fail(ctx, B200MSM_ERR_OOM, "cannot allocate the witness buffers"); }
    // warm everything the first proof would otherwise pay for: lane arenas (query q runs on lane q), domain tables
    int rc = b200msm_internal_reserve(ctx, 0, key->slot[0], key->cnt[0]);
    if (!rc) rc = b200msm_internal_reserve(ctx, 4, key->slot[4], key->cnt[4]);
    if (!rc) rc = b200msm_internal_reserve(ctx, 1, key->slot[1], key->cnt[1]);
    if (!rc) rc = b200msm_internal_reserve(ctx, 2, key->slot[2], key->cnt[2]);
    if (!rc) rc = b200msm_internal_reserve(ctx, 3, key->slot[3], key->cnt[3]);
    if (!rc && shard == 0) rc = b200msm_internal_fft_prepare(ctx, d);
    if (!rc) rc = key_warm_up(ctx, key);
    if (rc) { b200msm_key_free(ctx, key); return rc; }
    *out = key;
    return B200MSM_OK;
}

int b200msm_key_load(b200msm_ctx *ctx, const void *params_image, size_t bytes, b200msm_key **out) {
    return b200msm_key_load_shard(ctx, params_image, bytes, 0, 1, out);
}

// Every GPU takes its point range of every query out of ONE image of the parameter file; one host thread per GPU, so
// that the uploads and window-table builds of the shards run side by side (key load on eight GPUs: 1.8 s, on one: 8.2 s).
int b200msm_key_load_sharded_file(b200msm_ctx *const *ctxs, int nshards, const char *path, b200msm_key **keys) {
    if (!ctxs || nshards < 1 || !ctxs[0]) return B200MSM_ERR_ARG;
    b200msm_ctx *ctx = ctxs[0];
    if (!path || !keys) return fail(ctx, B200MSM_ERR_ARG, "bad argument");
    for (int g = 0; g < nshards; ++g) {
        if (!ctxs[g]) return fail(ctx, B200MSM_ERR_ARG, "null context for shard %d", g);
        keys[g] = nullptr;
    }
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(ctx, B200MSM_ERR_ARG, "cannot open %s", path);
    struct stat st;
    const size_t n = fstat(fd, &st) == 0 && st.st_size > 0 ? (size_t)st.st_size : 0;
    std::unique_ptr<char[]> buf(n ? new (std::nothrow) char[n] : nullptr);
    const bool ok = buf && read_range(fd, buf.get(), 0, n);
    close(fd);
    if (!ok) return fail(ctx, B200MSM_ERR_ARG, "cannot read %s", path);
    std::vector<int> rcs((size_t)nshards, B200MSM_OK);
    if (nshards == 1) rcs[0] = b200msm_key_load_shard(ctxs[0], buf.get(), n, 0, 1, &keys[0]);
    else {
        std::vector<std::thread> loaders;
        for (int g = 0; g < nshards; ++g)
            loaders.emplace_back([&, g] { rcs[(size_t)g] = b200msm_key_load_shard(ctxs[g], buf.get(), n, g, nshards, &keys[g]); });
        for (auto &th : loaders) th.join();
    }
    int rc = B200MSM_OK;
    for (int g = 0; g < nshards; ++g)
        if (rcs[(size_t)g] && !rc) {
            rc = rcs[(size_t)g];
            if (g > 0) ctx->err = "shard " + std::to_string(g) + ": " + ctxs[g]->err;
        }
    if (rc)
        for (int g = 0; g < nshards; ++g)
            if (keys[g]) { b200msm_key_free(ctxs[g], keys[g]); keys[g] = nullptr; }
    return rc;
}

int b200msm_key_load_file(b200msm_ctx *ctx, const char *path, b200msm_key **out) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (!out) return fail(ctx, B200MSM_ERR_ARG, "bad argument");
    return b200msm_key_load_sharded_file(&ctx, 1, path, out);
}

int b200msm_key_info(const b200msm_key *key, uint64_t info[2]) {
    if (!key || !info) return B200MSM_ERR_ARG;
    info[0] = key->d;
    info[1] = key->m;
    return B200MSM_OK;
}

void *b200msm_pinned_alloc(size_t bytes) {
    void *p = nullptr;
    return cudaMallocHost(&p, bytes ? bytes : 1) == cudaSuccess ? p : nullptr;
}
void b200msm_pinned_free(void *p) {
    if (p) cudaFreeHost(p);
}

size_t b200msm_proof_bytes(const b200msm_ctx *ctx) { return ctx ? (size_t)(2 * 192 + 192 * g2_deg(ctx)) : 0; }
size_t b200msm_input_bytes(const b200msm_key *key) { return key ? ((key->m + 1) + 3 * (key->d + 1) + 1) * 96 : 0; }

}  // extern "C"

namespace {
struct Partials { uint64_t A[36], rB1[36], B2[108], L[36], H[36]; };

// ---- the five MSMs of a proof side by side -----------------------------------------------------------------------
// The accumulation and the first reduction rounds of an MSM are persistent kernels of one block per SM, so MSMs
// enqueued on different lanes still run one after the other, each paying the latency of its rounds, of its reduction
// tree and of the serial tail on a GPU that is mostly idle (a 2^15-point G1 MSM: 5.4 ms, of which about 1 ms is
// arithmetic).  When the proof is small the lanes get DISJOINT parts of the GPU instead (b200msm_set_lane_sms).
// Five parties: the A, B1, B2 and L queries (lanes 0-3) and, on lane 4, the H query behind the FFTs of compute_H,
// which run on the SMs that lane 4 is going to use.  With work W_q (shrinks with the SMs) and latency L_q (does not)
// from the window-choice model, party q gets the fraction W_q / (T - L_q) of the SMs, T being the common finishing
// time that makes the fractions sum to one.  B200MSM_LANE_SPLIT: 0 = never, 1 (default) = when the model gains at
// least 15 % and the work is below 100 ms (beyond that the proof is bound by throughput and an imperfect split costs
// more than the latencies it hides: measured, profiles/r02_lane_split_ab.txt), 2 = always.
int lane_split_mode() {
    static const int mode = [] { const char *e = getenv("B200MSM_LANE_SPLIT"); return e ? atoi(e) : 1; }();
    return mode;
}

// B200MSM_TRACE=1: after every proof, the timeline of each lane on stderr (development; tools/lane_split_ab.sh)
bool trace_on() {
    static const bool on = [] { const char *e = getenv("B200MSM_TRACE"); return e && *e && *e != '0'; }();
    return on;
}
void trace_lanes(b200msm_ctx *ctx, const b200msm_key *key) {
    static const char *name[5] = {"A", "B1", "B2", "L", "H"};
    cudaSetDevice(ctx->device);
    for (int l = 0; l < NLANES; ++l) {
        const Lane &ln = ctx->lanes[l];
        if (!ln.timed) continue;
        float t[6] = {0, 0, 0, 0, 0, 0};
        for (int e = 0; e < 6; ++e) cudaEventElapsedTime(&t[e], key->t0, ln.ev[e]);
        fprintf(stderr, "[trace] shard %d lane %d %-2s sms %3d | start %7.2f sorted %7.2f accumulated %7.2f reduced %7.2f end %7.2f ms\n", key->shard, l, name[l],
                ctx->lane_sms[l], t[0], t[2], t[3], t[4], t[5]);
    }
    if (key->shard == 0) {
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, key->t0, ctx->fft.ev[0]);
        cudaEventElapsedTime(&b, key->t0, ctx->fft.ev[1]);
        fprintf(stderr, "[trace] shard 0 compute_H | start %7.2f end %7.2f ms\n", a, b);
    }
}

void lane_split_clear(b200msm_ctx *ctx) {
    for (int l = 0; l < NLANES; ++l) ctx->lane_sms[l] = 0;
}

// cost[q]: query q alone on the whole GPU; fft_delay_ns: how long after the witness MSMs the FFTs of compute_H can start
// (b200msm_prove_sharded_file reads their input, three quarters of the file, while the witness MSMs run).  Returns
// true (and the SMs of each lane) when the queries should run side by side.
bool lane_split_compute(const MsmCost cost[5], bool shard0, size_t d, double fft_delay_ns, int sm_count, int mode, int sms[5], double est_ns[2]) {
    double W[5], L[5];
    for (int q = 0; q < 5; ++q) { W[q] = cost[q].work_ns; L[q] = cost[q].latency_ns; sms[q] = 0; }
    // compute_H before the H query: seven transforms of M = d + 1 points, (M / 2) log2 M butterflies each, and the
    // pointwise passes, at the multiplier's 7.7 G products per second; three vectors of M elements cross PCIe first.
    // Shard 0 runs them on lane 4's SMs (work of that party); the other shards wait for them (latency: about a
    // third of shard 0's GPU is what the split gives them).
    const double M = double(d + 1);
    const double fft_w = 0.13 * (3.5 * log2(M > 2 ? M : 2) + 6.0) * M, fft_l = 3.0 * M * 96.0 / 40.0 + 150000.0;
    double serial = 0;
    for (int q = 0; q < 4; ++q) serial += W[q] + L[q];                 // one after the other on the whole GPU ...
    if (shard0) serial += fft_w;                                       // ... the FFTs squeezed in between, or after their input has arrived
    serial = std::max(serial, fft_delay_ns + fft_l + fft_w) + W[4] + L[4];
    if (W[4] > 0) {
        if (shard0) { W[4] += fft_w; L[4] += fft_delay_ns + fft_l; }
        else L[4] += fft_delay_ns + fft_l + 3.0 * fft_w;
    }
    double sumW = 0, lo = 0;
    for (int q = 0; q < 5; ++q) { sumW += W[q]; if (W[q] > 0 && L[q] > lo) lo = L[q]; }
    est_ns[0] = serial;
    est_ns[1] = serial;
    if (sumW <= 0 || mode <= 0) return false;
    double hi = lo + sumW;                                            // every term W_q / (hi - L_q) <= W_q / sumW
    for (int it = 0; it < 60; ++it) {
        const double T = 0.5 * (lo + hi);
        double f = 0;
        for (int q = 0; q < 5; ++q) if (W[q] > 0) f += W[q] / (T - L[q]);
        if (f > 1.0) lo = T; else hi = T;
    }
    const double T = hi;
    est_ns[1] = T;
    if (mode == 1 && !(T <= 0.85 * serial && sumW <= 100e6)) return false;
    // whole SMs: a handful at least for lane 4 (the FFT launches are short and many), one for any other MSM, the
    // rounding goes to (or comes from) the largest party
    int total = 0, big = 0;
    for (int q = 0; q < 5; ++q) {
        if (W[q] <= 0) continue;
        sms[q] = std::max(q == 4 ? std::min(8, std::max(1, sm_count / 8)) : 1, (int)(W[q] / (T - L[q]) * sm_count));
        total += sms[q];
        if (W[q] > W[big]) big = q;
    }
    while (total > sm_count && sms[big] > 1) { --sms[big]; --total; }
    if (total < sm_count) sms[big] += sm_count - total;
    return true;
}

void lane_split_plan(b200msm_ctx *ctx, const b200msm_key *key, double fft_delay_ns) {
    lane_split_clear(ctx);
    MsmCost cost[5];
    for (int q = 0; q < 5; ++q) {
        const BaseSet &bs = ctx->sets[key->slot[q]];
        cost[q] = key->cnt[q] ? model_cost(key->cnt[q], degree_of(ctx->curve, bs.group), cfg_for_set(ctx, bs, key->cnt[q])) : MsmCost{0.0, 0.0};
    }
    int sms[5];
    double est[2];
    if (lane_split_compute(cost, key->shard == 0, key->d, fft_delay_ns, ctx->sm_count, lane_split_mode(), sms, est))
        for (int q = 0; q < 5; ++q) ctx->lane_sms[q] = sms[q];
}

// First half of a proof on one shard: its witness slice and r on its device, its four witness MSMs in flight.
int prove_begin(b200msm_ctx *ctx, const b200msm_key *key, const uint64_t *w, const uint64_t *r, double fft_delay_ns, Partials &P) {
    int rc;
    // The witness crosses PCIe once; the B1 query runs on r * w so that its result is the r * Bt1 term of C directly
    // (sum (r w_i) B1_i = r * sum w_i B1_i: the same group element, no 753-step scalar multiplication afterwards).
    CU(cudaSetDevice(ctx->device));
    CU(cudaEventRecord(key->t0, key->stream));
    CU(cudaMemcpyAsync(key->w_dev, w + key->w_lo * 12, key->w_cnt * 96, cudaMemcpyHostToDevice, key->stream));
    CU(cudaMemcpyAsync(key->r_dev, r, 96, cudaMemcpyHostToDevice, key->stream));
    const uint32_t *wa = key->w_dev + (key->lo[0] - key->w_lo) * 24;          // first witness element of the A / B1 / B2 slice
    const uint32_t *wl = key->w_dev + (2 + key->lo[3] - key->w_lo) * 24;      // and of the L slice: w[2..m] (:167)
    if ((rc = b200msm_internal_fr_scale(ctx, key->cnt[0], wa, key->r_dev, key->rw_dev, key->stream))) return rc;
    CU(cudaEventRecord(key->ready, key->stream));
    for (int l = 0; l < 4; ++l) CU(cudaStreamWaitEvent(ctx->lanes[l].stream, key->ready, 0));
    // the four witness MSMs in flight together (cuda_prover_piecewise.cu:162-167), side by side when the proof is small
    lane_split_plan(ctx, key, fft_delay_ns);
    if ((rc = b200msm_msm_async(ctx, 0, key->slot[0], 0, reinterpret_cast<const uint64_t *>(wa), key->cnt[0], P.A))) return rc;
    if ((rc = b200msm_msm_async(ctx, 1, key->slot[1], 0, reinterpret_cast<const uint64_t *>(key->rw_dev), key->cnt[1], P.rB1))) return rc;
    if ((rc = b200msm_msm_async(ctx, 2, key->slot[2], 0, reinterpret_cast<const uint64_t *>(wa), key->cnt[2], P.B2))) return rc;
    if ((rc = b200msm_msm_async(ctx, 3, key->slot[3], 0, reinterpret_cast<const uint64_t *>(wl), key->cnt[3], P.L))) return rc;
    return B200MSM_OK;
}
void prove_drain(b200msm_ctx *const *ctxs, int n) {
    b200msm_internal_compute_h_abort(ctxs[0]);
    for (int g = 0; g < n; ++g) {
        for (int l = 0; l < NLANES; ++l)
            if (ctxs[g]->lanes[l].pending) b200msm_wait(ctxs[g], l);
        lane_split_clear(ctxs[g]);
    }
}

// Second half: the H polynomial on shard 0's GPU beside the MSMs, its coefficients handed to the other shards over
// NVLink, the H query, the fold of the partial points, the assembly and the proof bytes.
// ca == nullptr: the three vectors have been staged already (b200msm_internal_compute_h_stage), only the finish is left.
int prove_finish(b200msm_ctx *const *ctxs, b200msm_key *const *keys, int n, const uint64_t *ca, const uint64_t *cb, const uint64_t *cc,
                 Partials *P, uint8_t *proof) {
    b200msm_ctx *c0 = ctxs[0];
    const size_t d = keys[0]->d;
    const int dg = g2_deg(c0);
    const uint64_t *h_dev = nullptr;
    int rc = ca ? b200msm_compute_h(c0, d, ca, cb, cc, nullptr, &h_dev) : b200msm_internal_compute_h_finish(c0, d, &h_dev);
    for (int g = 0; g < n && !rc; ++g) {
        const uint64_t *hs = h_dev + keys[g]->lo[4] * 12;
        if (g > 0) {
            if (cudaSetDevice(ctxs[g]->device) != cudaSuccess) { rc = fail(c0, B200MSM_ERR_CUDA, "cannot select device %d", ctxs[g]->device); break; }
            // on lane 4's stream of the receiving GPU: ordered before the H query enqueued on the same stream below,
            // and no device-wide synchronisation while lanes 0-3 are still working (the source is complete:
            // b200msm_compute_h is synchronous)
            if (cudaMemcpyPeerAsync(keys[g]->h_dev, ctxs[g]->device, hs, c0->device, keys[g]->cnt[4] * 96, ctxs[g]->lanes[4].stream) != cudaSuccess) {
                rc = fail(c0, B200MSM_ERR_CUDA, "peer copy of the H coefficients to device %d failed", ctxs[g]->device);
                break;
            }
            hs = reinterpret_cast<const uint64_t *>(keys[g]->h_dev);
        }
        rc = b200msm_msm_async(ctxs[g], 4, keys[g]->slot[4], 0, hs, keys[g]->cnt[4], P[g].H);   // beside the witness MSMs still running
    }
    for (int g = 0; g < n; ++g) {
        for (int l = 0; l < NLANES; ++l)
            if (ctxs[g]->lanes[l].pending) { const int rcw = b200msm_wait(ctxs[g], l); if (!rc) rc = rcw; }
        if (!rc && trace_on()) trace_lanes(ctxs[g], keys[g]);
        lane_split_clear(ctxs[g]);
    }
    if (rc) return rc;
    // Fold the shards' partial points and normalise, all on shard 0's GPU: A = sum A_g, B = sum B2_g,
    // C = sum (H_g + L_g + r Bt1_g)  (:198-200) -- one upload of the 5 n partial points, three folds and three affine
    // normalisations chained on the context's tail stream, one download of the proof bytes (A || B || C).
    const size_t J1 = 72, J2 = 72 * (size_t)dg, A1 = 48, A2 = 48 * (size_t)dg;      // 32-bit words of a Jacobian / affine point
    const size_t in_words = (size_t)n * J1 + (size_t)n * J2 + (size_t)3 * n * J1;
    const size_t sc_words = fold_scratch_points((size_t)3 * n) * J2;               // scratch of the largest fold, in the larger point size
    const size_t out_words = 2 * A1 + A2;
    b200msm_ctx *ctx = c0;      // for CU()
    CU(cudaSetDevice(c0->device));                 // the waits above left the last shard's device current
    if ((rc = tail_reserve(c0, (in_words + sc_words + 2 * J1 + J2 + out_words) * 4, std::max(in_words, out_words) * 4))) return rc;
    TailBuf &t = c0->tail;
    uint32_t *h = reinterpret_cast<uint32_t *>(t.h);
    uint32_t *hA = h, *hB = hA + (size_t)n * J1, *hC = hB + (size_t)n * J2;
    for (int g = 0; g < n; ++g) {
        memcpy(hA + (size_t)g * J1, P[g].A, J1 * 4);
        memcpy(hB + (size_t)g * J2, P[g].B2, J2 * 4);
        memcpy(hC + (size_t)(3 * g) * J1, P[g].H, J1 * 4);
        memcpy(hC + (size_t)(3 * g + 1) * J1, P[g].L, J1 * 4);
        memcpy(hC + (size_t)(3 * g + 2) * J1, P[g].rB1, J1 * 4);
    }
    uint32_t *dA = reinterpret_cast<uint32_t *>(t.d), *dB = dA + (size_t)n * J1, *dC = dB + (size_t)n * J2, *dsc = dC + (size_t)3 * n * J1;
    uint32_t *rA = dsc + sc_words, *rB = rA + J1, *rC = rB + J2, *dout = rC + J1;
    CU(cudaSetDevice(c0->device));
    CU(cudaMemcpyAsync(dA, h, in_words * 4, cudaMemcpyHostToDevice, t.st));
    const GroupOps &g1 = b200msm_internal_ops(c0->curve, B200MSM_G1), &g2 = b200msm_internal_ops(c0->curve, B200MSM_G2);
    if ((rc = g1.fold_dev(c0, t.st, dA, (size_t)n, dsc, rA))) return rc;
    if ((rc = g1.to_affine_dev(c0, t.st, 1, rA, dout))) return rc;
    if ((rc = g2.fold_dev(c0, t.st, dB, (size_t)n, dsc, rB))) return rc;
    if ((rc = g2.to_affine_dev(c0, t.st, 1, rB, dout + A1))) return rc;
    if ((rc = g1.fold_dev(c0, t.st, dC, (size_t)3 * n, dsc, rC))) return rc;
    if ((rc = g1.to_affine_dev(c0, t.st, 1, rC, dout + A1 + A2))) return rc;
    CU(cudaMemcpyAsync(h, dout, out_words * 4, cudaMemcpyDeviceToHost, t.st));
    CU(cudaStreamSynchronize(t.st));
    memcpy(proof, h, out_words * 4);
    return B200MSM_OK;
}

int check_shards(b200msm_ctx *const *ctxs, b200msm_key *const *keys, int n) {
    if (!ctxs || !keys || n < 1 || !ctxs[0]) return B200MSM_ERR_ARG;
    for (int g = 0; g < n; ++g) {
        if (!ctxs[g] || !keys[g]) return fail(ctxs[0], B200MSM_ERR_ARG, "null context or key for shard %d", g);
        if (keys[g]->shard != g || keys[g]->nshards != n || keys[g]->d != keys[0]->d || keys[g]->m != keys[0]->m || ctxs[g]->curve != ctxs[0]->curve)
            return fail(ctxs[0], B200MSM_ERR_ARG, "key %d is not shard %d of %d of the same proving key", g, g, n);
    }
    return B200MSM_OK;
}
}  // namespace

// tests: the split b200msm_prove would choose for shard `shard` of `nshards` of a key with m variables and degree d (base sets
// with the default table budget), from sizes alone -- pure host arithmetic.  ns[0] = modelled time of the five MSMs one after
// the other, ns[1] = side by side.  Returns 1 when the lanes are split (sms[q] > 0), 0 when not.
extern "C" int b200msm_internal_lane_split_model(int curve, size_t d, size_t m, int shard, int nshards, double fft_delay_ns, int sm_count, int mode, int sms[5],
                                                 double ns[2]) {
    const size_t count[5] = {m + 1, m + 1, m + 1, m >= 1 ? m - 1 : 0, d};
    const int deg[5] = {1, 1, curve == B200MSM_MNT4753 ? 2 : 3, 1, 1};
    MsmCost cost[5];
    for (int q = 0; q < 5; ++q) {
        size_t lo, cnt;
        if (b200msm_shard_range(count[q], shard, nshards, &lo, &cnt)) return -1;
        cost[q] = cnt ? model_cost(cnt, deg[q], choose_cfg(cnt, deg[q], 0, size_t(32) << 30, true)) : MsmCost{0.0, 0.0};
    }
    return lane_split_compute(cost, shard == 0, d, fft_delay_ns, sm_count, mode, sms, ns) ? 1 : 0;
}

// tests: the parallel file reader on its own (no GPU involved)
extern "C" int b200msm_internal_read_range(const char *path, size_t off, size_t len, void *dst) {
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return -1;
    const bool ok = read_range(fd, static_cast<char *>(dst), off, len);
    close(fd);
    return ok ? 0 : 1;
}

extern "C" {

int b200msm_prove_sharded(b200msm_ctx *const *ctxs, b200msm_key *const *keys, int n, const void *input_image, size_t bytes, uint8_t *proof) {
    int rc = check_shards(ctxs, keys, n);
    if (rc) return rc;
    if (!input_image || !proof) return fail(ctxs[0], B200MSM_ERR_ARG, "null pointer");
    const size_t d = keys[0]->d, m = keys[0]->m;
    if (bytes != b200msm_input_bytes(keys[0])) return fail(ctxs[0], B200MSM_ERR_ARG, "input image of %zu bytes, expected %zu", bytes, b200msm_input_bytes(keys[0]));
    // layout of <curve>-input (main.cpp:35-85): w[m+1], ca[d+1], cb[d+1], cc[d+1], r -- Fr, Montgomery limbs
    const uint64_t *w = static_cast<const uint64_t *>(input_image);
    const uint64_t *ca = w + (m + 1) * 12, *cb = ca + (d + 1) * 12, *cc = cb + (d + 1) * 12, *r = cc + (d + 1) * 12;
    std::vector<Partials> P((size_t)n);
    for (int g = 0; g < n && !rc; ++g) rc = prove_begin(ctxs[g], keys[g], w, r, 0.0, P[g]);
    if (rc) { prove_drain(ctxs, n); return rc; }
    return prove_finish(ctxs, keys, n, ca, cb, cc, P.data(), proof);
}

// The same straight from the reference's <curve>-input FILE: r (the last 96 bytes) and the witness are read first and
// the witness MSMs start on every shard; the three coefficient vectors of the H polynomial (three quarters of the
// file) are read while the GPUs work.  `buffer` is host scratch of b200msm_input_bytes() bytes (pinned for full-rate
// uploads: b200msm_pinned_alloc); it holds the file image afterwards.
int b200msm_prove_sharded_file(b200msm_ctx *const *ctxs, b200msm_key *const *keys, int n, const char *input_path, void *buffer, uint8_t *proof) {
    int rc = check_shards(ctxs, keys, n);
    if (rc) return rc;
    b200msm_ctx *ctx = ctxs[0];
    if (!input_path || !buffer || !proof) return fail(ctx, B200MSM_ERR_ARG, "null pointer");
    const size_t d = keys[0]->d, m = keys[0]->m, bytes = b200msm_input_bytes(keys[0]);
    const int fd = open(input_path, O_RDONLY);
    if (fd < 0) return fail(ctx, B200MSM_ERR_ARG, "cannot open %s", input_path);
    char *img = static_cast<char *>(buffer);
    const size_t w_bytes = (m + 1) * 96, h_bytes = 3 * (d + 1) * 96;
    struct stat st;
    bool ok = fstat(fd, &st) == 0 && (size_t)st.st_size == bytes;
    ok = ok && read_range(fd, img + bytes - 96, bytes - 96, 96) && read_range(fd, img, 0, w_bytes);
    if (!ok) { close(fd); return fail(ctx, B200MSM_ERR_ARG, "%s is not an input file of %zu bytes for this key", input_path, bytes); }
    uint64_t *w = reinterpret_cast<uint64_t *>(img);
    const uint64_t *ca = w + (m + 1) * 12, *cb = ca + (d + 1) * 12, *cc = cb + (d + 1) * 12, *r = cc + (d + 1) * 12;
    std::vector<Partials> P((size_t)n);
    const double read_ns = double(h_bytes) / READ_BYTES_PER_NS;           // the FFTs cannot start before their input is here
    for (int g = 0; g < n && !rc; ++g) rc = prove_begin(ctxs[g], keys[g], w, r, read_ns, P[g]);
    // each coefficient vector is uploaded and transformed on shard 0's FFT stream while the next one is being read
    const uint64_t *vec[3] = {ca, cb, cc};
    for (int v = 0; v < 3 && !rc; ++v) {
        const size_t off = w_bytes + (size_t)v * (h_bytes / 3);
        if (!read_range(fd, img + off, off, h_bytes / 3)) rc = fail(ctx, B200MSM_ERR_ARG, "short read of %s", input_path);
        else rc = b200msm_internal_compute_h_stage(ctx, d, v, vec[v]);
    }
    close(fd);
    if (rc) { prove_drain(ctxs, n); return rc; }
    return prove_finish(ctxs, keys, n, nullptr, nullptr, nullptr, P.data(), proof);
}

int b200msm_prove(b200msm_ctx *ctx, const b200msm_key *key, const void *input_image, size_t bytes, uint8_t *proof) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (!key) return fail(ctx, B200MSM_ERR_ARG, "null pointer");
    b200msm_key *k = const_cast<b200msm_key *>(key);
    return b200msm_prove_sharded(&ctx, &k, 1, input_image, bytes, proof);
}

int b200msm_prove_file(b200msm_ctx *ctx, const b200msm_key *key, const char *input_path, void *buffer, uint8_t *proof) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (!key) return fail(ctx, B200MSM_ERR_ARG, "null pointer");
    b200msm_key *k = const_cast<b200msm_key *>(key);
    return b200msm_prove_sharded_file(&ctx, &k, 1, input_path, buffer, proof);
}

}  // extern "C"
