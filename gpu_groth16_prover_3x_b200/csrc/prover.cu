// The whole `compute` step of the reference's GPU prover behind the C ABI: proving key resident in HBM, one
// call per proof.  b200msm_prove is run_prover of cuda_prover_piecewise.cu:96-230 with
//   * the five multi-exponentiations A, B1, B2, L, H on the engine (A and H ran on the CPU in the reference),
//   * compute_H on the device (b200msm_compute_h),
//   * the proof assembly  C = Ht + Lt + r * Bt1  and the three affine normalisations on the device
//     (the reference: libff on the host, lines 198-204),
// writing the proof bytes of groth16_output_write (A || B || C, affine, infinity as zeros;
// libsnark/serialization.hpp:43-67).  It only composes the other entry points of include/b200_msm.h.
#include <cstdio>
#include <cstring>
#include <vector>

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
};

int b200msm_internal_fr_scale(b200msm_ctx *ctx, size_t n, const uint32_t *in_dev, const uint32_t *k_dev, uint32_t *out_dev, cudaStream_t st);
extern "C" int b200msm_internal_reserve(b200msm_ctx *ctx, int lane, int slot, size_t n);
int b200msm_internal_fft_prepare(b200msm_ctx *ctx, size_t d);
const GroupOps &b200msm_internal_ops(int curve, int group);

namespace {
inline int g2_deg(const b200msm_ctx *ctx) { return ctx->curve == B200MSM_MNT4753 ? 2 : 3; }
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
              cudaEventCreateWithFlags(&key->ready, cudaEventDisableTiming) == cudaSuccess;
    if (ok && shard > 0) ok = cudaMalloc(&key->h_dev, (key->cnt[4] + 1) * 96) == cudaSuccess;
    if (!ok) { b200msm_key_free(ctx, key); return fail(ctx, B200MSM_ERR_OOM, "cannot allocate the witness buffers"); }
    // warm everything the first proof would otherwise pay for: lane arenas (A and H share lane 0), domain tables
    int rc = b200msm_internal_reserve(ctx, 0, key->slot[0], key->cnt[0]);
    if (!rc) rc = b200msm_internal_reserve(ctx, 0, key->slot[4], key->cnt[4]);
    if (!rc) rc = b200msm_internal_reserve(ctx, 1, key->slot[1], key->cnt[1]);
    if (!rc) rc = b200msm_internal_reserve(ctx, 2, key->slot[2], key->cnt[2]);
    if (!rc) rc = b200msm_internal_reserve(ctx, 3, key->slot[3], key->cnt[3]);
    if (!rc && shard == 0) rc = b200msm_internal_fft_prepare(ctx, d);
    if (rc) { b200msm_key_free(ctx, key); return rc; }
    *out = key;
    return B200MSM_OK;
}

int b200msm_key_load(b200msm_ctx *ctx, const void *params_image, size_t bytes, b200msm_key **out) {
    return b200msm_key_load_shard(ctx, params_image, bytes, 0, 1, out);
}

int b200msm_key_load_file(b200msm_ctx *ctx, const char *path, b200msm_key **out) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (!path || !out) return fail(ctx, B200MSM_ERR_ARG, "bad argument");
    FILE *f = fopen(path, "rb");
    if (!f) return fail(ctx, B200MSM_ERR_ARG, "cannot open %s", path);
    fseek(f, 0, SEEK_END);
    const long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> buf(n > 0 ? (size_t)n : 0);
    const bool ok = n > 0 && fread(buf.data(), 1, (size_t)n, f) == (size_t)n;
    fclose(f);
    if (!ok) return fail(ctx, B200MSM_ERR_ARG, "cannot read %s", path);
    return b200msm_key_load(ctx, buf.data(), buf.size(), out);
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

// First half of a proof on one shard: its witness slice and r on its device, its four witness MSMs in flight.
int prove_begin(b200msm_ctx *ctx, const b200msm_key *key, const uint64_t *w, const uint64_t *r, Partials &P) {
    int rc;
    // The witness crosses PCIe once; the B1 query runs on r * w so that its result is the r * Bt1 term of C directly
    // (sum (r w_i) B1_i = r * sum w_i B1_i: the same group element, no 753-step scalar multiplication afterwards).
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(key->w_dev, w + key->w_lo * 12, key->w_cnt * 96, cudaMemcpyHostToDevice, key->stream));
    CU(cudaMemcpyAsync(key->r_dev, r, 96, cudaMemcpyHostToDevice, key->stream));
    const uint32_t *wa = key->w_dev + (key->lo[0] - key->w_lo) * 24;          // first witness element of the A / B1 / B2 slice
    const uint32_t *wl = key->w_dev + (2 + key->lo[3] - key->w_lo) * 24;      // and of the L slice: w[2..m] (:167)
    if ((rc = b200msm_internal_fr_scale(ctx, key->cnt[0], wa, key->r_dev, key->rw_dev, key->stream))) return rc;
    CU(cudaEventRecord(key->ready, key->stream));
    for (int l = 0; l < 4; ++l) CU(cudaStreamWaitEvent(ctx->lanes[l].stream, key->ready, 0));
    // the four witness MSMs in flight together (cuda_prover_piecewise.cu:162-167)
    if ((rc = b200msm_msm_async(ctx, 0, key->slot[0], 0, reinterpret_cast<const uint64_t *>(wa), key->cnt[0], P.A))) return rc;
    if ((rc = b200msm_msm_async(ctx, 1, key->slot[1], 0, reinterpret_cast<const uint64_t *>(key->rw_dev), key->cnt[1], P.rB1))) return rc;
    if ((rc = b200msm_msm_async(ctx, 2, key->slot[2], 0, reinterpret_cast<const uint64_t *>(wa), key->cnt[2], P.B2))) return rc;
    if ((rc = b200msm_msm_async(ctx, 3, key->slot[3], 0, reinterpret_cast<const uint64_t *>(wl), key->cnt[3], P.L))) return rc;
    return B200MSM_OK;
}
void prove_drain(b200msm_ctx *const *ctxs, int n) {
    for (int g = 0; g < n; ++g)
        for (int l = 0; l < 4; ++l) b200msm_wait(ctxs[g], l);
}

// Second half: the H polynomial on shard 0's GPU beside the MSMs, its coefficients handed to the other shards over
// NVLink, the H query, the fold of the partial points, the assembly and the proof bytes.
int prove_finish(b200msm_ctx *const *ctxs, b200msm_key *const *keys, int n, const uint64_t *ca, const uint64_t *cb, const uint64_t *cc,
                 Partials *P, uint8_t *proof) {
    b200msm_ctx *c0 = ctxs[0];
    const size_t d = keys[0]->d;
    const int dg = g2_deg(c0);
    const uint64_t *h_dev = nullptr;
    int rc = b200msm_compute_h(c0, d, ca, cb, cc, nullptr, &h_dev);
    for (int g = 0; g < n && !rc; ++g) {
        rc = b200msm_wait(ctxs[g], 0);                    // lane 0 carries A, then H
        if (rc) break;
        const uint64_t *hs = h_dev + keys[g]->lo[4] * 12;
        if (g > 0) {
            // on lane 0's stream of the receiving GPU: ordered before the H query enqueued on the same stream below,
            // and no device-wide synchronisation while lanes 1-3 are still working (the source is complete:
            // b200msm_compute_h is synchronous)
            if (cudaMemcpyPeerAsync(keys[g]->h_dev, ctxs[g]->device, hs, c0->device, keys[g]->cnt[4] * 96, ctxs[g]->lanes[0].stream) != cudaSuccess) {
                rc = fail(c0, B200MSM_ERR_CUDA, "peer copy of the H coefficients to device %d failed", ctxs[g]->device);
                break;
            }
            hs = reinterpret_cast<const uint64_t *>(keys[g]->h_dev);
        }
        rc = b200msm_msm_async(ctxs[g], 0, keys[g]->slot[4], 0, hs, keys[g]->cnt[4], P[g].H);
    }
    for (int g = 0; g < n; ++g)
        for (int l = 0; l < 4; ++l) { const int rcw = b200msm_wait(ctxs[g], l); if (!rc) rc = rcw; }
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
    for (int g = 0; g < n && !rc; ++g) rc = prove_begin(ctxs[g], keys[g], w, r, P[g]);
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
    FILE *f = fopen(input_path, "rb");
    if (!f) return fail(ctx, B200MSM_ERR_ARG, "cannot open %s", input_path);
    char *img = static_cast<char *>(buffer);
    const size_t w_bytes = (m + 1) * 96, h_bytes = 3 * (d + 1) * 96;
    bool ok = fseek(f, 0, SEEK_END) == 0 && (size_t)ftell(f) == bytes;
    ok = ok && fseek(f, (long)(bytes - 96), SEEK_SET) == 0 && fread(img + bytes - 96, 1, 96, f) == 96;
    ok = ok && fseek(f, 0, SEEK_SET) == 0 && fread(img, 1, w_bytes, f) == w_bytes;
    if (!ok) { fclose(f); return fail(ctx, B200MSM_ERR_ARG, "%s is not an input file of %zu bytes for this key", input_path, bytes); }
    uint64_t *w = reinterpret_cast<uint64_t *>(img);
    const uint64_t *ca = w + (m + 1) * 12, *cb = ca + (d + 1) * 12, *cc = cb + (d + 1) * 12, *r = cc + (d + 1) * 12;
    std::vector<Partials> P((size_t)n);
    for (int g = 0; g < n && !rc; ++g) rc = prove_begin(ctxs[g], keys[g], w, r, P[g]);
    if (!rc && fread(img + w_bytes, 1, h_bytes, f) != h_bytes) rc = fail(ctx, B200MSM_ERR_ARG, "short read of %s", input_path);
    fclose(f);
    if (rc) { prove_drain(ctxs, n); return rc; }
    return prove_finish(ctxs, keys, n, ca, cb, cc, P.data(), proof);
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
