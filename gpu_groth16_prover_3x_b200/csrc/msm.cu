// Host side of the engine: the C ABI of include/b200_msm.h.  One context per (curve, GPU); five
// internal streams ("lanes"), one per query of a proof, so that its multiexps can be in flight
// together, as the reference does with one stream per MSM (cuda_prover_piecewise.cu:162-167).  No CPU
// fallback: every failure is reported through the return code and b200msm_last_error().
// The kernels are instantiated per group in inst_*.cu and reached through GroupOps (group_ops.cuh).
#include <cstdlib>

#include "host_ctx.cuh"

extern const GroupOps b200msm_ops_mnt4g1, b200msm_ops_mnt4g2, b200msm_ops_mnt6g1, b200msm_ops_mnt6g2;

namespace {

const GroupOps &ops_for(int curve, int group) {
    if (curve == B200MSM_MNT4753) return group == B200MSM_G1 ? b200msm_ops_mnt4g1 : b200msm_ops_mnt4g2;
    return group == B200MSM_G1 ? b200msm_ops_mnt6g1 : b200msm_ops_mnt6g2;
}

int dispatch_msm(b200msm_ctx *ctx, int lane, int slot, size_t offset, const uint64_t *scalars, size_t n, uint64_t *out) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (lane < 0 || lane >= NLANES) return fail(ctx, B200MSM_ERR_ARG, "lane %d out of range", lane);
    if (slot < 0 || slot >= (int)ctx->sets.size() || !ctx->sets[slot].used) return fail(ctx, B200MSM_ERR_ARG, "bad base slot %d", slot);
    const BaseSet &bs = ctx->sets[slot];
    if (offset > bs.n || n > bs.n - offset) return fail(ctx, B200MSM_ERR_ARG, "range [%zu, %zu) exceeds base set of %zu points", offset, offset + n, bs.n);
    if (n >= (size_t(1) << 31)) return fail(ctx, B200MSM_ERR_ARG, "n = %zu too large (max 2^31 - 1 per call)", n);
    if (!out || (n && !scalars)) return fail(ctx, B200MSM_ERR_ARG, "null pointer");
    if (ctx->lanes[lane].pending) return fail(ctx, B200MSM_ERR_ARG, "lane %d still has an un-waited MSM", lane);
    CU(cudaSetDevice(ctx->device));
    return ops_for(ctx->curve, bs.group).enqueue(ctx, lane, bs, offset, scalars, n, out);
}

int build_tables_any(b200msm_ctx *ctx, BaseSet &bs) {
    return ops_for(ctx->curve, bs.group).build_tables(ctx, bs);
}

// decide the table layout of a new base set of n points and allocate it (table 0 still to be filled)
int alloc_base_set(b200msm_ctx *ctx, int group, size_t n, bool tables, BaseSet &bs) {
    const int deg = degree_of(ctx->curve, group);
    bs = BaseSet();
    bs.used = true;
    bs.group = group;
    bs.n = n;
    const TabCfg cfg = choose_cfg(n ? n : 1, deg, ctx->c_override, ctx->table_budget, tables);
    if (cfg.NT > 1) { bs.c_tab = cfg.c; bs.NT = cfg.NT; bs.G = cfg.G; bs.Wd = cfg.Wd; bs.glv = cfg.glv; }
    const size_t bytes = (size_t)bs.NT * n * 2 * deg * NLIMB * 4;
    CU(cudaMalloc(&bs.pts, bytes ? bytes : 256));
    cudaError_t e = cudaMalloc(&bs.inf, n ? n : 256);
    if (e != cudaSuccess) { cudaFree(bs.pts); return fail(ctx, B200MSM_ERR_OOM, "cudaMalloc: %s", cudaGetErrorString(e)); }
    return B200MSM_OK;
}

int register_base_set(b200msm_ctx *ctx, const BaseSet &bs) {
    int id = -1;
    for (size_t i = 0; i < ctx->sets.size(); ++i) if (!ctx->sets[i].used) { id = (int)i; break; }
    if (id < 0) { ctx->sets.push_back(bs); id = (int)ctx->sets.size() - 1; } else ctx->sets[id] = bs;
    return id;
}

int upload_impl(b200msm_ctx *ctx, int group, const uint64_t *affine, size_t n, int *slot, bool tables) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (!slot || (group != B200MSM_G1 && group != B200MSM_G2) || (n && !affine)) return fail(ctx, B200MSM_ERR_ARG, "bad argument");
    if (n >= (size_t(1) << 31)) return fail(ctx, B200MSM_ERR_ARG, "n = %zu too large", n);
    CU(cudaSetDevice(ctx->device));
    const int deg = degree_of(ctx->curve, group);
    BaseSet bs;
    int rc = alloc_base_set(ctx, group, n, tables, bs);
    if (rc) return rc;
    if (n) {
        cudaError_t e = cudaMemcpy(bs.pts, affine, n * 2 * deg * NLIMB * 4, cudaMemcpyDefault);
        if (e == cudaSuccess) {
            const unsigned grid = (unsigned)((n + 255) / 256);
            if (deg == 1) k_flag_inf<1><<<grid, 256>>>(bs.pts, (uint32_t)n, bs.inf);
            else if (deg == 2) k_flag_inf<2><<<grid, 256>>>(bs.pts, (uint32_t)n, bs.inf);
            else k_flag_inf<3><<<grid, 256>>>(bs.pts, (uint32_t)n, bs.inf);
            e = cudaDeviceSynchronize();
        }
        if (e != cudaSuccess) { cudaFree(bs.pts); cudaFree(bs.inf); return fail(ctx, B200MSM_ERR_CUDA, "base upload: %s", cudaGetErrorString(e)); }
        if ((rc = build_tables_any(ctx, bs))) { cudaFree(bs.pts); cudaFree(bs.inf); return rc; }
    }
    *slot = register_base_set(ctx, bs);
    return B200MSM_OK;
}

// ---- microbenchmarks ------------------------------------------------------------------------------
constexpr int MB_CHAINS = 16;
// 16 independent chains per thread.  Each step is one 32x32->64 product (IMAD.WIDE.U32) folded back into
// its chain by one LOP3, so nothing is loop-invariant and ptxas cannot split the instruction (a
// mad.wide.u32 with a 64-bit addend is rewritten by ptxas into IMAD.WIDE.U32 + IADD3 + IADD3.X, and an
// earlier version of this probe counted that wrongly: tools/pipe_probe.cu, profiles/r01_pipe_probe.txt).
__global__ void __launch_bounds__(256) k_mb_wide(uint32_t *out, int iters) {
    uint32_t acc[MB_CHAINS];
    const uint32_t b = 0x9e3779b9u + blockIdx.x;
    for (int j = 0; j < MB_CHAINS; ++j) acc[j] = (threadIdx.x + 1) * (j + 3);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < MB_CHAINS; ++j) {
            uint64_t p;
            asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(acc[j]), "r"(b));
            acc[j] = (uint32_t)p ^ (uint32_t)(p >> 32) ^ j;
        }
    }
    uint32_t s = 0;
    for (int j = 0; j < MB_CHAINS; ++j) s ^= acc[j];
    if (s == 0x12345678u) out[0] = 1;
}
__global__ void __launch_bounds__(256) k_mb_lo(uint32_t *out, int iters) {
    uint32_t acc[MB_CHAINS];
    const uint32_t b = 0x9e3779b9u + blockIdx.x;
    for (int j = 0; j < MB_CHAINS; ++j) acc[j] = (threadIdx.x + 1) * (j + 3);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < MB_CHAINS; ++j)
            asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(acc[j]) : "r"(acc[(j + 1) % MB_CHAINS]), "r"(b));
    }
    uint32_t s = 0;
    for (int j = 0; j < MB_CHAINS; ++j) s ^= acc[j];
    if (s == 0x12345678u) out[0] = 1;
}
// the engine's own register-resident Montgomery product, iterated
template <class M>
__global__ void __launch_bounds__(128) k_mb_fqmul(uint32_t *out, int iters) {
    fq_t x, y;
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) { x[i] = M::R1(i) ^ (threadIdx.x * 7u + i); y[i] = M::R2(i) ^ (blockIdx.x + i); }
    x[NLIMB - 1] &= 0xffffu;
    y[NLIMB - 1] &= 0xffffu;
    for (int it = 0; it < iters; ++it) {
        fq_mul<M>(x, x, y);
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) s ^= x[i];
    if (s == 0x12345678u) out[0] = 1;
}

// latency of ONE field inversion executed by a single thread (the situation inside k_batch_add's tile_inverse):
// a dependent chain of `iters` inversions on lane 0 of one warp.  MODE 0: fq_inv (fast path + fallback),
// 1: plain binary gcd, 2: fast path only (out[1] counts failures of its final check).
template <class M, int MODE>
__global__ void __launch_bounds__(32) k_mb_inv(uint32_t *out, int iters) {
    if (threadIdx.x != 0) return;
    fq_t x, r;
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) x[i] = M::R2(i) ^ (0x01010101u * (uint32_t)i);
    x[NLIMB - 1] &= 0xffffu;
    uint32_t fails = 0;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) fq_inv<M>(r, x);
        else if (MODE == 1) { if (!fq_inv_plain<M>(r, x)) ++fails; }
        else { if (!fq_inv_plain_fast<M>(r, x)) ++fails; }
#pragma unroll
        for (int i = 0; i < NLIMB; ++i) x[i] = r[i] ^ (uint32_t)(it + 1);
        x[NLIMB - 1] &= 0xffffu;
        x[0] |= 1u;
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) s ^= x[i];
    out[0] = s;
    out[1] = fails;
}

}  // namespace

// the per-group routines, for the proof assembly in prover.cu
const GroupOps &b200msm_internal_ops(int curve, int group) { return ops_for(curve, group); }

// ================================================================================================
extern "C" {

int b200msm_create(int curve, int device, b200msm_ctx **out) {
    if (!out || (curve != B200MSM_MNT4753 && curve != B200MSM_MNT6753)) return B200MSM_ERR_ARG;
    *out = nullptr;
    // A context runs up to eight streams at once (five lanes, the FFTs, the witness upload, the proof tail) plus the
    // copy streams of chunked scalar uploads.  The driver maps streams onto 8 hardware queues by default and streams
    // that share a queue wait for each other: the H query, enqueued on its own lane the moment its FFTs were done,
    // started only when the A query's kernels had drained (measured with B200MSM_TRACE).  More queues, unless the
    // caller chose a number; it takes effect when this is the process's first CUDA call (b200_prove), and is harmless
    // otherwise.
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return B200MSM_ERR_CUDA;
    b200msm_ctx *ctx = new b200msm_ctx;
    ctx->curve = curve;
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) {
        delete ctx;
        return B200MSM_ERR_CUDA;  // sm_100a code only: there is nothing to run elsewhere
    }
    ctx->sm_count = prop.multiProcessorCount;
    for (int i = 0; i < NLANES; ++i) {
        Lane &ln = ctx->lanes[i];
        bool ok = cudaStreamCreateWithFlags(&ln.own_stream, cudaStreamNonBlocking) == cudaSuccess;
        ln.stream = ln.own_stream;
        for (int e = 0; e < NEVENTS && ok; ++e) ok = cudaEventCreate(&ln.ev[e]) == cudaSuccess;
        ok = ok && cudaMallocHost(&ln.h_result, 3 * 3 * NLIMB * 4) == cudaSuccess;
        ok = ok && cudaMallocHost(&ln.h_ctl, BA_CTL_WORDS * 4) == cudaSuccess;
        if (!ok) { b200msm_destroy(ctx); return B200MSM_ERR_CUDA; }
    }
    *out = ctx;
    return B200MSM_OK;
}

void b200msm_destroy(b200msm_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    b200msm_fft_release(ctx);
    for (int i = 0; i < NLANES; ++i) {
        Lane &ln = ctx->lanes[i];
        if (ln.stream) cudaStreamSynchronize(ln.stream);
        if (ln.arena) cudaFree(ln.arena);
        if (ln.h_result) cudaFreeHost(ln.h_result);
        if (ln.h_ctl) cudaFreeHost(ln.h_ctl);
        for (int e = 0; e < NEVENTS; ++e) if (ln.ev[e]) cudaEventDestroy(ln.ev[e]);
        for (int e = 0; e <= NCOPY; ++e) if (ln.ev_copy[e]) cudaEventDestroy(ln.ev_copy[e]);
        if (ln.copy_stream) cudaStreamDestroy(ln.copy_stream);
        if (ln.own_stream) cudaStreamDestroy(ln.own_stream);
    }
    for (auto &s : ctx->sets) if (s.used) { cudaFree(s.pts); cudaFree(s.inf); }
    if (ctx->tail.st) { cudaStreamSynchronize(ctx->tail.st); cudaStreamDestroy(ctx->tail.st); }
    if (ctx->tail.d) cudaFree(ctx->tail.d);
    if (ctx->tail.h) cudaFreeHost(ctx->tail.h);
    delete ctx;
}

const char *b200msm_last_error(const b200msm_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int b200msm_bases_upload(b200msm_ctx *ctx, int group, const uint64_t *affine, size_t n, int *slot) {
    return upload_impl(ctx, group, affine, n, slot, true);
}

int b200msm_bases_free(b200msm_ctx *ctx, int slot) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (slot < 0 || slot >= (int)ctx->sets.size() || !ctx->sets[slot].used) return fail(ctx, B200MSM_ERR_ARG, "bad base slot %d", slot);
    CU(cudaSetDevice(ctx->device));
    CU(cudaDeviceSynchronize());
    cudaFree(ctx->sets[slot].pts);
    cudaFree(ctx->sets[slot].inf);
    ctx->sets[slot] = BaseSet();
    return B200MSM_OK;
}

int b200msm_msm_async(b200msm_ctx *ctx, int lane, int slot, size_t offset, const uint64_t *scalars_mont, size_t n, uint64_t *out_xyz) {
    return dispatch_msm(ctx, lane, slot, offset, scalars_mont, n, out_xyz);
}

int b200msm_wait(b200msm_ctx *ctx, int lane) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (lane < 0 || lane >= NLANES) return fail(ctx, B200MSM_ERR_ARG, "lane %d out of range", lane);
    Lane &ln = ctx->lanes[lane];
    if (!ln.pending) return fail(ctx, B200MSM_ERR_ARG, "nothing pending on lane %d", lane);
    ln.pending = false;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ln.stream));
    if (ln.timed) {
        for (int i = 0; i < 5; ++i) CU(cudaEventElapsedTime(&ln.ms[i + 1], ln.ev[i], ln.ev[i + 1]));
        CU(cudaEventElapsedTime(&ln.ms[0], ln.ev[0], ln.ev[5]));
    }
    memcpy(ln.user_out, ln.h_result, ln.out_words * 8);
    return B200MSM_OK;
}

int b200msm_msm(b200msm_ctx *ctx, int slot, size_t offset, const uint64_t *scalars_mont, size_t n, uint64_t *out_xyz) {
    int rc = dispatch_msm(ctx, 0, slot, offset, scalars_mont, n, out_xyz);
    if (rc) return rc;
    return b200msm_wait(ctx, 0);
}

int b200msm_ec_reduce(b200msm_ctx *ctx, int group, const uint64_t *bases_affine, const uint64_t *scalars_mont, size_t n, uint64_t *out_xyz) {
    int slot = -1;
    int rc = upload_impl(ctx, group, bases_affine, n, &slot, false);  // one-shot bases: no window tables
    if (rc) return rc;
    rc = b200msm_msm(ctx, slot, 0, scalars_mont, n, out_xyz);
    std::string keep = ctx->err;
    b200msm_bases_free(ctx, slot);
    if (rc) ctx->err = keep;
    return rc;
}

int b200msm_fold(b200msm_ctx *ctx, int group, const uint64_t *partials_xyz, size_t n, uint64_t *out_xyz) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (!out_xyz || n == 0 || !partials_xyz || n > (1u << 20) || (group != B200MSM_G1 && group != B200MSM_G2))
        return fail(ctx, B200MSM_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    return ops_for(ctx->curve, group).fold(ctx, partials_xyz, n, out_xyz);
}

int b200msm_shard_range(size_t n, int shard, int nshards, size_t *offset, size_t *length) {
    if (nshards < 1 || shard < 0 || shard >= nshards || !offset || !length) return B200MSM_ERR_ARG;
    // 128-bit product: n * nshards cannot overflow for any n a base set can hold, but the ABI takes size_t
    const unsigned __int128 N = n;
    const size_t lo = (size_t)(N * (unsigned)shard / (unsigned)nshards), hi = (size_t)(N * (unsigned)(shard + 1) / (unsigned)nshards);
    *offset = lo;
    *length = hi - lo;
    return B200MSM_OK;
}

int b200msm_to_affine(b200msm_ctx *ctx, int group, size_t n, const uint64_t *xyz, uint64_t *out_affine) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (!xyz || !out_affine || n == 0 || n > (1u << 24) || (group != B200MSM_G1 && group != B200MSM_G2))
        return fail(ctx, B200MSM_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    return ops_for(ctx->curve, group).to_affine(ctx, n, xyz, out_affine);
}

// internal (prover.cu): pre-size lane `lane` for MSMs of n points over base set `slot`
int b200msm_internal_reserve(b200msm_ctx *ctx, int lane, int slot, size_t n) {
    if (!ctx || lane < 0 || lane >= NLANES || slot < 0 || slot >= (int)ctx->sets.size() || !ctx->sets[slot].used) return B200MSM_ERR_ARG;
    CU(cudaSetDevice(ctx->device));
    const BaseSet &bs = ctx->sets[slot];
    return ops_for(ctx->curve, bs.group).reserve(ctx, lane, bs, n);
}

int b200msm_scalar_mul(b200msm_ctx *ctx, int group, const uint64_t *affine, const uint64_t *k_mont, uint64_t *out_xyz) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (!affine || !k_mont || !out_xyz || (group != B200MSM_G1 && group != B200MSM_G2)) return fail(ctx, B200MSM_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    return ops_for(ctx->curve, group).scalar_mul(ctx, affine, k_mont, out_xyz);
}

int b200msm_bases_synthetic(b200msm_ctx *ctx, int group, size_t n, const uint64_t *k_p0_mont, const uint64_t *k_q_mont, int *slot) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (!slot || !k_p0_mont || !k_q_mont || n == 0 || n >= (size_t(1) << 31) || (group != B200MSM_G1 && group != B200MSM_G2))
        return fail(ctx, B200MSM_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    BaseSet bs;
    int rc = alloc_base_set(ctx, group, n, true, bs);
    if (rc) return rc;
    rc = ops_for(ctx->curve, group).synthetic(ctx, n, k_p0_mont, k_q_mont, bs);
    if (!rc) rc = build_tables_any(ctx, bs);
    if (rc) { cudaFree(bs.pts); cudaFree(bs.inf); return rc; }
    *slot = register_base_set(ctx, bs);
    return B200MSM_OK;
}

int b200msm_bases_download(b200msm_ctx *ctx, int slot, size_t offset, size_t n, uint64_t *out_affine) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (slot < 0 || slot >= (int)ctx->sets.size() || !ctx->sets[slot].used) return fail(ctx, B200MSM_ERR_ARG, "bad base slot %d", slot);
    const BaseSet &bs = ctx->sets[slot];
    if (!out_affine || offset > bs.n || n > bs.n - offset) return fail(ctx, B200MSM_ERR_ARG, "bad range");
    CU(cudaSetDevice(ctx->device));
    const size_t pw = (size_t)2 * degree_of(ctx->curve, bs.group) * NLIMB;
    CU(cudaMemcpy(out_affine, bs.pts + offset * pw, n * pw * 4, cudaMemcpyDefault));
    return B200MSM_OK;
}

int b200msm_set_stream(b200msm_ctx *ctx, int lane, void *cuda_stream) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (lane < 0 || lane >= NLANES) return fail(ctx, B200MSM_ERR_ARG, "lane %d out of range", lane);
    Lane &ln = ctx->lanes[lane];
    if (ln.pending) return fail(ctx, B200MSM_ERR_ARG, "lane %d still has an un-waited MSM", lane);
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ln.stream));
    ln.stream = cuda_stream ? (cudaStream_t)cuda_stream : ln.own_stream;
    return B200MSM_OK;
}

int b200msm_set_lane_sms(b200msm_ctx *ctx, int lane, int sms) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (lane < 0 || lane >= NLANES) return fail(ctx, B200MSM_ERR_ARG, "lane %d out of range", lane);
    ctx->lane_sms[lane] = (sms > 0 && sms < ctx->sm_count) ? sms : 0;
    return B200MSM_OK;
}

int b200msm_set_table_budget(b200msm_ctx *ctx, size_t max_bytes_per_set) {
    if (!ctx) return B200MSM_ERR_ARG;
    ctx->table_budget = max_bytes_per_set;
    return B200MSM_OK;
}

int b200msm_bases_info(b200msm_ctx *ctx, int slot, uint64_t info[6]) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (slot < 0 || slot >= (int)ctx->sets.size() || !ctx->sets[slot].used || !info) return fail(ctx, B200MSM_ERR_ARG, "bad base slot %d", slot);
    const BaseSet &bs = ctx->sets[slot];
    info[0] = bs.n;
    info[1] = (uint64_t)bs.c_tab;
    info[2] = (uint64_t)bs.NT;
    info[3] = (uint64_t)bs.G;
    info[4] = (uint64_t)bs.NT * bs.n * 2 * degree_of(ctx->curve, bs.group) * NLIMB * 4;
    info[5] = (uint64_t)(bs.build_ms * 1000.0f);
    return B200MSM_OK;
}

int b200msm_set_window_bits(b200msm_ctx *ctx, int c) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (c != 0 && (c < 2 || c > 22)) return fail(ctx, B200MSM_ERR_ARG, "window bits %d outside [2, 22]", c);
    ctx->c_override = c;
    return B200MSM_OK;
}

int b200msm_last_rounds(b200msm_ctx *ctx, int lane, uint64_t info[4], uint32_t *pairs_per_round, size_t max_rounds) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (lane < 0 || lane >= NLANES || !info) return fail(ctx, B200MSM_ERR_ARG, "bad argument");
    const Lane &ln = ctx->lanes[lane];
    if (ln.pending) return fail(ctx, B200MSM_ERR_ARG, "lane %d still has an un-waited MSM", lane);
    memset(info, 0, 4 * sizeof(uint64_t));
    if (!ln.ctl_valid) return B200MSM_OK;
    const uint32_t *ctl = ln.h_ctl;   // BA_CTL_WORDS: rounds | largest bucket | additions | - | pairs of round r ...
    info[0] = ctl[0];
    info[1] = ln.info[3];
    info[2] = ctl[1];
    info[3] = ctl[2];
    for (uint32_t r = 0; r < ctl[0] && r < (uint32_t)(BA_CTL_WORDS - 4); ++r)
        if (pairs_per_round && (size_t)r < max_rounds) pairs_per_round[r] = ctl[4 + r];
    return B200MSM_OK;
}

// development introspection (not declared in the public header): arena offsets of the last MSM of a lane, raw reads
int b200msm_internal_debug_layout(b200msm_ctx *ctx, int lane, uint64_t dbg[16]) {
    if (!ctx || lane < 0 || lane >= NLANES) return B200MSM_ERR_ARG;
    memcpy(dbg, ctx->lanes[lane].dbg, sizeof ctx->lanes[lane].dbg);
    return B200MSM_OK;
}
int b200msm_internal_debug_read(b200msm_ctx *ctx, int lane, uint64_t offset, uint64_t bytes, void *out) {
    if (!ctx || lane < 0 || lane >= NLANES) return B200MSM_ERR_ARG;
    Lane &ln = ctx->lanes[lane];
    if (offset + bytes > ln.arena_bytes) return fail(ctx, B200MSM_ERR_ARG, "debug read beyond the arena");
    CU(cudaSetDevice(ctx->device));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(out, ln.arena + offset, bytes, cudaMemcpyDeviceToHost));
    return B200MSM_OK;
}

int b200msm_last_timings(b200msm_ctx *ctx, int lane, float ms[6], uint64_t info[8]) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (lane < 0 || lane >= NLANES) return fail(ctx, B200MSM_ERR_ARG, "lane %d out of range", lane);
    if (ms) memcpy(ms, ctx->lanes[lane].ms, sizeof(float) * 6);
    if (info) memcpy(info, ctx->lanes[lane].info, sizeof(uint64_t) * 8);
    return B200MSM_OK;
}

int b200msm_microbench(b200msm_ctx *ctx, int kind, int iters, double *gops) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (!gops || iters <= 0 || kind < 0 || kind > 18 || kind == 3) return fail(ctx, B200MSM_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    if (kind >= 15) {  // register multiplier at 4 (kind 15), 8 (16) or 12 (17) warps per SM: does ONE warp per scheduler fill the pipe?
        if (kind > 17) return fail(ctx, B200MSM_ERR_ARG, "bad argument");
        DevBuf buf;
        CU(cudaMalloc(&buf.p, 256));
        uint32_t *dd = static_cast<uint32_t *>(buf.p);
        EventPair ev;
        CU(cudaEventCreate(&ev.e0));
        CU(cudaEventCreate(&ev.e1));
        const int nb = ctx->sm_count * (kind - 14);
        for (int rep = 0; rep < 2; ++rep) {
            CU(cudaEventRecord(ev.e0, 0));
            k_mb_fqmul<ModA><<<nb, 128>>>(dd, iters);
            CU(cudaEventRecord(ev.e1, 0));
            CU(cudaEventSynchronize(ev.e1));
        }
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, ev.e0, ev.e1));
        *gops = double(nb) * 128 * iters / (double(ms) * 1e6);
        return B200MSM_OK;
    }
    if (kind >= 7) {  // slab multiplier: kind = 7 + 4 * (group - 1) + (blocks per SM - 1), blocks per SM in 1..4
        const int group = kind < 11 ? B200MSM_G1 : B200MSM_G2, bps = (kind - 7) % 4 + 1;
        if (kind > 14) return fail(ctx, B200MSM_ERR_ARG, "bad argument");
        return ops_for(ctx->curve, group).teammul_bench(ctx, bps, iters, gops);
    }
    DevBuf buf;
    CU(cudaMalloc(&buf.p, 256));
    uint32_t *d = static_cast<uint32_t *>(buf.p);
    EventPair ev;
    CU(cudaEventCreate(&ev.e0));
    CU(cudaEventCreate(&ev.e1));
    cudaEvent_t e0 = ev.e0, e1 = ev.e1;
    const int blocks = ctx->sm_count * 8;
    double ops = 0;
    if (kind >= 4) {  // single-thread inversion latency: returns microseconds per inversion
        for (int rep = 0; rep < 2; ++rep) {
            CU(cudaEventRecord(e0, 0));
            if (kind == 4) k_mb_inv<ModA, 0><<<1, 32>>>(d, iters);
            else if (kind == 5) k_mb_inv<ModA, 1><<<1, 32>>>(d, iters);
            else k_mb_inv<ModA, 2><<<1, 32>>>(d, iters);
            CU(cudaEventRecord(e1, 0));
            CU(cudaEventSynchronize(e1));
        }
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, e0, e1));
        uint32_t h[2] = {0, 0};
        CU(cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost));
        if (kind == 6 && h[1]) return fail(ctx, B200MSM_ERR_CUDA, "fast inversion failed its final check %u times", h[1]);
        *gops = double(ms) * 1e3 / iters;
        return B200MSM_OK;
    }
    for (int rep = 0; rep < 2; ++rep) {  // first repetition is the warm-up
        CU(cudaEventRecord(e0, 0));
        if (kind == 0) { k_mb_wide<<<blocks, 256>>>(d, iters); ops = double(blocks) * 256 * iters * MB_CHAINS; }
        else if (kind == 1) { k_mb_lo<<<blocks, 256>>>(d, iters); ops = double(blocks) * 256 * iters * MB_CHAINS; }
        else {
            if (ctx->curve == B200MSM_MNT4753) k_mb_fqmul<ModA><<<blocks, 128>>>(d, iters);
            else k_mb_fqmul<ModB><<<blocks, 128>>>(d, iters);
            ops = double(blocks) * 128 * iters;
        }
        CU(cudaEventRecord(e1, 0));
        CU(cudaEventSynchronize(e1));
    }
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    CU(cudaGetLastError());
    *gops = ops / (double(ms) * 1e6);
    return B200MSM_OK;
}

int b200msm_selftest_field(b200msm_ctx *ctx, int group, int op, size_t n, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (!a || !out || n == 0 || n > (1u << 24)) return fail(ctx, B200MSM_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    return ops_for(ctx->curve, group).selftest(ctx, false, op, n, a, b, nullptr, out);
}

int b200msm_selftest_point(b200msm_ctx *ctx, int group, int op, size_t n, const uint64_t *acc, const uint64_t *q, const uint32_t *flags, uint64_t *out) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (!acc || !out || n == 0 || n > (1u << 22) || op < 0 || op > 2 || (op != 2 && !q)) return fail(ctx, B200MSM_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    return ops_for(ctx->curve, group).selftest(ctx, true, op, n, acc, q, flags, out);
}

}  // extern "C"
