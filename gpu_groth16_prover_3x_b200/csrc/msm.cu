// Host side of the engine: the C ABI of include/b200_msm.h on top of the kernels in
// msm_kernels.cuh.  One context per (curve, GPU); four internal streams ("lanes") so that the
// A, B1, B2 and L multiexps of one proof can be in flight together, as the reference does with
// one stream per MSM (cuda_prover_piecewise.cu:162-167).  No CPU fallback: every failure is
// reported through the return code and b200msm_last_error().
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200_msm.h"
#include "msm_kernels.cuh"
#include "util_kernels.cuh"

using namespace mnt753;

namespace {

constexpr int NLANES = 4;
constexpr int NEVENTS = 7;

struct BaseSet {
    bool used = false;
    int group = 0;
    size_t n = 0;
    uint32_t *pts = nullptr;
    uint8_t *inf = nullptr;
};

struct Lane {
    cudaStream_t stream = nullptr;      // stream MSMs are enqueued on
    cudaStream_t own_stream = nullptr;  // the lane's internal stream (stream == own_stream unless overridden)
    cudaEvent_t ev[NEVENTS] = {};
    char *arena = nullptr;
    size_t arena_bytes = 0;
    uint32_t *h_result = nullptr;  // pinned staging for the Jacobian result
    uint64_t *user_out = nullptr;
    size_t out_words = 0;
    bool pending = false;
    bool timed = false;
    float ms[6] = {};
    uint64_t info[5] = {};
};

}  // namespace

struct b200msm_ctx {
    int curve = 0;
    int device = 0;
    int sm_count = 0;
    int c_override = 0;
    std::vector<BaseSet> sets;
    Lane lanes[NLANES];
    std::string err = "";
};

namespace {

int fail(b200msm_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, e_ == cudaErrorMemoryAllocation ? B200MSM_ERR_OOM : B200MSM_ERR_CUDA, \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

inline int degree_of(int curve, int group) { return group == B200MSM_G1 ? 1 : (curve == B200MSM_MNT4753 ? 2 : 3); }

// ---- window-size choice ---------------------------------------------------------------------
// cost model in Fq-tower multiplications: one mixed add (11) per point and window, two full adds
// (16 each) per bucket for the running-sum reduction, and the serial window combine.
int auto_window_bits(size_t n) {
    int best = 2;
    double best_cost = 1e300;
    for (int c = 2; c <= 18; ++c) {
        const double W = (MNT753_NUM_BITS + 1 + c - 1) / c;
        const double NB = double(1u << (c - 1));
        const double cost = W * (11.0 * double(n) + 32.0 * NB) + 2000.0 * NB / 32.0;
        if (cost < best_cost) { best_cost = cost; best = c; }
    }
    return best;
}

struct Plan {
    MsmArgs a;
    size_t bytes;
    uint32_t *bsum;
    uint32_t nscan;
};

size_t align_up(size_t x, size_t al) { return (x + al - 1) / al * al; }

template <class G>
Plan make_plan(const b200msm_ctx *ctx, size_t n, int c, char *base) {
    typedef AccCfg<G> AC;
    constexpr size_t JACB = 3 * G::F::DEG * NLIMB * 4;
    Plan p;
    MsmArgs &a = p.a;
    memset(&a, 0, sizeof a);
    a.n = (uint32_t)n;
    a.c = c;
    a.W = (MNT753_NUM_BITS + 1 + c - 1) / c;
    a.NB = 1u << (c - 1);
    a.K = (uint32_t)a.W * a.NB;
    const uint64_t emax = (uint64_t)n * a.W;
    // lanes resident on the device in k_accumulate; aim at ~6 chunks per lane for load balance
    const uint64_t lanes = (uint64_t)ctx->sm_count * AC::MINB * AC::TPB * 32;
    uint64_t L = emax / (lanes * 6);
    if (L < 8) L = 8;
    if (L > 256) L = 256;
    a.L = (uint32_t)L;
    a.max_chunks = (uint32_t)((emax + L - 1) / L) + 1;
    // bucket-reduce segment length: keep >= ~32k lanes of segments when there are that many buckets
    uint32_t m = 1;
    while ((uint64_t)a.K / (m * 2) >= 32768 && m * 2 <= a.NB && m < 64) m *= 2;
    a.m = m;
    a.nseg = a.NB / m;
    p.nscan = (a.K + SCAN_B - 1) / SCAN_B;

    size_t off = 0;
    auto take = [&](size_t bytes) { char *q = base + off; off += align_up(bytes, 256); return q; };
    a.scalars = (uint32_t *)take(n * NLIMB * 4);
    a.count = (uint32_t *)take((size_t)a.K * 4);
    a.offs = (uint32_t *)take(((size_t)a.K + 1) * 4);
    a.cursor = (uint32_t *)take((size_t)a.K * 4);
    p.bsum = (uint32_t *)take((size_t)p.nscan * 4);
    a.group_counter = (uint32_t *)take(4);
    a.entries = (uint32_t *)take((size_t)emax * 4 + 4);
    a.buckets = (uint32_t *)take((size_t)a.K * JACB);
    a.edges = (uint32_t *)take((size_t)a.max_chunks * 2 * JACB);
    a.edge_bucket = (uint32_t *)take((size_t)a.max_chunks * 2 * 4);
    {
        const size_t n1 = (a.max_chunks + FOLD_GS - 1) / FOLD_GS, n2 = (n1 + FOLD_GS - 1) / FOLD_GS;
        a.fold_pts[0] = (uint32_t *)take(n1 * 2 * JACB);
        a.fold_key[0] = (uint32_t *)take(n1 * 2 * 4);
        a.fold_pts[1] = (uint32_t *)take(n2 * 2 * JACB);
        a.fold_key[1] = (uint32_t *)take(n2 * 2 * 4);
    }
    a.segsum = (uint32_t *)take((size_t)a.W * a.nseg * JACB);
    const size_t lvl = (size_t)a.W * ((a.nseg + 31) / 32) * JACB;
    a.tmp_a = (uint32_t *)take(lvl);
    a.tmp_b = (uint32_t *)take(lvl);
    a.winsum = (uint32_t *)take((size_t)a.W * JACB);
    a.result = (uint32_t *)take(JACB);
    p.bytes = off;
    return p;
}

template <class K>
int set_smem(b200msm_ctx *ctx, K kernel, size_t bytes) {
    CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return B200MSM_OK;
}

int grow_arena(b200msm_ctx *ctx, Lane &ln, size_t bytes) {
    if (bytes <= ln.arena_bytes) return B200MSM_OK;
    CU(cudaStreamSynchronize(ln.stream));
    if (ln.arena) CU(cudaFree(ln.arena));
    ln.arena = nullptr;
    ln.arena_bytes = 0;
    const size_t want = bytes + bytes / 8;
    CU(cudaMalloc(&ln.arena, want));
    ln.arena_bytes = want;
    return B200MSM_OK;
}

// Enqueue one MSM on lane `li`.  scalars: host or device pointer (Montgomery Fr).
template <class G>
int enqueue_msm(b200msm_ctx *ctx, int li, const BaseSet &bs, size_t offset, const uint64_t *scalars, size_t n,
                uint64_t *out_xyz) {
    typedef typename G::F F;
    typedef AccCfg<G> AC;
    typedef TailCfg<G> TC;
    constexpr int DEG = F::DEG;
    constexpr size_t AFFW = 2 * DEG * NLIMB, JACW = 3 * DEG * NLIMB;
    Lane &ln = ctx->lanes[li];
    cudaStream_t st = ln.stream;
    ln.out_words = JACW / 2;
    ln.user_out = out_xyz;

    if (n == 0) {
        // empty sum: infinity, reported as (1, 1, 0) in Montgomery form like curves.cu:104-114
        memset(ln.h_result, 0, JACW * 4);
        for (int i = 0; i < NLIMB; ++i) {
            ln.h_result[i] = F::M::R1(i);
            ln.h_result[DEG * NLIMB + i] = F::M::R1(i);
        }
        ln.pending = true;
        ln.timed = false;
        memset(ln.ms, 0, sizeof ln.ms);
        memset(ln.info, 0, sizeof ln.info);
        return B200MSM_OK;
    }

    const int c = ctx->c_override ? ctx->c_override : auto_window_bits(n);
    Plan probe = make_plan<G>(ctx, n, c, nullptr);
    int rc = grow_arena(ctx, ln, probe.bytes);
    if (rc) return rc;
    Plan p = make_plan<G>(ctx, n, c, ln.arena);
    MsmArgs &a = p.a;
    a.bases = bs.pts + offset * AFFW;
    a.base_inf = bs.inf + offset;

    if ((rc = set_smem(ctx, k_accumulate<G>, AC::TS::SMEM))) return rc;
    if ((rc = set_smem(ctx, k_fold_edges<G>, TC::TS::SMEM))) return rc;
    if ((rc = set_smem(ctx, k_bucket_reduce<G>, TC::TS::SMEM))) return rc;
    if ((rc = set_smem(ctx, k_sum<G>, TC::TS::SMEM))) return rc;
    if ((rc = set_smem(ctx, k_horner<G>, TC::TS::SMEM))) return rc;

    uint64_t launches = 0;
    CU(cudaEventRecord(ln.ev[0], st));
    CU(cudaMemcpyAsync(a.scalars, scalars, n * NLIMB * 4, cudaMemcpyDefault, st));
    CU(cudaEventRecord(ln.ev[1], st));

    const unsigned nb128 = (unsigned)((n + 127) / 128), nb256 = (unsigned)((n + 255) / 256);
    k_from_mont<typename G::Fr><<<nb128, 128, 0, st>>>(a.scalars, a.n);
    CU(cudaMemsetAsync(a.count, 0, (size_t)a.K * 4, st));
    k_count<<<nb256, 256, 0, st>>>(a);
    k_scan_local<<<p.nscan, SCAN_T, 0, st>>>(a.count, a.offs, p.bsum, a.K);
    k_scan_bsum<<<1, SCAN_T, 0, st>>>(p.bsum, p.nscan, a.offs + a.K);
    k_scan_add<<<p.nscan, SCAN_T, 0, st>>>(a.offs, a.cursor, p.bsum, a.K);
    k_scatter<<<nb256, 256, 0, st>>>(a);
    launches += 6;
    CU(cudaEventRecord(ln.ev[2], st));

    CU(cudaMemsetAsync(a.group_counter, 0, 4, st));
    CU(cudaMemsetAsync(a.edge_bucket, 0xff, (size_t)a.max_chunks * 2 * 4, st));
    k_accumulate<G><<<ctx->sm_count * AC::MINB, AC::TS::THREADS, AC::TS::SMEM, st>>>(a);
    launches += 1;
    CU(cudaEventRecord(ln.ev[3], st));

    const unsigned tail_lanes = TC::TPB * 32;
    {
        const uint32_t *in_pts = a.edges, *in_key = a.edge_bucket;
        uint32_t n_in = a.max_chunks;
        unsigned long long span = a.L;
        int flip = 0;
        do {
            const uint32_t n_out = (n_in + FOLD_GS - 1) / FOLD_GS;
            span *= FOLD_GS;
            k_fold_edges<G><<<(n_out + tail_lanes - 1) / tail_lanes, TC::TS::THREADS, TC::TS::SMEM, st>>>(
                a, in_pts, in_key, n_in, a.fold_pts[flip], a.fold_key[flip], n_out, span);
            ++launches;
            in_pts = a.fold_pts[flip];
            in_key = a.fold_key[flip];
            n_in = n_out;
            flip ^= 1;
        } while (n_in > 1);
    }
    k_bucket_reduce<G><<<((unsigned)a.W * a.nseg + tail_lanes - 1) / tail_lanes, TC::TS::THREADS, TC::TS::SMEM, st>>>(a);
    launches += 1;
    {
        const uint32_t *in = a.segsum;
        uint32_t nin = a.nseg;
        uint32_t *bufs[2] = {a.tmp_a, a.tmp_b};
        int flip = 0;
        while (nin > 1) {
            const uint32_t nout = (nin + 31) / 32;
            uint32_t *out = nout == 1 ? a.winsum : bufs[flip];
            k_sum<G><<<((unsigned)a.W * nout + tail_lanes - 1) / tail_lanes, TC::TS::THREADS, TC::TS::SMEM, st>>>(
                in, out, (uint32_t)a.W, nin, 32u);
            ++launches;
            in = out;
            nin = nout;
            flip ^= 1;
        }
        if (in != a.winsum) a.winsum = const_cast<uint32_t *>(in);  // nseg == 1: segment sums are the window sums
    }
    k_horner<G><<<1, TC::TS::THREADS, TC::TS::SMEM, st>>>(a);
    ++launches;
    CU(cudaEventRecord(ln.ev[4], st));
    CU(cudaMemcpyAsync(ln.h_result, a.result, JACW * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaEventRecord(ln.ev[5], st));
    CU(cudaGetLastError());

    ln.pending = true;
    ln.timed = true;
    ln.info[0] = (uint64_t)c;
    ln.info[1] = (uint64_t)a.W;
    ln.info[2] = (uint64_t)n * a.W;
    ln.info[3] = 1;
    ln.info[4] = launches;
    return B200MSM_OK;
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
    if (ctx->curve == B200MSM_MNT4753)
        return bs.group == B200MSM_G1 ? enqueue_msm<Mnt4G1>(ctx, lane, bs, offset, scalars, n, out)
                                      : enqueue_msm<Mnt4G2>(ctx, lane, bs, offset, scalars, n, out);
    return bs.group == B200MSM_G1 ? enqueue_msm<Mnt6G1>(ctx, lane, bs, offset, scalars, n, out)
                                  : enqueue_msm<Mnt6G2>(ctx, lane, bs, offset, scalars, n, out);
}

// ---- self-test kernels: the field / point layer exposed elementwise (tests only) --------------
template <class G>
__global__ void __launch_bounds__(TailCfg<G>::TS::THREADS) k_test_field(int op, uint32_t n, const uint32_t *x, const uint32_t *y, uint32_t *out) {
    typedef typename G::F F;
    typedef TailCfg<G> C;
    constexpr int EW = F::DEG * NLIMB;
    extern __shared__ uint4 smem[];
    __shared__ uint32_t s_flags[C::TPB][4];
    int team;
    const Team<F> T = C::TS::make(smem, s_flags, team);
    const uint32_t id = (blockIdx.x * C::TPB + team) * 32 + (threadIdx.x & 31);
    const bool v = id < n;
    g2s(T, 0, x + (size_t)id * EW, v);
    g2s(T, 1, (y ? y : x) + (size_t)id * EW, v);
    T.set_zero(0, !v);
    T.set_zero(1, !v);
    T.sync();
    switch (op) {
        case 0: T.mul(2, 0, 1); break;
        case 1: T.add(2, 0, 1); break;
        case 2: T.sub(2, 0, 1); break;
        case 3: T.sqr(2, 0); break;
        case 5: T.neg_if(2, 0, true); break;
        case 6: T.mul_by_a(2, 0); break;
        case 7: T.mul(0, 0, 1); T.copy(2, 0); break;
        default: T.dbl(2, 0); break;
    }
    T.sync();
    s2g(T, out + (size_t)id * EW, 2, v);
}

// op 0: acc (Jacobian) += q (affine, optional negation flag bit0; bit1: acc is infinity); op 1: full add; op 2: dbl
template <class G>
__global__ void __launch_bounds__(TailCfg<G>::TS::THREADS) k_test_point(int op, uint32_t n, const uint32_t *acc, const uint32_t *q, const uint32_t *flags, uint32_t *out) {
    typedef typename G::F F;
    typedef TailCfg<G> C;
    constexpr int EW = F::DEG * NLIMB;
    extern __shared__ uint4 smem[];
    __shared__ uint32_t s_flags[C::TPB][4];
    int team;
    const Team<F> T = C::TS::make(smem, s_flags, team);
    const uint32_t id = (blockIdx.x * C::TPB + team) * 32 + (threadIdx.x & 31);
    const bool v = id < n;
    const PtSlots s = {0, 1, 2, 6, 7, 8, 9, 10, 11};
    load_jac(T, s.X1, s.Y1, s.Z1, acc + (size_t)id * 3 * EW, v);
    T.set_zero(s.X1, !v); T.set_zero(s.Y1, !v); T.set_zero(s.Z1, !v);
    const uint32_t fl = (v && flags) ? flags[id] : 0u;
    if (op == 0) {
        g2s(T, s.X2, q + (size_t)id * 2 * EW, v);
        g2s(T, s.Y2, q + (size_t)id * 2 * EW + EW, v);
        T.set_zero(s.X2, !v); T.set_zero(s.Y2, !v);
        bool acc_inf = (fl & 2u) != 0;
        Ec<F>::madd(T, s, (fl & 1u) != 0, v, acc_inf);
        T.set_zero(s.Z1, acc_inf);
    } else if (op == 1) {
        load_jac(T, s.X2, s.Y2, s.Z2, q + (size_t)id * 3 * EW, v);
        T.set_zero(s.X2, !v); T.set_zero(s.Y2, !v); T.set_zero(s.Z2, !v);
        Ec<F>::add(T, s, v);
    } else {
        Ec<F>::dbl(T, s, v);
    }
    T.sync();
    store_jac(T, out + (size_t)id * 3 * EW, s.X1, s.Y1, s.Z1, v);
}

template <class G>
int run_test(b200msm_ctx *ctx, bool point, int op, size_t n, const uint64_t *a, const uint64_t *b, const uint32_t *flags, uint64_t *out) {
    typedef TailCfg<G> TC;
    constexpr size_t EB = G::F::DEG * NLIMB * 4;
    const size_t ab = point ? 3 * EB : EB, bb = point ? (op == 0 ? 2 * EB : 3 * EB) : EB, ob = ab;
    uint32_t *da = nullptr, *db = nullptr, *dout = nullptr, *dfl = nullptr;
    CU(cudaMalloc(&da, n * ab));
    CU(cudaMalloc(&dout, n * ob));
    CU(cudaMemcpy(da, a, n * ab, cudaMemcpyHostToDevice));
    if (b) { CU(cudaMalloc(&db, n * bb)); CU(cudaMemcpy(db, b, n * bb, cudaMemcpyHostToDevice)); }
    if (flags) { CU(cudaMalloc(&dfl, n * 4)); CU(cudaMemcpy(dfl, flags, n * 4, cudaMemcpyHostToDevice)); }
    const unsigned lanes = TC::TPB * 32, grid = (unsigned)((n + lanes - 1) / lanes);
    if (point) {
        CU(cudaFuncSetAttribute(k_test_point<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC::TS::SMEM));
        k_test_point<G><<<grid, TC::TS::THREADS, TC::TS::SMEM>>>(op, (uint32_t)n, da, db, dfl, dout);
    } else {
        CU(cudaFuncSetAttribute(k_test_field<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC::TS::SMEM));
        k_test_field<G><<<grid, TC::TS::THREADS, TC::TS::SMEM>>>(op, (uint32_t)n, da, db, dout);
    }
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(out, dout, n * ob, cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(db); cudaFree(dout); cudaFree(dfl);
    return B200MSM_OK;
}

// sum of n Jacobian partial results (one per GPU shard) on the device -> one Jacobian point
template <class G>
int run_fold(b200msm_ctx *ctx, const uint64_t *xyz, size_t n, uint64_t *out) {
    typedef TailCfg<G> TC;
    constexpr size_t JACW = 3 * G::F::DEG * NLIMB, JACB = JACW * 4;
    const size_t lvl = (n + 31) / 32 + 1;
    uint32_t *din = nullptr, *buf[2] = {nullptr, nullptr}, *dres = nullptr;
    CU(cudaMalloc(&din, (n ? n : 1) * JACB));
    CU(cudaMalloc(&buf[0], lvl * JACB));
    CU(cudaMalloc(&buf[1], lvl * JACB));
    CU(cudaMalloc(&dres, JACB));
    CU(cudaMemcpy(din, xyz, n * JACB, cudaMemcpyDefault));
    CU(cudaFuncSetAttribute(k_sum<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC::TS::SMEM));
    CU(cudaFuncSetAttribute(k_horner<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC::TS::SMEM));
    const unsigned tail_lanes = TC::TPB * 32;
    const uint32_t *in = din;
    uint32_t nin = (uint32_t)n;
    int flip = 0;
    do {  // at least one pass so that the input is never aliased
        const uint32_t nout = (nin + 31) / 32;
        k_sum<G><<<(nout + tail_lanes - 1) / tail_lanes, TC::TS::THREADS, TC::TS::SMEM>>>(in, buf[flip], 1u, nin, 32u);
        in = buf[flip];
        nin = nout;
        flip ^= 1;
    } while (nin > 1);
    MsmArgs a;
    memset(&a, 0, sizeof a);
    a.W = 1;
    a.c = 1;
    a.winsum = const_cast<uint32_t *>(in);
    a.result = dres;
    k_horner<G><<<1, TC::TS::THREADS, TC::TS::SMEM>>>(a);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(out, dres, JACB, cudaMemcpyDeviceToHost));
    cudaFree(din); cudaFree(buf[0]); cudaFree(buf[1]); cudaFree(dres);
    return B200MSM_OK;
}

// ---- affine normalisation / synthetic bases ---------------------------------------------------
// exponent q^DEG - 2 of the Fermat inversion in Fq^DEG, little-endian 32-bit words
template <class G>
std::vector<uint32_t> fermat_exponent() {
    constexpr int DEG = G::F::DEG;
    std::vector<uint32_t> q(NLIMB), acc(1, 1u);
    for (int i = 0; i < NLIMB; ++i) q[i] = G::F::M::P(i);
    for (int d = 0; d < DEG; ++d) {
        std::vector<uint32_t> r(acc.size() + NLIMB, 0u);
        for (size_t i = 0; i < acc.size(); ++i) {
            uint64_t carry = 0;
            for (int j = 0; j < NLIMB; ++j) {
                uint64_t t = (uint64_t)acc[i] * q[j] + r[i + j] + carry;
                r[i + j] = (uint32_t)t;
                carry = t >> 32;
            }
            r[i + NLIMB] = (uint32_t)carry;
        }
        acc.swap(r);
    }
    uint64_t borrow = 2;  // acc -= 2 (q is odd and > 2, so no underflow)
    for (size_t i = 0; i < acc.size() && borrow; ++i) {
        uint64_t t = (uint64_t)acc[i] - borrow;
        acc[i] = (uint32_t)t;
        borrow = (t >> 63) & 1u;
    }
    return acc;
}

struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
};

template <class G>
int upload_exponent(b200msm_ctx *ctx, DevBuf &d, int &bits) {
    std::vector<uint32_t> e = fermat_exponent<G>();
    bits = G::F::DEG * MNT753_NUM_BITS;
    CU(cudaMalloc(&d.p, e.size() * 4));
    CU(cudaMemcpy(d.p, e.data(), e.size() * 4, cudaMemcpyHostToDevice));
    return B200MSM_OK;
}

template <class G>
int run_to_affine(b200msm_ctx *ctx, size_t n, const uint64_t *xyz, uint64_t *out) {
    typedef TailCfg<G> TC;
    constexpr size_t EB = G::F::DEG * NLIMB * 4;
    DevBuf e, in, o;
    int bits = 0, rc = upload_exponent<G>(ctx, e, bits);
    if (rc) return rc;
    CU(cudaMalloc(&in.p, n * 3 * EB));
    CU(cudaMalloc(&o.p, n * 2 * EB));
    CU(cudaMemcpy(in.p, xyz, n * 3 * EB, cudaMemcpyDefault));
    CU(cudaFuncSetAttribute(k_to_affine<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC::TS::SMEM));
    const unsigned lanes = TC::TPB * 32;
    k_to_affine<G><<<(unsigned)((n + lanes - 1) / lanes), TC::TS::THREADS, TC::TS::SMEM>>>((uint32_t)n, (const uint32_t *)in.p, (uint32_t *)o.p,
                                                                                         (const uint32_t *)e.p, bits);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(out, o.p, n * 2 * EB, cudaMemcpyDefault));
    return B200MSM_OK;
}

template <class G>
const uint32_t *generator_words() {
    static const uint32_t c0g1[] = MNT753_GEN_C0_G1_U32, c0g2[] = MNT753_GEN_C0_G2_U32, c1g1[] = MNT753_GEN_C1_G1_U32,
                          c1g2[] = MNT753_GEN_C1_G2_U32;
    return G::CURVE == 0 ? (G::GROUP == 1 ? c0g1 : c0g2) : (G::GROUP == 1 ? c1g1 : c1g2);
}

// bases[i] = k_p0 * G + i * (k_q * G), written straight into a new resident base set
template <class G>
int run_synthetic(b200msm_ctx *ctx, size_t n, const uint64_t *k_p0, const uint64_t *k_q, BaseSet &bs) {
    typedef TailCfg<G> TC;
    constexpr size_t EB = G::F::DEG * NLIMB * 4;
    constexpr uint32_t B = 64;
    DevBuf e, gen, ks, pj, pa, jac, pre;
    int bits = 0, rc = upload_exponent<G>(ctx, e, bits);
    if (rc) return rc;
    CU(cudaMalloc(&gen.p, 2 * EB));
    CU(cudaMalloc(&ks.p, 2 * NLIMB * 4));
    CU(cudaMalloc(&pj.p, 2 * 3 * EB));
    CU(cudaMalloc(&pa.p, 2 * 2 * EB));
    CU(cudaMalloc(&jac.p, n * 3 * EB));
    CU(cudaMalloc(&pre.p, n * EB));
    CU(cudaMemcpy(gen.p, generator_words<G>(), 2 * EB, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(ks.p, k_p0, NLIMB * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy((uint32_t *)ks.p + NLIMB, k_q, NLIMB * 4, cudaMemcpyHostToDevice));
    CU(cudaFuncSetAttribute(k_scalar_mul<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC::TS::SMEM));
    CU(cudaFuncSetAttribute(k_to_affine<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC::TS::SMEM));
    CU(cudaFuncSetAttribute(k_synth_bases<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC::TS::SMEM));
    for (int i = 0; i < 2; ++i)
        k_scalar_mul<G><<<1, TC::TS::THREADS, TC::TS::SMEM>>>((const uint32_t *)gen.p, (const uint32_t *)ks.p + i * NLIMB,
                                                             (uint32_t *)pj.p + i * 3 * (EB / 4));
    k_to_affine<G><<<1, TC::TS::THREADS, TC::TS::SMEM>>>(2u, (const uint32_t *)pj.p, (uint32_t *)pa.p, (const uint32_t *)e.p, bits);
    const unsigned lanes = TC::TPB * 32;
    const size_t runs = (n + B - 1) / B;
    k_synth_bases<G><<<(unsigned)((runs + lanes - 1) / lanes), TC::TS::THREADS, TC::TS::SMEM>>>(
        (uint32_t)n, B, (const uint32_t *)pa.p, (const uint32_t *)pa.p + 2 * (EB / 4), bs.pts, (uint32_t *)jac.p, (uint32_t *)pre.p,
        (const uint32_t *)e.p, bits);
    constexpr int DEG = G::F::DEG;
    k_flag_inf<DEG><<<(unsigned)((n + 255) / 256), 256>>>(bs.pts, (uint32_t)n, bs.inf);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    return B200MSM_OK;
}

// ---- microbenchmarks ------------------------------------------------------------------------------
constexpr int MB_CHAINS = 16;
// 16 independent chains per thread; the multiplier of every step is the low word of the chain's own
// accumulator, so nothing is loop-invariant (an earlier version with constant multipliers was hoisted
// by ptxas into 64-bit adds and reported twice the real rate).
__global__ void __launch_bounds__(256) k_mb_wide(uint32_t *out, int iters) {
    uint64_t acc[MB_CHAINS];
    const uint32_t b = 0x9e3779b9u + blockIdx.x;
    for (int j = 0; j < MB_CHAINS; ++j) acc[j] = (uint64_t)(threadIdx.x + 1) * (j + 3) + 0x100000001ull * j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < MB_CHAINS; ++j) {
            const uint32_t m = (uint32_t)acc[j];
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(m), "r"(b));
        }
    }
    uint64_t s = 0;
    for (int j = 0; j < MB_CHAINS; ++j) s ^= acc[j];
    if (s == 0x123456789abcdefull) out[0] = 1;
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
template <class M, bool RR>
__global__ void __launch_bounds__(128) k_mb_fqmul(uint32_t *out, int iters) {
    fq_t x, y;
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) { x[i] = M::R1(i) ^ (threadIdx.x * 7u + i); y[i] = M::R2(i) ^ (blockIdx.x + i); }
    x[NLIMB - 1] &= 0xffffu;
    y[NLIMB - 1] &= 0xffffu;
    for (int it = 0; it < iters; ++it) {
        if (RR) fq_mul_rr<M>(x, x, y); else fq_mul<M>(x, x, y);
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) s ^= x[i];
    if (s == 0x12345678u) out[0] = 1;
}

}  // namespace

// ================================================================================================
extern "C" {

int b200msm_create(int curve, int device, b200msm_ctx **out) {
    if (!out || (curve != B200MSM_MNT4753 && curve != B200MSM_MNT6753)) return B200MSM_ERR_ARG;
    *out = nullptr;
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
        if (!ok) { b200msm_destroy(ctx); return B200MSM_ERR_CUDA; }
    }
    *out = ctx;
    return B200MSM_OK;
}

void b200msm_destroy(b200msm_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    for (int i = 0; i < NLANES; ++i) {
        Lane &ln = ctx->lanes[i];
        if (ln.stream) cudaStreamSynchronize(ln.stream);
        if (ln.arena) cudaFree(ln.arena);
        if (ln.h_result) cudaFreeHost(ln.h_result);
        for (int e = 0; e < NEVENTS; ++e) if (ln.ev[e]) cudaEventDestroy(ln.ev[e]);
        if (ln.own_stream) cudaStreamDestroy(ln.own_stream);
    }
    for (auto &s : ctx->sets) if (s.used) { cudaFree(s.pts); cudaFree(s.inf); }
    delete ctx;
}

const char *b200msm_last_error(const b200msm_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int b200msm_bases_upload(b200msm_ctx *ctx, int group, const uint64_t *affine, size_t n, int *slot) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (!slot || (group != B200MSM_G1 && group != B200MSM_G2) || (n && !affine)) return fail(ctx, B200MSM_ERR_ARG, "bad argument");
    if (n >= (size_t(1) << 31)) return fail(ctx, B200MSM_ERR_ARG, "n = %zu too large", n);
    CU(cudaSetDevice(ctx->device));
    const int deg = degree_of(ctx->curve, group);
    const size_t bytes = n * 2 * deg * NLIMB * 4;
    BaseSet bs;
    bs.used = true;
    bs.group = group;
    bs.n = n;
    CU(cudaMalloc(&bs.pts, bytes ? bytes : 256));
    cudaError_t e = cudaMalloc(&bs.inf, n ? n : 256);
    if (e != cudaSuccess) { cudaFree(bs.pts); return fail(ctx, B200MSM_ERR_OOM, "cudaMalloc: %s", cudaGetErrorString(e)); }
    if (n) {
        e = cudaMemcpy(bs.pts, affine, bytes, cudaMemcpyDefault);
        if (e == cudaSuccess) {
            const unsigned grid = (unsigned)((n + 255) / 256);
            if (deg == 1) k_flag_inf<1><<<grid, 256>>>(bs.pts, (uint32_t)n, bs.inf);
            else if (deg == 2) k_flag_inf<2><<<grid, 256>>>(bs.pts, (uint32_t)n, bs.inf);
            else k_flag_inf<3><<<grid, 256>>>(bs.pts, (uint32_t)n, bs.inf);
            e = cudaDeviceSynchronize();
        }
        if (e != cudaSuccess) { cudaFree(bs.pts); cudaFree(bs.inf); return fail(ctx, B200MSM_ERR_CUDA, "base upload: %s", cudaGetErrorString(e)); }
    }
    int id = -1;
    for (size_t i = 0; i < ctx->sets.size(); ++i) if (!ctx->sets[i].used) { id = (int)i; break; }
    if (id < 0) { ctx->sets.push_back(bs); id = (int)ctx->sets.size() - 1; } else ctx->sets[id] = bs;
    *slot = id;
    return B200MSM_OK;
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
    int rc = b200msm_bases_upload(ctx, group, bases_affine, n, &slot);
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
    if (ctx->curve == B200MSM_MNT4753)
        return group == B200MSM_G1 ? run_fold<Mnt4G1>(ctx, partials_xyz, n, out_xyz) : run_fold<Mnt4G2>(ctx, partials_xyz, n, out_xyz);
    return group == B200MSM_G1 ? run_fold<Mnt6G1>(ctx, partials_xyz, n, out_xyz) : run_fold<Mnt6G2>(ctx, partials_xyz, n, out_xyz);
}

int b200msm_to_affine(b200msm_ctx *ctx, int group, size_t n, const uint64_t *xyz, uint64_t *out_affine) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (!xyz || !out_affine || n == 0 || n > (1u << 24) || (group != B200MSM_G1 && group != B200MSM_G2))
        return fail(ctx, B200MSM_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    if (ctx->curve == B200MSM_MNT4753)
        return group == B200MSM_G1 ? run_to_affine<Mnt4G1>(ctx, n, xyz, out_affine) : run_to_affine<Mnt4G2>(ctx, n, xyz, out_affine);
    return group == B200MSM_G1 ? run_to_affine<Mnt6G1>(ctx, n, xyz, out_affine) : run_to_affine<Mnt6G2>(ctx, n, xyz, out_affine);
}

int b200msm_bases_synthetic(b200msm_ctx *ctx, int group, size_t n, const uint64_t *k_p0_mont, const uint64_t *k_q_mont, int *slot) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (!slot || !k_p0_mont || !k_q_mont || n == 0 || n >= (size_t(1) << 31) || (group != B200MSM_G1 && group != B200MSM_G2))
        return fail(ctx, B200MSM_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    const int deg = degree_of(ctx->curve, group);
    BaseSet bs;
    bs.used = true;
    bs.group = group;
    bs.n = n;
    CU(cudaMalloc(&bs.pts, n * 2 * deg * NLIMB * 4));
    cudaError_t e = cudaMalloc(&bs.inf, n);
    if (e != cudaSuccess) { cudaFree(bs.pts); return fail(ctx, B200MSM_ERR_OOM, "cudaMalloc: %s", cudaGetErrorString(e)); }
    int rc;
    if (ctx->curve == B200MSM_MNT4753)
        rc = group == B200MSM_G1 ? run_synthetic<Mnt4G1>(ctx, n, k_p0_mont, k_q_mont, bs) : run_synthetic<Mnt4G2>(ctx, n, k_p0_mont, k_q_mont, bs);
    else
        rc = group == B200MSM_G1 ? run_synthetic<Mnt6G1>(ctx, n, k_p0_mont, k_q_mont, bs) : run_synthetic<Mnt6G2>(ctx, n, k_p0_mont, k_q_mont, bs);
    if (rc) { cudaFree(bs.pts); cudaFree(bs.inf); return rc; }
    int id = -1;
    for (size_t i = 0; i < ctx->sets.size(); ++i) if (!ctx->sets[i].used) { id = (int)i; break; }
    if (id < 0) { ctx->sets.push_back(bs); id = (int)ctx->sets.size() - 1; } else ctx->sets[id] = bs;
    *slot = id;
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

int b200msm_set_window_bits(b200msm_ctx *ctx, int c) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (c != 0 && (c < 2 || c > 20)) return fail(ctx, B200MSM_ERR_ARG, "window bits %d outside [2, 20]", c);
    ctx->c_override = c;
    return B200MSM_OK;
}

int b200msm_last_timings(b200msm_ctx *ctx, int lane, float ms[6], uint64_t info[5]) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (lane < 0 || lane >= NLANES) return fail(ctx, B200MSM_ERR_ARG, "lane %d out of range", lane);
    if (ms) memcpy(ms, ctx->lanes[lane].ms, sizeof(float) * 6);
    if (info) memcpy(info, ctx->lanes[lane].info, sizeof(uint64_t) * 5);
    return B200MSM_OK;
}

int b200msm_microbench(b200msm_ctx *ctx, int kind, int iters, double *gops) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (!gops || iters <= 0 || kind < 0 || kind > 3) return fail(ctx, B200MSM_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    uint32_t *d = nullptr;
    CU(cudaMalloc(&d, 256));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    const int blocks = ctx->sm_count * 8;
    double ops = 0;
    for (int rep = 0; rep < 2; ++rep) {  // first repetition is the warm-up
        CU(cudaEventRecord(e0, 0));
        if (kind == 0) { k_mb_wide<<<blocks, 256>>>(d, iters); ops = double(blocks) * 256 * iters * MB_CHAINS; }
        else if (kind == 1) { k_mb_lo<<<blocks, 256>>>(d, iters); ops = double(blocks) * 256 * iters * MB_CHAINS; }
        else {
            if (kind == 3) k_mb_fqmul<ModA, true><<<blocks, 128>>>(d, iters);
            else if (ctx->curve == B200MSM_MNT4753) k_mb_fqmul<ModA, false><<<blocks, 128>>>(d, iters);
            else k_mb_fqmul<ModB, false><<<blocks, 128>>>(d, iters);
            ops = double(blocks) * 128 * iters;
        }
        CU(cudaEventRecord(e1, 0));
        CU(cudaEventSynchronize(e1));
    }
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, e0, e1));
    CU(cudaGetLastError());
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *gops = ops / (double(ms) * 1e6);
    return B200MSM_OK;
}

int b200msm_selftest_field(b200msm_ctx *ctx, int group, int op, size_t n, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (!a || !out || n == 0 || n > (1u << 24)) return fail(ctx, B200MSM_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    if (ctx->curve == B200MSM_MNT4753)
        return group == B200MSM_G1 ? run_test<Mnt4G1>(ctx, false, op, n, a, b, nullptr, out) : run_test<Mnt4G2>(ctx, false, op, n, a, b, nullptr, out);
    return group == B200MSM_G1 ? run_test<Mnt6G1>(ctx, false, op, n, a, b, nullptr, out) : run_test<Mnt6G2>(ctx, false, op, n, a, b, nullptr, out);
}

int b200msm_selftest_point(b200msm_ctx *ctx, int group, int op, size_t n, const uint64_t *acc, const uint64_t *q, const uint32_t *flags, uint64_t *out) {
    if (!ctx) return B200MSM_ERR_ARG;
    if (!acc || !out || n == 0 || n > (1u << 22) || op < 0 || op > 2 || (op != 2 && !q)) return fail(ctx, B200MSM_ERR_ARG, "bad argument");
    CU(cudaSetDevice(ctx->device));
    if (ctx->curve == B200MSM_MNT4753)
        return group == B200MSM_G1 ? run_test<Mnt4G1>(ctx, true, op, n, acc, q, flags, out) : run_test<Mnt4G2>(ctx, true, op, n, acc, q, flags, out);
    return group == B200MSM_G1 ? run_test<Mnt6G1>(ctx, true, op, n, acc, q, flags, out) : run_test<Mnt6G2>(ctx, true, op, n, acc, q, flags, out);
}

}  // extern "C"
