// Host-side state shared by the translation units of libb200msm.so (msm.cu holds the C ABI, inst_*.cu
// the per-group kernels).  One context per (curve, GPU); five internal streams ("lanes"), one per query of a
// proof, so that the A, B1, B2 and L multiexps can be in flight together, as the reference does with one stream
// per MSM (cuda_prover_piecewise.cu:162-167), and the H query can follow its FFTs without waiting for any of them.
#pragma once
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/b200_msm.h"
#include "msm_kernels.cuh"
#include "batch_affine.cuh"
#include "bucket_tree.cuh"
#include "glv.cuh"
#include "util_kernels.cuh"

using namespace mnt753;

constexpr int NLANES = 5;
constexpr int NEVENTS = 7;
constexpr int NCOPY = 4;     // chunks of a host-scalar upload

// A resident base set.  Besides the points themselves (table 0) it may hold NT - 1 further
// "window tables": table t = 2^(c*G*t) * P_i, affine, row t*n + i of `pts`.  Signed digit w of a
// scalar then adds table (w / G) into bucket set (w % G), so an MSM needs only G = ceil(Wd / NT)
// bucket sets, G running-sum reductions and (G - 1) * c doublings in the window combine, instead of
// Wd = ceil(754 / c) of each.  With NT = Wd there is a single bucket set and no combine at all.
// This is the engine's use of HBM capacity (180 GB): the reference spent 25 GB of *disk* on its
// 31x multiples table (main.cpp:311-339); these tables are built on the device at upload time.
struct BaseSet {
    bool used = false;
    int group = 0;
    size_t n = 0;
    uint32_t *pts = nullptr;
    uint8_t *inf = nullptr;
    int c_tab = 0;  // window bits the tables were built for (0: no tables, any c allowed)
    int NT = 1;     // number of tables
    int G = 0;      // bucket sets when the tables are used
    int Wd = 0;     // digits per scalar when the tables are used
    bool glv = false;  // G2: scalars are split k = k0 + k1 lam, tables NT/2 .. NT-1 are psi of tables 0 .. NT/2-1 (glv.cuh)
    float build_ms = 0.f;
};

struct Lane {
    cudaStream_t stream = nullptr;      // stream MSMs are enqueued on
    cudaStream_t own_stream = nullptr;  // the lane's internal stream (stream == own_stream unless overridden)
    cudaEvent_t ev[NEVENTS] = {};
    cudaStream_t copy_stream = nullptr;       // H2D of host scalars, overlapped with the recode of earlier chunks
    cudaEvent_t ev_copy[NCOPY + 1] = {};
    char *arena = nullptr;
    size_t arena_bytes = 0;
    uint32_t *h_result = nullptr;  // pinned staging for the Jacobian result
    uint32_t *h_ctl = nullptr;     // pinned copy of the accumulation's statistics (BA_CTL_WORDS: rounds, largest bucket, additions, pairs per round)
    bool ctl_valid = false;
    uint64_t *user_out = nullptr;
    size_t out_words = 0;
    bool pending = false;
    bool timed = false;
    float ms[6] = {};
    uint64_t info[8] = {};
    uint64_t dbg[16] = {};         // arena offsets of the last MSM's lists (development introspection, b200msm_internal_debug_*)
};

// H-polynomial state (fft.cu): tables of the evaluation domain of size 2^logm and three work vectors
struct FftState {
    int logm = -1;
    uint32_t *consts = nullptr, *tw = nullptr, *twi = nullptr, *cg_br = nullptr, *cgi_br = nullptr;
    uint32_t *a = nullptr, *b = nullptr, *c = nullptr, *out = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[2] = {};
    float last_ms = 0.f, tables_ms = 0.f;
    unsigned staged = 0;     // bit v: input vector v of the compute_H in progress is uploaded and transformed (fft.cu)
};

// Scratch of the small host-facing routines (fold of partial points, affine normalisation, proof assembly): one
// device buffer, one pinned host buffer and one stream per context, grown on demand and kept -- no allocation, no
// device-wide synchronisation on the per-proof path.
struct TailBuf {
    cudaStream_t st = nullptr;
    char *d = nullptr, *h = nullptr;
    size_t dbytes = 0, hbytes = 0;
};

struct b200msm_ctx {
    int curve = 0;
    int device = 0;
    int sm_count = 0;
    int c_override = 0;
    bool kernels_ready[3] = {false, false, false};  // per group: shared-memory attributes of its kernels set on this device
    size_t table_budget = size_t(32) << 30;  // bytes of window tables per base set (0: never build tables)
    std::vector<BaseSet> sets;
    Lane lanes[NLANES];
    int lane_sms[NLANES] = {0, 0, 0, 0, 0};     // SMs the MSMs of a lane may occupy (0: all of them), b200msm_set_lane_sms
    FftState fft;
    TailBuf tail;
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

// a device allocation / a pair of timing events that are released on every return path (CU() returns early)
struct DevBuf {
    void *p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
};
struct EventPair {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    ~EventPair() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); }
};

inline int degree_of(int curve, int group) { return group == B200MSM_G1 ? 1 : (curve == B200MSM_MNT4753 ? 2 : 3); }

// ---- window-size / table choice ---------------------------------------------------------------
// Time model in nanoseconds, calibrated on B200 (profiles/r02_c_sweep.txt, profiles/r02_size_sweep_1gpu.txt): ~1.0 ns per
// batched-affine addition (one per point and digit; x3 / x6 in the towers); ~0.3 ms per round of the accumulation
// (plan, tile inversion, two passes -- its time follows the NUMBER OF ROUNDS much more than the number of additions
// below 2^18 points: 2^12 points take 2.9 ms at c = 12 and 0.8 ms at c = 18); the bucket reduction (bucket_tree.cuh)
// ~3 ns per bucket, ~0.1 ms per tree level, ~30 us per doubling of the final terms and ~0.45 ms for the butterflies of
// the short lists and the window sum; bucket sets after the first cost c doublings each in the window combine.
// Rounds = log2 of the LARGEST bucket: a typical one, and the top window's, which holds only rem = 754 - (Wd - 1) c
// bits and therefore piles n / 2^(rem-1) points on each of its few buckets (c = 16: rem = 2, a quarter of all points
// in one bucket).
struct TabCfg { int c, Wd, NT, G; bool glv; };
inline int digits_for(int c) { return (MNT753_NUM_BITS + 1 + c - 1) / c; }
inline int half_digits_for(int c) { return (MNT753_GLV_HALF_BITS + 1 + c - 1) / c; }   // of one half of a split G2 scalar

// The model's two terms for one MSM of n points in a configuration: `work_ns` shrinks with the SMs the MSM runs on
// (additions, per-bucket work of the reduction), `latency_ns` does not (rounds, tree levels, the serial tail).
struct MsmCost { double work_ns, latency_ns; };
inline MsmCost model_cost(size_t n, int deg, const TabCfg &cfg) {
    const double k = deg == 1 ? 1.0 : (deg == 2 ? 3.0 : 6.0);
    const int c = cfg.c;
    // bits of the top window (of a half's top window for split scalars; <= 0: it stays empty)
    const int rem = cfg.glv ? MNT753_GLV_HALF_BITS + 1 - (cfg.Wd / 2 - 1) * c : MNT753_NUM_BITS + 1 - (cfg.Wd - 1) * c;
    const double NB = double(1u << (c - 1));
    const double avg = double(n) * cfg.NT / NB;                                 // digits per bucket of a set
    const double top = (rem >= c || rem <= 0) ? 0.0 : double(n) / double(1u << (rem > 1 ? rem - 1 : 0));
    // rounds = log2 of the largest bucket: of a typical one (Poisson tail over the average occupancy), which every
    // share goes through, plus the further ones of the top window's overfull buckets, which only the shares holding
    // their pieces run (and whose keys contend in the sort)
    const double typical = avg + 4.0 * sqrt(avg) + 3.0;
    double rounds = 1.0, rounds_top = 0.0;
    for (double occ = typical; occ > 1.0; occ *= 0.5) rounds += 1.0;
    for (double occ = (typical + top) / typical; occ > 1.0; occ *= 0.5) rounds_top += 1.0;
    const double lat = deg == 1 ? 1.0 : (deg == 2 ? 1.4 : 1.9);               // latency of one field operation in the towers
    const double levels = c > 6 ? double(c - 6) : 0.0;
    MsmCost m;
    m.work_ns = double(cfg.Wd) * double(n) * 1.0 * k + double(cfg.G) * NB * 3.0 * k;
    m.latency_ns = (rounds * 300000.0 + rounds_top * 200000.0) * lat + (levels * 100000.0 + 450000.0 + double(c) * 30000.0) * lat +
                   double(cfg.G - 1) * c * 35000.0 * lat;
    return m;
}

inline TabCfg choose_cfg(size_t n, int deg, int c_fixed, size_t budget_bytes, bool tables) {
    const size_t affb = (size_t)2 * deg * NLIMB * 4;
    TabCfg best = {2, digits_for(2), 1, digits_for(2), false};
    double best_cost = 1e300;
    for (int c = (c_fixed ? c_fixed : 2); c <= (c_fixed ? c_fixed : 22); ++c) {
        int Wd = digits_for(c);
        size_t nt = 1;
        if (tables && n >= 256 && budget_bytes) {
            nt = budget_bytes / (n * affb);
            const size_t idx_cap = ((size_t(1) << 31) - 1) / n;
            if (nt > idx_cap) nt = idx_cap;
            if (nt < 1) nt = 1;
        }
        // G2 with at least two tables: split scalars (glv.cuh) -- two halves of Wh digits, Wh a multiple of the
        // number of bucket sets so that a table belongs to one half
        TabCfg cfg;
        cfg.c = c;
        if (deg > 1 && nt >= 2) {
            int Wh = half_digits_for(c);
            if (nt > (size_t)(2 * Wh)) nt = 2 * Wh;
            nt &= ~size_t(1);                                                      // the two halves have the same number of tables
            cfg.G = (2 * Wh + (int)nt - 1) / (int)nt;
            Wh = (Wh + cfg.G - 1) / cfg.G * cfg.G;
            cfg.Wd = 2 * Wh;
            cfg.NT = cfg.Wd / cfg.G;
            cfg.glv = true;
        } else {
            if (nt > (size_t)Wd) nt = Wd;
            cfg.Wd = Wd;
            cfg.G = (Wd + (int)nt - 1) / (int)nt;
            cfg.NT = (Wd + cfg.G - 1) / cfg.G;
            cfg.glv = false;
        }
        const MsmCost m = model_cost(n, deg, cfg);
        const double cost = m.work_ns + m.latency_ns;
        if (cost < best_cost) { best_cost = cost; best = cfg; }
    }
    return best;
}

// configuration of an MSM of n points over a resident base set: its window tables are used when they exist and the
// caller did not force a different window width
inline TabCfg cfg_for_set(const b200msm_ctx *ctx, const BaseSet &bs, size_t n) {
    if (bs.c_tab && (ctx->c_override == 0 || ctx->c_override == bs.c_tab)) return TabCfg{bs.c_tab, bs.Wd, bs.NT, bs.G, bs.glv};
    return choose_cfg(n, degree_of(ctx->curve, bs.group), ctx->c_override, 0, false);
}

// SMs the MSMs of lane `li` may occupy
inline int lane_sm_cap(const b200msm_ctx *ctx, int li) {
    const int s = ctx->lane_sms[li];
    return (s > 0 && s < ctx->sm_count) ? s : ctx->sm_count;
}

// make the context's tail buffers at least this large (contents are not preserved)
inline int tail_reserve(b200msm_ctx *ctx, size_t dbytes, size_t hbytes) {
    TailBuf &t = ctx->tail;
    if (!t.st) CU(cudaStreamCreateWithFlags(&t.st, cudaStreamNonBlocking));
    if (dbytes > t.dbytes) {
        CU(cudaStreamSynchronize(t.st));
        if (t.d) CU(cudaFree(t.d));
        t.d = nullptr; t.dbytes = 0;
        const size_t want = std::max<size_t>(dbytes + dbytes / 4, size_t(1) << 20);
        CU(cudaMalloc(&t.d, want));
        t.dbytes = want;
    }
    if (hbytes > t.hbytes) {
        CU(cudaStreamSynchronize(t.st));
        if (t.h) CU(cudaFreeHost(t.h));
        t.h = nullptr; t.hbytes = 0;
        const size_t want = std::max<size_t>(hbytes + hbytes / 4, size_t(1) << 16);
        CU(cudaMallocHost(&t.h, want));
        t.hbytes = want;
    }
    return B200MSM_OK;
}

// scratch points of GroupOps::fold_dev for n inputs
inline size_t fold_scratch_points(size_t n) { return 2 * ((n + 31) / 32 + 1); }

struct Plan {
    MsmArgs a;
    BaArgs b;        // the batched-affine accumulation (batch_affine.cuh)
    TreeArgs t;      // the bucket-reduction tree (bucket_tree.cuh)
    uint32_t *fin;   // its k + 1 Jacobian terms per set
    uint64_t slots;  // scratch points in all
    size_t bytes;
    uint32_t *bsum;
    uint32_t nscan;
};

size_t align_up(size_t x, size_t al) { return (x + al - 1) / al * al; }

}  // namespace

// per-group entry points (inst_*.cu)
struct GroupOps {
    int (*enqueue)(b200msm_ctx *, int, const BaseSet &, size_t, const uint64_t *, size_t, uint64_t *);
    int (*selftest)(b200msm_ctx *, bool, int, size_t, const uint64_t *, const uint64_t *, const uint32_t *, uint64_t *);
    int (*fold)(b200msm_ctx *, const uint64_t *, size_t, uint64_t *);
    int (*to_affine)(b200msm_ctx *, size_t, const uint64_t *, uint64_t *);
    // device-side pieces of the two (no copies, no synchronisation): din n Jacobian points -> dres one; scratch of fold_scratch(n) bytes
    int (*fold_dev)(b200msm_ctx *, cudaStream_t, const uint32_t *, size_t, uint32_t *, uint32_t *);
    int (*to_affine_dev)(b200msm_ctx *, cudaStream_t, size_t, const uint32_t *, uint32_t *);
    int (*synthetic)(b200msm_ctx *, size_t, const uint64_t *, const uint64_t *, BaseSet &);
    int (*build_tables)(b200msm_ctx *, BaseSet &);
    int (*teammul_bench)(b200msm_ctx *, int, int, double *);
    int (*scalar_mul)(b200msm_ctx *, const uint64_t *, const uint64_t *, uint64_t *);
    int (*reserve)(b200msm_ctx *, int, const BaseSet &, size_t);
};
