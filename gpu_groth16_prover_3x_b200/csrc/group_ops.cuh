// Per-group (curve x G1/G2) host routines: everything that launches a kernel templated on the group.
// Each of the four groups is instantiated in its own translation unit (inst_*.cu) so that the library
// builds in parallel; msm.cu reaches them through the GroupOps table.
#pragma once
#include "host_ctx.cuh"

namespace {

#ifndef B200_TREE_AFFINE_MIN
#define B200_TREE_AFFINE_MIN 100000
#endif
// additions in a reduction round from which batched-affine beats one Jacobian addition per lane, for Fq; the towers'
// Jacobian additions are dearer in proportion (measured: profiles/r02_ab_tree_threshold.txt)
constexpr uint64_t TREE_AFFINE_MIN = B200_TREE_AFFINE_MIN;

// teams that take a share of an MSM with at most `emax` sorted entries on `sms` SMs: every team of the persistent grid,
// but no share below 256 entries (eight additions per lane: below that a round is all fixed cost)
template <class G>
uint32_t shares_for(int sms, uint64_t emax) {
    const uint64_t teams = (uint64_t)sms * BaCfg<G>::TPB;
    if (const char *e = getenv("B200MSM_SHARES")) return (uint32_t)std::max(1, atoi(e));   // development knob
    return (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(teams, emax / 256));
}

template <class G>
Plan make_plan(int sms, size_t n, const TabCfg &cfg, char *base) {
    constexpr size_t JACB = 3 * G::F::DEG * NLIMB * 4, AFFB = 2 * G::F::DEG * NLIMB * 4;
    Plan p;
    MsmArgs &a = p.a;
    BaArgs &b = p.b;
    memset(&a, 0, sizeof a);
    memset(&b, 0, sizeof b);
    const int c = cfg.c;
    a.n = (uint32_t)n;
    a.c = c;
    a.Wd = cfg.Wd;
    a.glv = cfg.glv ? 1 : 0;
    a.Wh = cfg.Wd / 2;
    a.W = cfg.G;
    a.NB = 1u << (c - 1);
    a.K = (uint32_t)a.W * a.NB;
    const uint64_t emax = (uint64_t)n * a.Wd;
    p.nscan = (a.K + SCAN_B - 1) / SCAN_B;

    size_t off = 0;
    auto take = [&](size_t bytes) { char *q = base + off; off += align_up(bytes, 256); return q; };
    a.scalars = (uint32_t *)take(n * NLIMB * 4);
    a.count = (uint32_t *)take((size_t)a.K * 4);
    a.offs = (uint32_t *)take(((size_t)a.K + 1) * 4);
    a.cursor = (uint32_t *)take((size_t)a.K * 4);
    p.bsum = (uint32_t *)take((size_t)p.nscan * 4);
    a.entries = (uint32_t *)take((size_t)emax * 4 + 4);
    // batched-affine accumulation: reference lists, sums of even rounds (region A: a piece starting at list entry
    // e puts sum j at (e >> 1) + j), of odd rounds (region B: ((e + bucket + share) >> 2) + j), the fix-up's own
    // small regions, pair list and codes (share t starts at (E0 >> 1) + t)
    const uint32_t U = shares_for<G>(sms, emax);
    const uint64_t capA = emax / 2 + 1, capB = (emax + a.K + U) / 4 + 2, capF = 2 * (uint64_t)U + 2;
    // bucket-reduction tree (bucket_tree.cuh): a reference and a scratch slot per node, 2 NB nodes per set
    TreeArgs &t = p.t;
    memset(&t, 0, sizeof t);
    t.W = (uint32_t)a.W;
    t.NB = a.NB;
    t.k = (uint32_t)(c - 1);
    t.h = t.k > 5u ? t.k - 5u : 0u;
    t.nodes = 2u * a.NB;
    // affine rounds while a round has enough additions to pay for its tile inversion (bucket_tree.cuh)
    t.hA = 0;
    while (t.hA < t.h && (uint64_t)a.W * (a.NB >> (t.hA + 2)) * (3u + t.hA) * G::F::DEG >= TREE_AFFINE_MIN) ++t.hA;
    t.jbase = t.hA < t.h ? a.NB - (a.NB >> t.hA) : t.nodes;
    t.jnodes = t.nodes - t.jbase;
    const uint64_t capT = (uint64_t)a.W * t.nodes;
    p.slots = capA + capB + capF + capT;
    b.K = a.K;
    b.U = U;
    b.offs = a.offs;
    b.refs[0] = a.entries;
    b.refs[1] = (uint32_t *)take((size_t)emax * 4 + 4);
    b.scratch = (uint32_t *)take((size_t)p.slots * AFFB);
    b.capA = (uint32_t)capA;
    b.fx_scratch_base = (uint32_t)(capA + capB);
    b.pairs = (uint4 *)take((size_t)(emax / 2 + U + 1) * 16);
    b.codes = (uint8_t *)take((size_t)(emax / 2 + U + 1));
    b.cntv = (uint32_t *)take(((size_t)a.K + U) * 4);
    b.bucket_ref = (uint32_t *)take((size_t)a.K * 4);
    b.bnd_ref = (uint32_t *)take((size_t)2 * U * 4);
    b.bnd_bucket = (uint32_t *)take((size_t)2 * U * 4);
    b.ctl = (uint32_t *)take((size_t)BA_CTL_WORDS * 4);
    for (int i = 0; i < 2; ++i) b.fx_refs[i] = (uint32_t *)take((size_t)2 * U * 4);
    b.fx_offs = (uint32_t *)take(((size_t)2 * U + 1) * 4);
    b.fx_cntv = (uint32_t *)take(((size_t)2 * U + 1) * 4);
    b.fx_bucket = (uint32_t *)take((size_t)2 * U * 4);
    t.slot_base = (uint32_t)(capA + capB + capF);
    t.bucket_ref = b.bucket_ref;
    t.R = (uint32_t *)take((size_t)capT * 4);
    t.codes = (uint8_t *)take((size_t)a.K);            // the largest round has 3 K / 4 additions
    p.fin = (uint32_t *)take((size_t)a.W * (t.k + 1) * JACB);
    t.J = (uint32_t *)take((size_t)a.W * t.jnodes * JACB);
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

// shared-memory attributes of every kernel of the group, once per context (they are per device)
template <class G>
int prepare_kernels(b200msm_ctx *ctx) {
    typedef TailCfg<G> TC;
    bool &done = ctx->kernels_ready[G::GROUP];
    if (done) return B200MSM_OK;
    int rc;
    if ((rc = set_smem(ctx, k_batch_add<G>, BaCfg<G>::TS::SMEM))) return rc;
    if ((rc = set_smem(ctx, k_ba_fixup<G>, BaCfg<G>::TS::SMEM))) return rc;
    if ((rc = set_smem(ctx, k_tree_round<G>, BaCfg<G>::TS::SMEM))) return rc;
    if ((rc = set_smem(ctx, k_tree_finish<G>, TC::TS::SMEM))) return rc;
    if ((rc = set_smem(ctx, k_tree_jac<G>, TC::TS::SMEM))) return rc;
    if ((rc = set_smem(ctx, k_sum<G>, TC::TS::SMEM))) return rc;
    if ((rc = set_smem(ctx, k_horner<G>, TC::TS::SMEM))) return rc;
    if ((rc = set_smem(ctx, k_to_affine<G>, TC::TS::SMEM))) return rc;
    if ((rc = set_smem(ctx, k_scalar_mul<G>, TC::TS::SMEM))) return rc;
    if ((rc = set_smem(ctx, k_synth_bases<G>, TC::TS::SMEM))) return rc;
    if ((rc = set_smem(ctx, k_batch_normalise<G>, TC::TS::SMEM))) return rc;
    if ((rc = set_smem(ctx, k_dbl_many<G>, DblCfg<G>::TS::SMEM))) return rc;
    if ((rc = set_smem(ctx, k_psi_many<G>, TC::TS::SMEM))) return rc;
    done = true;
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

// Size lane `li`'s arena for MSMs of n points over `bs` and set the kernels' shared-memory attributes now, so
// that the first MSM does not pay for them (b200msm_key_load warms every lane it is going to use).
template <class G>
int reserve_lane(b200msm_ctx *ctx, int li, const BaseSet &bs, size_t n) {
    int rc = prepare_kernels<G>(ctx);
    if (rc || n == 0) return rc;
    Plan probe = make_plan<G>(ctx->sm_count, n, cfg_for_set(ctx, bs, n), nullptr);   // all SMs: the most shares, the largest arena
    return grow_arena(ctx, ctx->lanes[li], probe.bytes);
}

// Enqueue one MSM on lane `li`.  scalars: host or device pointer (Montgomery Fr).
template <class G>
int enqueue_msm(b200msm_ctx *ctx, int li, const BaseSet &bs, size_t offset, const uint64_t *scalars, size_t n,
                uint64_t *out_xyz) {
    typedef typename G::F F;
    typedef TailCfg<G> TC;
    typedef BaCfg<G> BC;
    constexpr int DEG = F::DEG;
    constexpr size_t AFFW = 2 * DEG * NLIMB, JACW = 3 * DEG * NLIMB;
    Lane &ln = ctx->lanes[li];
    cudaStream_t st = ln.stream;
    ln.out_words = JACW / 2;
    ln.user_out = out_xyz;
    ln.ctl_valid = false;

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

    const TabCfg cfg = cfg_for_set(ctx, bs, n);
    const int c = cfg.c;
    // the SMs this lane may occupy (b200msm_set_lane_sms): persistent grids and the number of shares follow it
    const int sms = lane_sm_cap(ctx, li);
    // 32-bit positions in the sorted list and 30-bit table rows (reference = row | scratch << 30 | sign << 31)
    if ((uint64_t)n * (uint64_t)cfg.Wd >= (uint64_t(1) << 32) - 4096 || (uint64_t)bs.n * (uint64_t)cfg.NT >= (uint64_t(1) << 30))
        return fail(ctx, B200MSM_ERR_ARG, "n = %zu with %d digits per scalar exceeds the 2^32 sorted entries / 2^30 table rows of one call; shard the MSM", n, cfg.Wd);
    int rc = prepare_kernels<G>(ctx);
    if (rc) return rc;
    Plan probe = make_plan<G>(sms, n, cfg, nullptr);
    if (probe.slots >= (uint64_t(1) << 30))
        return fail(ctx, B200MSM_ERR_ARG, "n = %zu with %d digits per scalar needs %llu scratch points (limit 2^30 per call); shard the MSM", n, cfg.Wd, (unsigned long long)probe.slots);
    if ((rc = grow_arena(ctx, ln, probe.bytes))) return rc;
    Plan p = make_plan<G>(sms, n, cfg, ln.arena);
    MsmArgs &a = p.a;
    BaArgs &b = p.b;
    a.bases = bs.pts + offset * AFFW;
    a.base_inf = bs.inf + offset;
    a.tab_stride = (uint32_t)bs.n;
    b.bases = a.bases;

    uint64_t launches = 0;
    CU(cudaEventRecord(ln.ev[0], st));
    // Scalars: device-resident ones are copied in one piece; host scalars are uploaded in chunks on the lane's
    // copy stream so that the de-Montgomery + digit histogram of chunk k overlap the PCIe transfer of chunk k+1.
    cudaPointerAttributes pattr;
    const bool on_device = cudaPointerGetAttributes(&pattr, scalars) == cudaSuccess && pattr.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    const int nchunk = (!on_device && n >= (size_t(1) << 16)) ? NCOPY : 1;
    CU(cudaMemsetAsync(a.count, 0, (size_t)a.K * 4, st));
    if (nchunk > 1 && !ln.copy_stream) {      // created on first use: a lane that only sees device scalars never needs it
        CU(cudaStreamCreateWithFlags(&ln.copy_stream, cudaStreamNonBlocking));
        for (int e = 0; e <= NCOPY; ++e) CU(cudaEventCreateWithFlags(&ln.ev_copy[e], cudaEventDisableTiming));
    }
    if (nchunk > 1) {
        CU(cudaEventRecord(ln.ev_copy[NCOPY], st));                 // the arena is free once earlier work on st is done
        CU(cudaStreamWaitEvent(ln.copy_stream, ln.ev_copy[NCOPY], 0));
    }
    for (int ch = 0; ch < nchunk; ++ch) {
        const size_t lo = n * ch / nchunk, hi = n * (ch + 1) / nchunk;
        if (nchunk > 1) {
            CU(cudaMemcpyAsync(a.scalars + lo * NLIMB, (const uint32_t *)scalars + lo * NLIMB, (hi - lo) * NLIMB * 4, cudaMemcpyDefault, ln.copy_stream));
            CU(cudaEventRecord(ln.ev_copy[ch], ln.copy_stream));
            CU(cudaStreamWaitEvent(st, ln.ev_copy[ch], 0));
        } else {
            CU(cudaMemcpyAsync(a.scalars, scalars, n * NLIMB * 4, cudaMemcpyDefault, st));
            CU(cudaEventRecord(ln.ev[1], st));
        }
        k_from_mont<typename G::Fr><<<(unsigned)((hi - lo + 127) / 128), 128, 0, st>>>(a.scalars + lo * NLIMB, (uint32_t)(hi - lo));
        if (a.glv) { k_glv_split<G::CURVE><<<(unsigned)((hi - lo + 127) / 128), 128, 0, st>>>(a.scalars + lo * NLIMB, (uint32_t)(hi - lo)); ++launches; }
        a.i0 = (uint32_t)lo;
        a.i1 = (uint32_t)hi;
        k_count<<<(unsigned)((hi - lo + 255) / 256), 256, 0, st>>>(a);
        launches += 2;
    }
    if (nchunk > 1) CU(cudaEventRecord(ln.ev[1], st));    // chunked: "H2D" is folded into recode+sort
    const unsigned nb256 = (unsigned)((n + 255) / 256);
    k_scan_local<<<p.nscan, SCAN_T, 0, st>>>(a.count, a.offs, p.bsum, a.K);
    k_scan_bsum<<<1, SCAN_T, 0, st>>>(p.bsum, p.nscan, a.offs + a.K);
    k_scan_add<<<p.nscan, SCAN_T, 0, st>>>(a.offs, a.cursor, p.bsum, a.K);
    k_scatter<<<nb256, 256, 0, st>>>(a);
    launches += 4;
    CU(cudaEventRecord(ln.ev[2], st));

    // bucket accumulation: every team takes its share of the sorted list through all rounds in one launch
    CU(cudaMemsetAsync(b.ctl, 0, (size_t)BA_CTL_WORDS * 4, st));
    CU(cudaMemsetAsync(b.bucket_ref, 0xff, (size_t)a.K * 4, st));
    CU(cudaMemsetAsync(b.bnd_bucket, 0xff, (size_t)2 * b.U * 4, st));
    const unsigned ba_blocks = (unsigned)std::min<uint64_t>((uint64_t)sms, (uint64_t)b.U);   // shares are dealt round-robin over the blocks
    k_batch_add<G><<<ba_blocks, BC::TS::THREADS, BC::TS::SMEM, st>>>(b);
    k_ba_fixup<G><<<1, BC::TS::THREADS, BC::TS::SMEM, st>>>(b);
    launches += 2;
    CU(cudaEventRecord(ln.ev[3], st));

    // bucket reduction (bucket_tree.cuh): h rounds of pairwise affine additions, then the k + 1 short lists of every set
    {
        TreeArgs t = p.t;
        for (uint32_t r = 1; r <= t.h; ++r) {
            t.r = r;
            t.q = t.NB >> (r + 1);
            t.logq = t.k - (r + 1);
            t.P = t.W * t.q * (2u + r);
            if (r <= t.hA) {
                const uint64_t teams = std::min<uint64_t>((uint64_t)sms * BC::TPB, ((uint64_t)t.P + 31) / 32);
                const unsigned blocks = (unsigned)std::min<uint64_t>((uint64_t)sms, teams);
                k_tree_round<G><<<blocks, BC::TS::THREADS, BC::TS::SMEM, st>>>(b, t);
            } else {
                const unsigned lanes = TC::TPB * 32;
                k_tree_jac<G><<<(t.P + lanes - 1) / lanes, TC::TS::THREADS, TC::TS::SMEM, st>>>(b, t);
            }
            ++launches;
        }
        const unsigned lists = t.W * (t.k + 1u);
        k_tree_finish<G><<<(lists + TC::TPB - 1) / TC::TPB, TC::TS::THREADS, TC::TS::SMEM, st>>>(b, t, p.fin);
        k_sum<G><<<((unsigned)a.W + TC::TPB - 1) / TC::TPB, TC::TS::THREADS, TC::TS::SMEM, st>>>(p.fin, a.winsum, (uint32_t)a.W, t.k + 1u, 32u);
        launches += 2;
    }
    k_horner<G><<<1, TC::TS::THREADS, TC::TS::SMEM, st>>>(a);
    ++launches;
    CU(cudaEventRecord(ln.ev[4], st));
    CU(cudaMemcpyAsync(ln.h_result, a.result, JACW * 4, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(ln.h_ctl, b.ctl, (size_t)BA_CTL_WORDS * 4, cudaMemcpyDeviceToHost, st));
    ln.ctl_valid = true;
    CU(cudaEventRecord(ln.ev[5], st));
    CU(cudaGetLastError());

    ln.pending = true;
    ln.timed = true;
    ln.info[0] = (uint64_t)c;
    ln.info[1] = (uint64_t)a.Wd;
    ln.info[2] = (uint64_t)n * a.Wd;
    ln.info[3] = (uint64_t)b.U;
    ln.info[4] = launches;
    ln.info[5] = (uint64_t)a.W;
    ln.info[6] = (uint64_t)cfg.NT;
    ln.info[7] = 0;
    {
        auto o = [&](const void *q) { return (uint64_t)((const char *)q - ln.arena); };
        const uint64_t d[16] = {a.K, o(a.offs), o(b.refs[0]), o(b.refs[1]), o(b.scratch), o(b.bucket_ref), o(b.cntv), b.capA, (uint64_t)n * a.Wd, o(a.scalars), (uint64_t)a.W, a.NB, b.U, o(b.bnd_ref), o(b.bnd_bucket), o(b.pairs)};
        memcpy(ln.dbg, d, sizeof d);
    }
    return B200MSM_OK;
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
        case 8:   // Team::inv_lane0 (the tile inversion of the batched-affine accumulation) on every lane's element in turn
            for (int src = 0; src < 32; ++src) {
                T.copy_lane(3, 0, src);
                T.sync();
                if (team_any(!T.is_zero(3))) {
                    T.inv_lane0(4, 3, 5, 6);
                    T.copy_lane(5, 4, 0);
                } else T.set_zero(5);
                T.copy(2, 5, (threadIdx.x & 31) == src);
            }
            break;
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
    DevBuf ba, bb2, bo, bf;
    CU(cudaMalloc(&ba.p, n * ab));
    CU(cudaMalloc(&bo.p, n * ob));
    CU(cudaMemcpy(ba.p, a, n * ab, cudaMemcpyHostToDevice));
    if (b) { CU(cudaMalloc(&bb2.p, n * bb)); CU(cudaMemcpy(bb2.p, b, n * bb, cudaMemcpyHostToDevice)); }
    if (flags) { CU(cudaMalloc(&bf.p, n * 4)); CU(cudaMemcpy(bf.p, flags, n * 4, cudaMemcpyHostToDevice)); }
    uint32_t *da = (uint32_t *)ba.p, *db = (uint32_t *)bb2.p, *dout = (uint32_t *)bo.p, *dfl = (uint32_t *)bf.p;
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
    return B200MSM_OK;
}

// sum of n Jacobian points (the partial results of the shards) -> one Jacobian point, all on the device and on
// stream st: 32-ary butterfly levels (k_sum), then k_horner with one window for the (1, 1, 0) form of infinity.
// scratch: 2 * (ceil(n / 32) + 1) Jacobian points.
template <class G>
int fold_dev(b200msm_ctx *ctx, cudaStream_t st, const uint32_t *din, size_t n, uint32_t *scratch, uint32_t *dres) {
    typedef TailCfg<G> TC;
    constexpr size_t JACW = 3 * G::F::DEG * NLIMB;
    int rc = prepare_kernels<G>(ctx);
    if (rc) return rc;
    uint32_t *buf[2] = {scratch, scratch + ((n + 31) / 32 + 1) * JACW};
    const uint32_t *in = din;
    uint32_t nin = (uint32_t)n;
    int flip = 0;
    do {  // at least one pass so that the input is never aliased
        const uint32_t nout = (nin + 31) / 32;
        k_sum<G><<<(nout + TC::TPB - 1) / TC::TPB, TC::TS::THREADS, TC::TS::SMEM, st>>>(in, buf[flip], 1u, nin, 32u);
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
    k_horner<G><<<1, TC::TS::THREADS, TC::TS::SMEM, st>>>(a);
    CU(cudaGetLastError());
    return B200MSM_OK;
}

template <class G>
int to_affine_dev(b200msm_ctx *ctx, cudaStream_t st, size_t n, const uint32_t *din, uint32_t *dout) {
    typedef TailCfg<G> TC;
    int rc = prepare_kernels<G>(ctx);
    if (rc) return rc;
    const unsigned lanes = TC::TPB * 32;
    k_to_affine<G><<<(unsigned)((n + lanes - 1) / lanes), TC::TS::THREADS, TC::TS::SMEM, st>>>((uint32_t)n, din, dout);
    CU(cudaGetLastError());
    return B200MSM_OK;
}

// host-facing: points in from host or device memory, result to the host, through the context's tail buffers
template <class G>
int run_fold(b200msm_ctx *ctx, const uint64_t *xyz, size_t n, uint64_t *out) {
    constexpr size_t JACW = 3 * G::F::DEG * NLIMB, JACB = JACW * 4;
    int rc = tail_reserve(ctx, (n + fold_scratch_points(n) + 1) * JACB, JACB);
    if (rc) return rc;
    TailBuf &t = ctx->tail;
    uint32_t *din = (uint32_t *)t.d, *scratch = din + n * JACW, *dres = scratch + fold_scratch_points(n) * JACW;
    CU(cudaMemcpyAsync(din, xyz, n * JACB, cudaMemcpyDefault, t.st));
    if ((rc = fold_dev<G>(ctx, t.st, din, n, scratch, dres))) return rc;
    CU(cudaMemcpyAsync(t.h, dres, JACB, cudaMemcpyDeviceToHost, t.st));
    CU(cudaStreamSynchronize(t.st));
    memcpy(out, t.h, JACB);
    return B200MSM_OK;
}

// ---- affine normalisation / synthetic bases ---------------------------------------------------
template <class G>
int run_to_affine(b200msm_ctx *ctx, size_t n, const uint64_t *xyz, uint64_t *out) {
    constexpr size_t EB = G::F::DEG * NLIMB * 4;
    int rc = tail_reserve(ctx, n * 5 * EB, n * 2 * EB);
    if (rc) return rc;
    TailBuf &t = ctx->tail;
    uint32_t *din = (uint32_t *)t.d, *dout = din + n * 3 * (EB / 4);
    CU(cudaMemcpyAsync(din, xyz, n * 3 * EB, cudaMemcpyDefault, t.st));
    if ((rc = to_affine_dev<G>(ctx, t.st, n, din, dout))) return rc;
    cudaPointerAttributes pattr;
    const bool out_on_device = cudaPointerGetAttributes(&pattr, out) == cudaSuccess && pattr.type == cudaMemoryTypeDevice;
    cudaGetLastError();
    if (out_on_device) {
        CU(cudaMemcpyAsync(out, dout, n * 2 * EB, cudaMemcpyDeviceToDevice, t.st));
        CU(cudaStreamSynchronize(t.st));
    } else {
        CU(cudaMemcpyAsync(t.h, dout, n * 2 * EB, cudaMemcpyDeviceToHost, t.st));
        CU(cudaStreamSynchronize(t.st));
        memcpy(out, t.h, n * 2 * EB);
    }
    return B200MSM_OK;
}

template <class G>
const uint32_t *generator_words() {
    static const uint32_t c0g1[] = MNT753_GEN_C0_G1_U32, c0g2[] = MNT753_GEN_C0_G2_U32, c1g1[] = MNT753_GEN_C1_G1_U32,
                          c1g2[] = MNT753_GEN_C1_G2_U32;
    return G::CURVE == 0 ? (G::GROUP == 1 ? c0g1 : c0g2) : (G::GROUP == 1 ? c1g1 : c1g2);
}

// Build window tables 1 .. NT-1 of a base set whose table 0 (the points) is already in place:
// table t = 2^(c*G) * table t-1, each normalised back to affine with one inversion per 32-point run.
template <class G>
int build_tables(b200msm_ctx *ctx, BaseSet &bs) {
    typedef TailCfg<G> TC;
    constexpr size_t EB = G::F::DEG * NLIMB * 4;
    constexpr uint32_t B = 32;
    if (bs.NT <= 1 || bs.n == 0) return B200MSM_OK;
    DevBuf jac, pre;
    const size_t n = bs.n;
    CU(cudaMalloc(&jac.p, n * 3 * EB));
    CU(cudaMalloc(&pre.p, n * EB));
    CU(cudaFuncSetAttribute(k_dbl_many<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DblCfg<G>::TS::SMEM));
    CU(cudaFuncSetAttribute(k_batch_normalise<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC::TS::SMEM));
    EventPair ev;
    CU(cudaEventCreate(&ev.e0));
    CU(cudaEventCreate(&ev.e1));
    CU(cudaEventRecord(ev.e0, 0));
    const unsigned lanes = TC::TPB * 32;
    const size_t runs = (n + B - 1) / B;
    const size_t tabw = n * 2 * (EB / 4);
    // with split scalars (glv.cuh) only the lower half of the tables is built by doubling; table NT/2 + t = psi(table t)
    const int by_doubling = bs.glv ? bs.NT / 2 : bs.NT;
    for (int t = 1; t < by_doubling; ++t) {
        const unsigned dl = DblCfg<G>::TPB * 32;
        k_dbl_many<G><<<(unsigned)((n + dl - 1) / dl), DblCfg<G>::TS::THREADS, DblCfg<G>::TS::SMEM>>>((uint32_t)n, bs.c_tab * bs.G, bs.pts + (t - 1) * tabw,
                                                                                                    (uint32_t *)jac.p);
        k_batch_normalise<G><<<(unsigned)((runs + lanes - 1) / lanes), TC::TS::THREADS, TC::TS::SMEM>>>(
            (uint32_t)n, B, (const uint32_t *)jac.p, (uint32_t *)pre.p, bs.pts + t * tabw);
    }
    if (bs.glv) {
        int rc = prepare_kernels<G>(ctx);
        if (rc) return rc;
        for (int t = by_doubling; t < bs.NT; ++t)
            k_psi_many<G><<<(unsigned)((n + lanes - 1) / lanes), TC::TS::THREADS, TC::TS::SMEM>>>((uint32_t)n, bs.pts + (size_t)(t - by_doubling) * tabw, bs.pts + (size_t)t * tabw);
    }
    CU(cudaEventRecord(ev.e1, 0));
    CU(cudaGetLastError());
    CU(cudaEventSynchronize(ev.e1));
    CU(cudaEventElapsedTime(&bs.build_ms, ev.e0, ev.e1));
    return B200MSM_OK;
}

// bases[i] = k_p0 * G + i * (k_q * G), written straight into a new resident base set
template <class G>
int run_synthetic(b200msm_ctx *ctx, size_t n, const uint64_t *k_p0, const uint64_t *k_q, BaseSet &bs) {
    typedef TailCfg<G> TC;
    constexpr size_t EB = G::F::DEG * NLIMB * 4;
    constexpr uint32_t B = 64;
    DevBuf gen, ks, pj, pa, jac, pre;
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
    k_to_affine<G><<<1, TC::TS::THREADS, TC::TS::SMEM>>>(2u, (const uint32_t *)pj.p, (uint32_t *)pa.p);
    const unsigned lanes = TC::TPB * 32;
    const size_t runs = (n + B - 1) / B;
    k_synth_bases<G><<<(unsigned)((runs + lanes - 1) / lanes), TC::TS::THREADS, TC::TS::SMEM>>>(
        (uint32_t)n, B, (const uint32_t *)pa.p, (const uint32_t *)pa.p + 2 * (EB / 4), bs.pts, (uint32_t *)jac.p, (uint32_t *)pre.p);
    constexpr int DEG = G::F::DEG;
    k_flag_inf<DEG><<<(unsigned)((n + 255) / 256), 256>>>(bs.pts, (uint32_t)n, bs.inf);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    return B200MSM_OK;
}

// out (Jacobian) = k * P for one affine point P and one Montgomery-form scalar k of Fr (the r * Bt1 of the proof
// assembly, cuda_prover_piecewise.cu:198; the reference does it on the host with libff)
template <class G>
int run_scalar_mul(b200msm_ctx *ctx, const uint64_t *affine, const uint64_t *k_mont, uint64_t *out_xyz) {
    typedef TailCfg<G> TC;
    constexpr size_t EB = G::F::DEG * NLIMB * 4;
    DevBuf p, k, o;
    CU(cudaMalloc(&p.p, 2 * EB));
    CU(cudaMalloc(&k.p, NLIMB * 4));
    CU(cudaMalloc(&o.p, 3 * EB));
    CU(cudaMemcpy(p.p, affine, 2 * EB, cudaMemcpyDefault));
    CU(cudaMemcpy(k.p, k_mont, NLIMB * 4, cudaMemcpyDefault));
    CU(cudaFuncSetAttribute(k_scalar_mul<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC::TS::SMEM));
    k_scalar_mul<G><<<1, TC::TS::THREADS, TC::TS::SMEM>>>((const uint32_t *)p.p, (const uint32_t *)k.p, (uint32_t *)o.p);
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(out_xyz, o.p, 3 * EB, cudaMemcpyDefault));
    return B200MSM_OK;
}

// throughput of the slab multiplier (Team::mul, operands in shared memory) at 4 / 8 / 12 warps per SM:
// `iters` x 2 dependent products per lane.  Returns 10^9 Fq-tower products per second in *gops.
template <class G>
__global__ void __launch_bounds__(BaCfg<G>::TS::THREADS, BaCfg<G>::MINB) k_mb_teammul(int iters, uint32_t *out) {
    typedef typename G::F F;
    typedef BaCfg<G> C;
    extern __shared__ uint4 smem[];
    __shared__ uint32_t s_flags[C::TPB][4];
    int team;
    const Team<F> T = C::TS::make(smem, s_flags, team);
    T.set_one(0);
    T.set_one(1);
    T.dbl(1, 1);
    T.sync();
    for (int it = 0; it < iters; ++it) { T.mul(0, 0, 1); T.mul(1, 1, 0); }
    T.sync();
    if (T.is_zero(0) && out) out[0] = 1;
}

template <class G>
int run_teammul_bench(b200msm_ctx *ctx, int blocks_per_sm, int iters, double *gops) {
    typedef BaCfg<G> C;
    CU(cudaFuncSetAttribute(k_mb_teammul<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::TS::SMEM));
    EventPair ev;
    CU(cudaEventCreate(&ev.e0));
    CU(cudaEventCreate(&ev.e1));
    const int blocks = ctx->sm_count * blocks_per_sm;
    for (int rep = 0; rep < 2; ++rep) {
        CU(cudaEventRecord(ev.e0, 0));
        k_mb_teammul<G><<<blocks, C::TS::THREADS, C::TS::SMEM>>>(iters, nullptr);
        CU(cudaEventRecord(ev.e1, 0));
        CU(cudaEventSynchronize(ev.e1));
    }
    float ms = 0;
    CU(cudaEventElapsedTime(&ms, ev.e0, ev.e1));
    CU(cudaGetLastError());
    *gops = double(blocks) * C::TPB * 32 * 2.0 * iters / (double(ms) * 1e6);
    return B200MSM_OK;
}

}  // namespace


template <class G>
constexpr GroupOps make_group_ops() {
    return GroupOps{&enqueue_msm<G>, &run_test<G>, &run_fold<G>, &run_to_affine<G>, &fold_dev<G>, &to_affine_dev<G>, &run_synthetic<G>, &build_tables<G>, &run_teammul_bench<G>, &run_scalar_mul<G>, &reserve_lane<G>};
}
