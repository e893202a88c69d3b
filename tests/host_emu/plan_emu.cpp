// CPU model check of the round planning of the batched-affine bucket accumulation: the REAL csrc/ba_plan.cuh
// (ba_boundary, ba_plan with its warp shuffles, the scratch-slot formulas) runs on a simulated warp (warp_sim.hpp);
// the additions themselves are replaced by set unions of entry ids.  What the model checks:
//   * every source a pass reads holds the point it is supposed to hold (a scratch slot is never rewritten -- by a
//     prefix product of the forward pass or by a later sum -- while a reference to it is still going to be read);
//   * every sum and prefix lands inside the region the host sized for it (capA / capB / the fix-up's regions: the
//     formulas of make_plan, csrc/group_ops.cuh), and inside the part of it that belongs to the share;
//   * after all rounds every bucket is exactly the union of its entries -- whole buckets through bucket_ref, buckets
//     cut between shares through the boundary list and the fix-up's rounds.
// The share loop, the tile walk and the fix-up's compaction are restated here from k_batch_add / ba_share / ba_tile /
// k_ba_fixup (batch_affine.cuh); the planning they call is the shipped code.  A second model, at the end of the file,
// does the same for the index logic of the bucket-reduction tree (csrc/tree_plan.cuh).  Test-only.
#include "warp_sim.hpp"

#include <algorithm>
#include <map>
#include <string>

#include "../../gpu_groth16_prover_3x_b200/csrc/prim.cuh"      // uint4 and the carry-chain primitives of the host emulation
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }
static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }

#include "../../gpu_groth16_prover_3x_b200/csrc/ba_plan.cuh"
#include "../../gpu_groth16_prover_3x_b200/csrc/tree_plan.cuh"
#include "../../gpu_groth16_prover_3x_b200/csrc/recode.cuh"
#include "../../gpu_groth16_prover_3x_b200/csrc/glv_split.cuh"

using namespace mnt753;

namespace {

typedef std::vector<uint32_t> Ids;      // sorted entry ids = the "point"

struct Model {
    std::string err;
    std::map<uint32_t, Ids> slot;        // scratch slot -> point; an empty vector marks a parked prefix product
    std::map<uint32_t, int> owner;       // scratch slot -> share that wrote it (-1: the fix-up)
    uint64_t capA = 0, capB = 0, capF = 0;
    uint32_t cancel_every = 0, additions = 0;   // every cancel_every-th addition sums to infinity (P + (-P)) ...
    Ids lost;                                   // ... and its entries leave their bucket
    bool fail(const std::string &m) { if (err.empty()) err = m; return false; }

    bool read(uint32_t ref, Ids &out) {
        if (ref == REF_INF) return fail("an infinite operand reached a pass");
        if (!(ref & REF_SCRATCH)) { out.assign(1, ref & REF_IDX); return true; }      // a table row: entry id = row
        auto it = slot.find(ref & REF_IDX);
        if (it == slot.end()) return fail("read of scratch slot " + std::to_string(ref & REF_IDX) + " that was never written");
        if (it->second.empty()) return fail("read of scratch slot " + std::to_string(ref & REF_IDX) + " while it holds a prefix product");
        out = it->second;
        return true;
    }
    bool write(uint32_t s, const Ids &v, int share, uint64_t lo, uint64_t hi) {
        if (s < lo || s >= hi) return fail("scratch slot " + std::to_string(s) + " outside [" + std::to_string(lo) + ", " + std::to_string(hi) + ") of its region");
        auto it = owner.find(s);
        if (it != owner.end() && it->second != share) return fail("scratch slot " + std::to_string(s) + " written by two shares");
        owner[s] = share;
        slot[s] = v;
        return true;
    }
};

// one share through all its rounds (ba_share + ba_tile of batch_affine.cuh, with unions for additions)
bool run_share(Model &M, const BaArgs &a, const BaView &v, int share, uint32_t bmax, uint32_t &rounds_out) {
    const uint64_t regA_lo = share < 0 ? a.fx_scratch_base : 0, regA_hi = share < 0 ? a.fx_scratch_base + a.U + 1 : M.capA;
    const uint64_t regB_lo = share < 0 ? a.fx_scratch_base + a.U + 1 : M.capA, regB_hi = share < 0 ? a.fx_scratch_base + M.capF : M.capA + M.capB;
    uint32_t r = 0;
    for (;; ++r) {
        uint32_t P[warpsim::LANES];
        warpsim::run_warp([&](int lane) {
            uint32_t maxc = 0;
            P[lane] = ba_plan(v, r, lane, maxc);
        });
        for (int l = 1; l < warpsim::LANES; ++l)
            if (P[l] != P[0]) return M.fail("ba_plan returned different pair counts on different lanes");
        if (P[0] == 0) break;
        if (r > 40) return M.fail("rounds do not terminate");
        uint32_t *nxt = v.refs[(r + 1u) & 1u];
        const uint64_t lo = (r & 1u) ? regB_lo : regA_lo, hi = (r & 1u) ? regB_hi : regA_hi;
        uint32_t p0 = 0;
        while (p0 < P[0]) {
            const uint32_t left = (P[0] - p0 + 31u) / 32u;
            uint32_t B = left;
            if (left > bmax) B = left >= 2u * bmax ? bmax : (left + 1u) / 2u;
            const uint4 idle = make_uint4(REF_INF, REF_INF, 0u, 0u);
            // forward: abscissae of both operands are read, the prefix is parked in the output slot
            for (uint32_t i = 0; i < B; ++i)
                for (uint32_t lane = 0; lane < 32; ++lane) {
                    const uint32_t p = p0 + i * 32u + lane;
                    const uint4 d = p < P[0] ? v.pairs[p] : idle;
                    if (d.x == REF_INF) continue;
                    Ids t;
                    if (!M.read(d.x, t)) return false;
                    if (d.y != REF_INF && !M.read(d.y, t)) return false;
                    if (!M.write(d.z, Ids(), share, lo, hi)) return false;
                }
            // backward: both operands and the prefix are read, the sum replaces the prefix
            for (int i = (int)B - 1; i >= 0; --i)
                for (uint32_t lane = 0; lane < 32; ++lane) {
                    const uint32_t p = p0 + (uint32_t)i * 32u + lane;
                    const uint4 d = p < P[0] ? v.pairs[p] : idle;
                    if (d.x == REF_INF) continue;
                    Ids s1, s2;
                    if (!M.read(d.x, s1)) return false;
                    if (d.y != REF_INF && !M.read(d.y, s2)) return false;
                    auto it = M.slot.find(d.z);
                    if (it == M.slot.end() || !it->second.empty()) return M.fail("the prefix parked in slot " + std::to_string(d.z) + " was overwritten before the backward pass");
                    Ids sum(s1.size() + s2.size());
                    std::merge(s1.begin(), s1.end(), s2.begin(), s2.end(), sum.begin());
                    const bool cancel = d.y != REF_INF && M.cancel_every && ++M.additions % M.cancel_every == 0;
                    if (cancel) {           // the kernel writes zeros to the slot and hands infinity on
                        M.lost.insert(M.lost.end(), sum.begin(), sum.end());
                        sum.assign(1, 0xfffffffeu);
                    }
                    if (!M.write(d.z, sum, share, lo, hi)) return false;
                    nxt[d.w] = cancel ? REF_INF : (REF_SCRATCH | d.z);
                }
            p0 += 32u * B;
        }
    }
    rounds_out = r;
    return true;
}

}  // namespace

extern "C" {

// counts[K]: points per bucket (the sorted list is 0, 1, 2, ... grouped by bucket), U shares, bmax additions per lane
// and tile, every cancel_every-th addition (0: none) sums to infinity.  Returns 0 when the model holds, 1 otherwise (message in err, at most errlen bytes).  stats[0] = most
// rounds of a share, [1] = buckets cut between shares, [2] = scratch slots written.
int emu_plan_check(uint32_t K, const uint32_t *counts, uint32_t U, uint32_t bmax, uint32_t cancel_every, uint64_t *stats, char *err, size_t errlen) {
    std::vector<uint32_t> offs(K + 1, 0);
    for (uint32_t b = 0; b < K; ++b) offs[b + 1] = offs[b] + counts[b];
    const uint32_t E = offs[K];
    if (E >= (1u << 30)) return 2;
    Model M;
    M.cancel_every = cancel_every;
    // capacities as the host sizes them (make_plan, csrc/group_ops.cuh): emax = E here
    M.capA = (uint64_t)E / 2 + 1;
    M.capB = ((uint64_t)E + K + U) / 4 + 2;
    M.capF = 2 * (uint64_t)U + 2;
    std::vector<uint32_t> refs0(E + 1), refs1(E + 1, 0), cntv(K + U, 0), bucket_ref(K, REF_INF), bnd_ref(2 * U, REF_INF), bnd_bucket(2 * U, REF_INF);
    std::vector<uint4> pairs((size_t)E / 2 + U + 1);
    for (uint32_t e = 0; e < E; ++e) refs0[e] = e;
    BaArgs a;
    memset(&a, 0, sizeof a);
    a.K = K;
    a.U = U;
    a.offs = offs.data();
    a.refs[0] = refs0.data();
    a.refs[1] = refs1.data();
    a.capA = (uint32_t)M.capA;
    a.pairs = pairs.data();
    a.cntv = cntv.data();
    a.bucket_ref = bucket_ref.data();
    a.bnd_ref = bnd_ref.data();
    a.bnd_bucket = bnd_bucket.data();
    a.fx_scratch_base = (uint32_t)(M.capA + M.capB);
    a.T = std::max<uint32_t>(1u, (E + U - 1u) / U);
    uint32_t max_rounds = 0, cut = 0;
    bool ok = true;
    // ---- k_batch_add: every share through its rounds, then what is left of its pieces
    for (uint32_t t = 0; t < U && ok; ++t) {
        const uint32_t E0 = ba_boundary(a, t, E), E1 = ba_boundary(a, t + 1u, E);
        if (E0 > E1) { ok = M.fail("share boundaries are not monotone"); break; }
        if (E0 >= E1) continue;
        BaView v;
        v.offs = a.offs;
        v.E0 = E0;
        v.E1 = E1;
        v.b0 = bucket_of(a.offs, a.K, E0);
        const uint32_t b1 = bucket_of(a.offs, a.K, E1 - 1u);
        v.npieces = b1 - v.b0 + 1u;
        v.id = t;
        v.refs[0] = a.refs[0];
        v.refs[1] = a.refs[1];
        v.cntv = a.cntv;
        v.pairs = a.pairs + (E0 >> 1) + t;
        v.codes = nullptr;
        v.a_base = 0u;
        v.b_base = a.capA;
        uint32_t rounds = 0;
        ok = run_share(M, a, v, (int)t, bmax, rounds);
        if (!ok) break;
        max_rounds = std::max(max_rounds, rounds);
        const uint32_t *cur = a.refs[rounds & 1u];
        for (uint32_t q = 0; q < v.npieces; ++q) {
            const uint32_t b = v.b0 + q;
            const uint32_t lo = a.offs[b], hi = a.offs[b + 1];
            const uint32_t ps = std::max(lo, v.E0), pe = std::min(hi, v.E1);
            const uint32_t c = rounds ? a.cntv[b + t] : pe - ps;
            const uint32_t ref = c ? cur[ps] : REF_INF;
            if (c > 1u) { ok = M.fail("a piece is left with more than one point"); break; }
            if (lo >= v.E0 && hi <= v.E1) a.bucket_ref[b] = ref;
            else {
                const uint32_t s = 2u * t + (lo < v.E0 ? 0u : 1u);
                if (a.bnd_bucket[s] != REF_INF) { ok = M.fail("two cut pieces in one boundary slot"); break; }
                a.bnd_ref[s] = ref;
                a.bnd_bucket[s] = b;
            }
        }
    }
    // the shares must tile the list: every entry belongs to exactly one of them (monotone boundaries from 0 to E)
    if (ok && (ba_boundary(a, 0, E) != 0u || ba_boundary(a, U, E) != E)) ok = M.fail("the shares do not cover the list");
    // ---- k_ba_fixup: the cut pieces as a list of their own (slot order = list order), reduced by the same rounds
    if (ok) {
        std::vector<uint32_t> fr0(2 * U + 1, REF_INF), fr1(2 * U + 1, REF_INF), foffs, fbucket, fcntv(2 * U + 1, 0);
        uint32_t n = 0, prev = REF_INF;
        for (uint32_t i = 0; i < 2 * U; ++i) {
            if (a.bnd_bucket[i] == REF_INF) continue;
            if (a.bnd_bucket[i] != prev) { foffs.push_back(n); fbucket.push_back(a.bnd_bucket[i]); prev = a.bnd_bucket[i]; }
            fr0[n++] = a.bnd_ref[i];
        }
        foffs.push_back(n);
        cut = (uint32_t)fbucket.size();
        if (n) {
            std::vector<uint4> fpairs((size_t)n / 2 + 2);
            BaView v;
            v.offs = foffs.data();
            v.E0 = 0u;
            v.E1 = n;
            v.b0 = 0u;
            v.npieces = cut;
            v.id = 0u;
            v.refs[0] = fr0.data();
            v.refs[1] = fr1.data();
            v.cntv = fcntv.data();
            v.pairs = fpairs.data();
            v.codes = nullptr;
            v.a_base = a.fx_scratch_base;
            v.b_base = a.fx_scratch_base + a.U + 1u;
            uint32_t rounds = 0;
            ok = run_share(M, a, v, -1, bmax, rounds);
            if (ok) {
                const uint32_t *cur = rounds & 1u ? fr1.data() : fr0.data();
                for (uint32_t q = 0; q < cut; ++q) {
                    const uint32_t ps = foffs[q];
                    const uint32_t c = rounds ? fcntv[q] : foffs[q + 1] - ps;
                    if (a.bucket_ref[fbucket[q]] != REF_INF) { ok = M.fail("a cut bucket already has a final reference"); break; }
                    a.bucket_ref[fbucket[q]] = c ? cur[ps] : REF_INF;
                }
            }
        }
    }
    // ---- every bucket is the union of its entries (minus what cancelled)
    std::sort(M.lost.begin(), M.lost.end());
    for (uint32_t b = 0; b < K && ok; ++b) {
        const uint32_t ref = a.bucket_ref[b];
        Ids want;
        for (uint32_t i = 0; i < counts[b]; ++i)
            if (!std::binary_search(M.lost.begin(), M.lost.end(), offs[b] + i)) want.push_back(offs[b] + i);
        if (want.empty()) { if (ref != REF_INF) ok = M.fail("bucket " + std::to_string(b) + " should be empty"); continue; }
        Ids got;
        if (ref == REF_INF) { ok = M.fail("bucket " + std::to_string(b) + " lost its points"); break; }
        if (!M.read(ref, got)) { ok = false; break; }
        if (got != want) ok = M.fail("bucket " + std::to_string(b) + " is not the sum of its entries");
    }
    if (stats) { stats[0] = max_rounds; stats[1] = cut; stats[2] = M.slot.size(); }
    if (err && errlen) { strncpy(err, M.err.c_str(), errlen - 1); err[errlen - 1] = 0; }
    return ok ? 0 : 1;
}

// ---- the bucket-reduction tree (csrc/tree_plan.cuh, bucket_tree.cuh) with integers modulo 2^61 - 1 for points ----------
// val[W * 2^k]: the bucket sums left by the accumulation, 0 = empty bucket.  Rounds 1 .. h run TreePairs::passthrough /
// get / locate (the shipped index code) as k_tree_round (r <= hA) or k_tree_jac (r > hA) do; the k + 1 short lists and
// the window sum are restated from k_tree_finish / k_sum.  Checked: every input of a round was written by an EARLIER
// round (or is a bucket), every node is written exactly once, Jacobian nodes fall inside the J array the host sized
// (make_plan), and the result of every set is sum_b (b + 1) B_b.  Returns 0 / 1 (message in err).
int emu_tree_check(uint32_t W, uint32_t k, uint32_t hA_in, const uint64_t *val, char *err, size_t errlen) {
    const uint64_t MOD = (uint64_t(1) << 61) - 1;
    auto addm = [&](uint64_t a, uint64_t b) { const uint64_t s = a + b; return s >= MOD ? s - MOD : s; };
    std::string msg;
    auto fail = [&](const std::string &m) { if (msg.empty()) msg = m; return false; };
    TreeArgs t;
    memset(&t, 0, sizeof t);
    t.W = W;
    t.k = k;
    t.NB = 1u << k;
    t.h = k > 5u ? k - 5u : 0u;
    t.hA = std::min(hA_in, t.h);
    t.nodes = 2u * t.NB;
    t.jbase = t.hA < t.h ? t.NB - (t.NB >> t.hA) : t.nodes;
    t.jnodes = t.nodes - t.jbase;
    t.slot_base = 1000u;
    std::vector<uint32_t> bucket_ref((size_t)W * t.NB), R((size_t)W * t.nodes, 0xdeadbeefu);
    std::vector<int> written((size_t)W * t.nodes, -1);           // round that wrote the node
    std::vector<uint64_t> nodeval((size_t)W * t.nodes, 0), jval((size_t)W * t.jnodes + 1, 0);
    for (size_t i = 0; i < bucket_ref.size(); ++i) bucket_ref[i] = val[i] ? (uint32_t)i : REF_INF;     // a table row per bucket
    t.bucket_ref = bucket_ref.data();
    t.R = R.data();
    bool ok = true;
    int round = 0;
    // value behind a reference; inputs must come from an earlier round
    auto value = [&](uint32_t ref, uint32_t set, uint64_t &out) -> bool {
        if (ref == REF_INF) { out = 0; return true; }
        if ((ref & REF_JAC) == REF_JAC) {
            const uint32_t node = ref & REF_IDX;
            const long long j = (long long)node - (long long)set * t.nodes - (long long)t.jbase;
            if (j < 0 || j >= (long long)t.jnodes) return fail("Jacobian node " + std::to_string(node) + " outside the J array");
            if (written[node] < 0 || written[node] >= round) return fail("Jacobian node read before it was written");
            out = jval[(size_t)set * t.jnodes + (size_t)j];
            return true;
        }
        if (ref & REF_SCRATCH) {
            const uint32_t slot = ref & REF_IDX;
            if (slot < t.slot_base || slot - t.slot_base >= W * t.nodes) return fail("scratch reference outside the tree's slots");
            const uint32_t node = slot - t.slot_base;
            if (written[node] < 0 || written[node] >= round) return fail("node " + std::to_string(node) + " read in the round that writes it (or never written)");
            out = nodeval[node];
            return true;
        }
        if (ref >= bucket_ref.size()) return fail("table reference out of range");
        out = val[ref];
        return true;
    };
    auto own_set = [&](const uint32_t *ptr) -> bool {        // an input pointer must lie in bucket_ref or R
        return (ptr >= bucket_ref.data() && ptr < bucket_ref.data() + bucket_ref.size()) || (ptr >= R.data() && ptr < R.data() + R.size());
    };
    for (uint32_t r = 1; r <= t.h && ok; ++r) {
        round = (int)r;
        t.r = r;
        t.q = t.NB >> (r + 1);
        t.logq = t.k - (r + 1);
        t.P = t.W * t.q * (2u + r);
        const TreePairs src{t, nullptr, t.P};
        std::vector<std::pair<uint32_t, uint32_t>> wr;       // (node, reference) written this round, applied afterwards
        std::vector<std::pair<uint32_t, uint64_t>> wv;
        for (uint32_t p = 0; p < t.P && ok; ++p) {
            const uint32_t *i0, *i1;
            uint32_t node;
            src.locate(p, i0, i1, node);
            if (!own_set(i0) || !own_set(i1)) { ok = fail("an input of round " + std::to_string(r) + " lies outside the reference arrays"); break; }
            if (node >= W * t.nodes) { ok = fail("node out of range"); break; }
            if (written[node] >= 0) { ok = fail("node " + std::to_string(node) + " written twice"); break; }
            const uint32_t set = node / t.nodes;
            const uint32_t r0 = *i0, r1 = *i1;
            uint64_t v0, v1;
            if (!value(r0, set, v0) || !value(r1, set, v1)) { ok = false; break; }
            if (r <= t.hA) {
                // k_tree_round: passthrough for an empty operand, an affine addition otherwise
                const uint4 d = src.get(p);
                if (r0 == REF_INF || r1 == REF_INF) {
                    if (d.x != REF_INF) { ok = fail("get() returned an addition for an empty operand"); break; }
                    wr.push_back({node, r0 == REF_INF ? r1 : r0});
                } else {
                    if (d.x != r0 || d.y != r1 || d.w != node || d.z != t.slot_base + node) { ok = fail("get() and locate() disagree"); break; }
                    wr.push_back({node, REF_SCRATCH | d.z});
                    wv.push_back({node, addm(v0, v1)});
                }
            } else {
                // k_tree_jac: always a Jacobian node (infinity when both operands are empty)
                const long long j = (long long)node - (long long)set * t.nodes - (long long)t.jbase;
                if (j < 0 || j >= (long long)t.jnodes) { ok = fail("round " + std::to_string(r) + ": Jacobian node " + std::to_string(node) + " outside the J array"); break; }
                jval[(size_t)set * t.jnodes + (size_t)j] = addm(v0, v1);
                wr.push_back({node, REF_JAC | node});
            }
            written[node] = (int)r;
        }
        for (auto &w : wr) R[w.first] = w.second;
        for (auto &w : wv) nodeval[w.first] = w.second;
    }
    // ---- k_tree_finish + k_sum: the k + 1 short lists of every set
    round = (int)t.h + 1;
    for (uint32_t set = 0; set < W && ok; ++set) {
        uint64_t total = 0;
        const uint32_t nt = t.NB >> t.h;
        for (uint32_t l = 0; l <= t.k && ok; ++l) {
            uint64_t term = 0;
            for (uint32_t lane = 0; lane < 32u && ok; ++lane) {
                uint32_t ref = REF_INF;
                if (l == 0u) { if (lane < nt) ref = tree_tlist(t, set, t.h)[lane]; }
                else if (l - 1u < t.h) { if (lane < (nt >> 1)) ref = t.R[(size_t)set * t.nodes + tree_ooff(t, l - 1u, t.h - (l - 1u)) + lane]; }
                else if (lane < nt && ((lane >> (l - 1u - t.h)) & 1u)) ref = tree_tlist(t, set, t.h)[lane];
                uint64_t v;
                if (!value(ref, set, v)) { ok = false; break; }
                term = addm(term, v);
            }
            for (uint32_t i = 1; i < l; ++i) term = addm(term, term);          // l = 1 + j: j doublings
            total = addm(total, term);
        }
        if (!ok) break;
        uint64_t want = 0;
        for (uint32_t b = 0; b < t.NB; ++b) want = (uint64_t)(((unsigned __int128)want + (unsigned __int128)(b + 1u) * val[(size_t)set * t.NB + b]) % MOD);
        if (total != want) ok = fail("set " + std::to_string(set) + ": the tree does not sum to sum_b (b + 1) B_b");
    }
    if (err && errlen) { strncpy(err, msg.c_str(), errlen - 1); err[errlen - 1] = 0; }
    return ok ? 0 : 1;
}

// ---- signed-digit recoding (csrc/recode.cuh): the W digits of the nl-limb integer k for window width c ----------------
void emu_recode(const uint32_t *k, int nl, int c, int W, int32_t *digits) {
    for (int w = 0; w < W; ++w) digits[w] = 0;
    for_each_digit(k, nl, c, W, [&](int w, int d) { digits[w] = d; });
}

// ---- split of a G2 scalar (csrc/glv_split.cuh): k (24 limbs, plain integer) -> the two stored halves of 12 limbs each,
// |k_h| with its sign in bit 31 of the top limb -- exactly what k_glv_split leaves for the digit kernels
void emu_glv_split(int curve, const uint32_t *k_in, uint32_t *out) {
    uint32_t k[NLIMB], r[2][GLV_L];
    memcpy(k, k_in, sizeof k);
    if (curve == 0) glv_split<0>(k, r); else glv_split<1>(k, r);
    for (int h = 0; h < 2; ++h) memcpy(out + 12 * h, r[h], 48);
    // limbs 12, 13 of a half must be zero once the sign has been folded into limb 11 (|k_h| < 2^377)
    out[24] = r[0][12] | r[0][13] | r[1][12] | r[1][13];
}
}
