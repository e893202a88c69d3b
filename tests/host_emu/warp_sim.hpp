// Lock-step simulation of ONE CUDA warp on the host, for the warp-cooperative device code (shuffles, ballots).
//
// The 32 lanes are user-level contexts (ucontext) scheduled round-robin by one host thread: a lane runs until its
// next warp intrinsic, publishes its operand and yields; when it is resumed every lane has published, and it reads
// what it needs.  Operand slots are double-buffered by the parity of the intrinsic's sequence number, so one yield
// per intrinsic is enough.  Every intrinsic carries a tag (kind and sequence number) and the simulator aborts when the
// lanes of the warp disagree on it -- the "must be called by all lanes of a converged warp" contract of the device
// code is checked, not assumed.  Test-only; never shipped.
#pragma once
#include <ucontext.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

namespace warpsim {

constexpr int LANES = 32;

struct Warp {
    ucontext_t main;
    ucontext_t ctx[LANES];
    std::vector<char> stack[LANES];
    bool done[LANES];
    int cur = 0;
    uint64_t slot[2][LANES];
    uint32_t tag[2][LANES];
    uint32_t seq[LANES];
    std::function<void(int)> body;
};

inline Warp *&current() {
    static thread_local Warp *w = nullptr;
    return w;
}

inline int lane_id() { return current()->cur; }

// publish `raw` under `kind`, wait for the rest of the warp, return the buffer index to read from
inline int collective(uint64_t raw, uint32_t kind) {
    Warp *w = current();
    const int l = w->cur;
    const uint32_t s = w->seq[l]++;
    const int buf = (int)(s & 1u);
    w->slot[buf][l] = raw;
    w->tag[buf][l] = (kind << 24) ^ (s & 0xffffffu);
    swapcontext(&w->ctx[l], &w->main);
    for (int i = 0; i < LANES; ++i)
        if (w->tag[buf][i] != w->tag[buf][l]) {
            fprintf(stderr, "warpsim: lanes %d and %d disagree on warp intrinsic #%u (diverged warp)\n", l, i, s);
            abort();
        }
    return buf;
}

template <class T>
inline T exchange(T v, int src, uint32_t kind) {
    static_assert(sizeof(T) <= 8, "operand too wide");
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    const int buf = collective(raw, kind);
    T out;
    memcpy(&out, &current()->slot[buf][src & (LANES - 1)], sizeof(T));
    return out;
}

inline unsigned ballot(bool pred) {
    const int buf = collective(pred ? 1u : 0u, 7);
    unsigned m = 0;
    for (int i = 0; i < LANES; ++i) m |= (unsigned)(current()->slot[buf][i] & 1u) << i;
    return m;
}

inline void trampoline(int lane) {
    Warp *w = current();
    w->body(lane);
    w->done[lane] = true;
    swapcontext(&w->ctx[lane], &w->main);
}

// run body(lane) on the 32 lanes of one warp in lock step
inline void run_warp(const std::function<void(int)> &body) {
    Warp w;
    w.body = body;
    Warp *saved = current();
    current() = &w;
    for (int l = 0; l < LANES; ++l) {
        w.stack[l].resize(256 << 10);
        w.done[l] = false;
        w.seq[l] = 0;
        getcontext(&w.ctx[l]);
        w.ctx[l].uc_stack.ss_sp = w.stack[l].data();
        w.ctx[l].uc_stack.ss_size = w.stack[l].size();
        w.ctx[l].uc_link = &w.main;
        makecontext(&w.ctx[l], (void (*)())trampoline, 1, l);
    }
    for (;;) {
        bool any = false;
        for (int l = 0; l < LANES; ++l) {
            if (w.done[l]) continue;
            any = true;
            w.cur = l;
            swapcontext(&w.main, &w.ctx[l]);
        }
        if (!any) break;
    }
    current() = saved;
}

struct ThreadIdx { int x; };
inline ThreadIdx thread_idx() { return ThreadIdx{lane_id()}; }

}  // namespace warpsim

// ---- the CUDA spellings the device code uses ------------------------------------------------------------------------
#define __device__
#define __forceinline__ inline
#define __noinline__
#define threadIdx (warpsim::thread_idx())

template <class T> inline T __shfl_sync(unsigned, T v, int src) { return warpsim::exchange(v, src, 1); }
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned d) {
    const int l = warpsim::lane_id();
    return warpsim::exchange(v, l >= (int)d ? l - (int)d : l, 2);
}
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned d) {
    const int l = warpsim::lane_id();
    return warpsim::exchange(v, l + (int)d < warpsim::LANES ? l + (int)d : l, 3);
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m) { return warpsim::exchange(v, warpsim::lane_id() ^ m, 4); }
inline unsigned __ballot_sync(unsigned, bool pred) { return warpsim::ballot(pred); }
inline bool __any_sync(unsigned, bool pred) { return warpsim::ballot(pred) != 0u; }
inline bool __all_sync(unsigned, bool pred) { return warpsim::ballot(pred) == 0xffffffffu; }
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
