// Simulation of ONE CUDA thread block on the host, for kernels whose threads meet only at __syncthreads(): every thread
// is a user-level context (ucontext); a thread runs until the next barrier and yields, the barrier opens when all
// threads of the block have arrived (or returned).  Between barriers the threads run one after the other, which is a
// legal schedule for race-free kernels; a kernel with a race between barriers is not what this is for.  Test-only.
#pragma once
#include <ucontext.h>

#include <cstdint>
#include <functional>
#include <vector>

namespace blocksim {

struct Dim { unsigned x, y, z; };
struct State {
    ucontext_t main;
    std::vector<ucontext_t> ctx;
    std::vector<std::vector<char>> stack;
    std::vector<char> done;
    unsigned cur = 0;
    Dim block_idx{0, 0, 0}, block_dim{1, 1, 1}, grid_dim{1, 1, 1};
    std::function<void()> body;
};
inline State *&current() {
    static thread_local State *s = nullptr;
    return s;
}
inline void syncthreads() {
    State *s = current();
    swapcontext(&s->ctx[s->cur], &s->main);
}
inline void trampoline() {
    State *s = current();
    s->body();
    s->done[s->cur] = 1;
    swapcontext(&s->ctx[s->cur], &s->main);
}
// kernel<<<grid, threads>>>(...): body() is the kernel call; blockIdx / threadIdx / blockDim read the simulator's state
inline void launch(unsigned grid, unsigned threads, const std::function<void()> &body) {
    State st;
    st.body = body;
    st.grid_dim = Dim{grid, 1, 1};
    st.block_dim = Dim{threads, 1, 1};
    st.ctx.resize(threads);
    st.stack.resize(threads);
    st.done.resize(threads);
    State *saved = current();
    current() = &st;
    for (unsigned t = 0; t < threads; ++t) st.stack[t].resize(128 << 10);
    for (unsigned b = 0; b < grid; ++b) {
        st.block_idx = Dim{b, 0, 0};
        for (unsigned t = 0; t < threads; ++t) {
            st.done[t] = 0;
            getcontext(&st.ctx[t]);
            st.ctx[t].uc_stack.ss_sp = st.stack[t].data();
            st.ctx[t].uc_stack.ss_size = st.stack[t].size();
            st.ctx[t].uc_link = &st.main;
            makecontext(&st.ctx[t], (void (*)())trampoline, 0);
        }
        for (;;) {
            bool any = false;
            for (unsigned t = 0; t < threads; ++t) {
                if (st.done[t]) continue;
                any = true;
                st.cur = t;
                swapcontext(&st.main, &st.ctx[t]);
            }
            if (!any) break;
        }
    }
    current() = saved;
}
inline Dim thread_idx() { return Dim{current()->cur, 0, 0}; }

}  // namespace blocksim

#define __global__
#define __device__
#define __forceinline__ inline
#define __noinline__
#define __shared__
#define __launch_bounds__(...)
#define threadIdx (blocksim::thread_idx())
#define blockIdx (blocksim::current()->block_idx)
#define blockDim (blocksim::current()->block_dim)
#define gridDim (blocksim::current()->grid_dim)
inline void __syncthreads() { blocksim::syncthreads(); }
inline unsigned __brev(unsigned v) {
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0f0f0f0fu) | ((v & 0x0f0f0f0fu) << 4);
    v = ((v >> 8) & 0x00ff00ffu) | ((v & 0x00ff00ffu) << 8);
    return (v >> 16) | (v << 16);
}
