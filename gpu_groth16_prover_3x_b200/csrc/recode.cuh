// Signed-digit recoding of a scalar (k_count / k_scatter, msm_kernels.cuh).  In a header of its own so that the CPU
// test-suite can run exactly this code for every window width (tests/host_emu/plan_emu.cpp).
#pragma once
#include <cstdint>

namespace mnt753 {

// signed-digit recoding of a 753-bit integer into W windows of c bits: digits in
// [-2^(c-1)+1, 2^(c-1)], a borrow of 2^c carried into the next window; W*c >= 754 so the top
// window never overflows.  f(w, digit) is called for non-zero digits only.
template <class Fn>
__device__ __forceinline__ void for_each_digit(const uint32_t *k, int nl, int c, int W, Fn f) {
    uint32_t carry = 0;
    const uint32_t mask = (1u << c) - 1u, half = 1u << (c - 1);
    for (int w = 0; w < W; ++w) {
        const int o = w * c, word = o >> 5, sh = o & 31;
        uint64_t v = word < nl ? k[word] : 0u;
        if (word + 1 < nl) v |= (uint64_t)k[word + 1] << 32;
        uint32_t raw = ((uint32_t)(v >> sh) & mask) + carry;
        int d;
        if (raw > half) { d = (int)raw - (int)(1u << c); carry = 1; } else { d = (int)raw; carry = 0; }
        if (d != 0) f(w, d);
    }
}

}  // namespace mnt753
