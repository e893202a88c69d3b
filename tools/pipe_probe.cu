// Pipe-rate probe for the roofline denominator (sm_100a).  Each kernel runs `iters` trips of an
// unrolled body of CH independent dependency chains per thread; the host prints body-ops per clock
// per SM.  The SASS form of each body is checked with cuobjdump (see profiles/*pipe_probe*).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../gpu_groth16_prover_3x_b200/csrc/fq.cuh"
constexpr int CH = 16;

// 0: mad.wide.u32 with 64-bit addend (ptxas decides the SASS form)
__global__ void __launch_bounds__(256) p_wide_acc(uint32_t *out, int iters, uint32_t b) {
    uint64_t acc[CH];
    for (int j = 0; j < CH; ++j) acc[j] = (uint64_t)(threadIdx.x + 1) * (j + 3) + 0x100000001ull * j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CH; ++j) { uint32_t m = (uint32_t)acc[j]; asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[j]) : "r"(m), "r"(b)); }
    }
    uint64_t s = 0; for (int j = 0; j < CH; ++j) s ^= acc[j];
    if (s == 0x123456789abcdefull) out[0] = 1;
}
// 1: mul.wide.u32 (no addend), folded into the chain by one LOP3 (xor of hi and lo)
__global__ void __launch_bounds__(256) p_wide_noacc(uint32_t *out, int iters, uint32_t b) {
    uint32_t acc[CH];
    for (int j = 0; j < CH; ++j) acc[j] = (threadIdx.x + 1) * (j + 3);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CH; ++j) { uint64_t p; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(acc[j]), "r"(b)); acc[j] = (uint32_t)p ^ (uint32_t)(p >> 32) ^ j; }
    }
    uint32_t s = 0; for (int j = 0; j < CH; ++j) s ^= acc[j];
    if (s == 0x12345678u) out[0] = 1;
}
// 2: 32-bit mad.lo
__global__ void __launch_bounds__(256) p_lo(uint32_t *out, int iters, uint32_t b) {
    uint32_t acc[CH];
    for (int j = 0; j < CH; ++j) acc[j] = (threadIdx.x + 1) * (j + 3);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CH; ++j) asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(acc[j]) : "r"(b));
    }
    uint32_t s = 0; for (int j = 0; j < CH; ++j) s ^= acc[j];
    if (s == 0x12345678u) out[0] = 1;
}
// 3: mad.hi.u32
__global__ void __launch_bounds__(256) p_hi(uint32_t *out, int iters, uint32_t b) {
    uint32_t acc[CH];
    for (int j = 0; j < CH; ++j) acc[j] = (threadIdx.x + 1) * (j + 3) + 0x9e3779b9u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CH; ++j) asm volatile("mad.hi.u32 %0, %0, %1, %0;" : "+r"(acc[j]) : "r"(b));
    }
    uint32_t s = 0; for (int j = 0; j < CH; ++j) s ^= acc[j];
    if (s == 0x12345678u) out[0] = 1;
}
// 4: the engine's Montgomery product (fq.cuh: 1152 + 24 IMAD.WIDE.U32(.X) per product)
__global__ void __launch_bounds__(128) p_cios_row(uint32_t *out, int iters, uint32_t b0) {
    using namespace mnt753;
    fq_t x, y;
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) { x[i] = ModA::R1(i) ^ (threadIdx.x * 7u + i); y[i] = ModA::R2(i) ^ (blockIdx.x + i + b0); }
    x[NLIMB - 1] &= 0xffffu; y[NLIMB - 1] &= 0xffffu;
    for (int it = 0; it < iters; ++it) fq_mul<ModA>(x, x, y);
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) s ^= x[i];
    if (s == 0x12345678u) out[0] = 1;
}
// 5: IADD3 (three live inputs)
__global__ void __launch_bounds__(256) p_iadd3(uint32_t *out, int iters, uint32_t b) {
    uint32_t acc[CH];
    for (int j = 0; j < CH; ++j) acc[j] = (threadIdx.x + 1) * (j + 3);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CH; ++j) acc[j] = acc[j] + acc[(j + 1) % CH] + b;
    }
    uint32_t s = 0; for (int j = 0; j < CH; ++j) s ^= acc[j];
    if (s == 0x12345678u) out[0] = 1;
}
// 6: DFMA
__global__ void __launch_bounds__(256) p_dfma(uint32_t *out, int iters, double b) {
    double acc[CH];
    for (int j = 0; j < CH; ++j) acc[j] = 1.0 + threadIdx.x * 1e-3 + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CH; ++j) acc[j] = fma(acc[j], b, 1e-9);
    }
    double s = 0; for (int j = 0; j < CH; ++j) s += acc[j];
    if (s == 0.12345) out[0] = 1;
}
// 7: mul.wide (no addend) + one 64-bit add into a 64-bit accumulator, written in C (ptxas decides)
__global__ void __launch_bounds__(256) p_wide_add64(uint32_t *out, int iters, uint32_t b) {
    uint64_t acc[CH]; uint32_t m[CH];
    for (int j = 0; j < CH; ++j) { acc[j] = (uint64_t)(threadIdx.x + 1) * (j + 3); m[j] = threadIdx.x * 77 + j; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CH; ++j) { acc[j] += (uint64_t)m[j] * b; m[j] = (uint32_t)acc[(j + 5) % CH]; }
    }
    uint64_t s = 0; for (int j = 0; j < CH; ++j) s ^= acc[j];
    if (s == 0x123456789abcdefull) out[0] = 1;
}
// 8: co-issue: one mul.wide (no addend) + two IADD3 per op (hi and lo folded into separate 32-bit sums)
__global__ void __launch_bounds__(256) p_wide_2add(uint32_t *out, int iters, uint32_t b) {
    uint32_t lo[CH], hi[CH], m[CH];
    for (int j = 0; j < CH; ++j) { lo[j] = j; hi[j] = 3 * j; m[j] = threadIdx.x * 77 + j; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CH; ++j) { uint64_t p; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(m[j]), "r"(b)); lo[j] += (uint32_t)p; hi[j] += (uint32_t)(p >> 32); m[j] = lo[(j + 5) % CH]; }
    }
    uint32_t s = 0; for (int j = 0; j < CH; ++j) s ^= lo[j] ^ hi[j];
    if (s == 0x12345678u) out[0] = 1;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    uint32_t *d; cudaMalloc(&d, 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int sms = prop.multiProcessorCount; int iters = 4096;
    const char *names[] = {"mad.wide.u32 +acc64", "mul.wide.u32 + LOP3", "mad.lo.u32", "mad.hi.u32", "Fq modmul x1176 MAC", "IADD3", "DFMA", "mul.wide + add64 (C)", "mul.wide + 2 IADD"};
    printf("SMs %d  clock %.0f MHz\n", sms, clk_khz / 1e3);
    for (int k = 0; k < 9; ++k) {
        float best = 1e30f;
        int threads = k == 4 ? 128 : 256, blocks = sms * (k == 4 ? 8 : 8);
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            switch (k) {
                case 0: p_wide_acc<<<blocks, threads>>>(d, iters, 0x9e3779b9u); break;
                case 1: p_wide_noacc<<<blocks, threads>>>(d, iters, 0x9e3779b9u); break;
                case 2: p_lo<<<blocks, threads>>>(d, iters, 0x9e3779b9u); break;
                case 3: p_hi<<<blocks, threads>>>(d, iters, 0x9e3779b9u); break;
                case 4: p_cios_row<<<blocks, threads>>>(d, iters, 0x9e3779b9u); break;
                case 5: p_iadd3<<<blocks, threads>>>(d, iters, 0x9e3779b9u); break;
                case 6: p_dfma<<<blocks, threads>>>(d, iters, 1.0000001); break;
                case 7: p_wide_add64<<<blocks, threads>>>(d, iters, 0x9e3779b9u); break;
                case 8: p_wide_2add<<<blocks, threads>>>(d, iters, 0x9e3779b9u); break;
            }
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double ops = double(blocks) * threads * iters * (k == 4 ? 1176.0 : CH);
        printf("%-24s %8.3f ms  %9.1f Gop/s  %6.2f op/clk/SM (at %.0f MHz)\n", names[k], best, ops / best / 1e6,
               ops / (best * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1e3);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
