// Throughput of the two Montgomery multipliers -- the 32-bit CIOS on the IMAD.WIDE pipe (fq.cuh, the engine's) and
// the 52-bit form on the FP64 pipe (fq_fp64.cuh, experiment) -- alone and side by side, for a chosen number of
// warps per SM AND a chosen split of those warps into blocks (the same twelve warps run 12 % faster as one block
// of 384 threads than as three of 128).  Also checks on the device that both multipliers give identical results.
// Results: profiles/r01_mul_sched_probe.txt.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/mul_sched_probe tools/mul_sched_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "experiments/fq_fp64.cuh"
#include "experiments/fq_experiments.cuh"

using namespace mnt753;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

template <class M>
__device__ __forceinline__ void seed(fq_t &x, fq_t &y) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) { x[i] = M::R1(i) ^ (g * 0x9e3779b9u + i * 0x85ebca6bu); y[i] = M::R2(i) ^ (g * 0xc2b2ae35u + i); }
    x[NLIMB - 1] &= 0xffffu;
    y[NLIMB - 1] &= 0xffffu;
}

// MODE 0: CIOS only, 1: FP64 only, 2: every warp alternates (odd warps start with the FP64 form),
// 3: warps 0 mod 3 run CIOS, the others FP64
template <class M, int MODE>
__global__ void __launch_bounds__(384) k_mul(uint32_t *out, int iters) {
    extern __shared__ uint4 dummy[];
    fq_t x, y;
    seed<M>(x, y);
    const int warp = (threadIdx.x >> 5) + blockIdx.x * (blockDim.x >> 5);
    if (MODE == 0) { for (int it = 0; it < iters; ++it) fq_mul<M>(x, x, y); }
    else if (MODE == 1) { for (int it = 0; it < iters; ++it) fq_mul_fp<M>(x, x, y); }
    else if (MODE == 2) {
        int ph = warp & 1;
        for (int it = 0; it < iters; ++it) { if (ph) fq_mul_fp<M>(x, x, y); else fq_mul<M>(x, x, y); ph ^= 1; }
    } else if (MODE == 4 || MODE == 5) {     // rolled CIOS, b streamed from shared memory (own lane's column)
        uint4 *sl = dummy + threadIdx.x;
        for (int q = 0; q < 6; ++q) sl[q * blockDim.x] = make_uint4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
        BQuads<1> src; src.stride = blockDim.x; src.p[0] = sl;
        for (int it = 0; it < iters; ++it) {
            uint32_t aa[1][NLIMB];
#pragma unroll
            for (int i = 0; i < NLIMB; ++i) aa[0][i] = x[i];
            if (MODE == 4) fq_dot_rolled<M, 1, 4>(x, aa, src); else fq_dot_rolled<M, 1, 8>(x, aa, src);
        }
    } else if (MODE == 6) {     // CIOS, unrolled, b streamed from shared memory like Team::mul
        uint4 *sl = dummy + threadIdx.x;
        for (int q = 0; q < 6; ++q) sl[q * blockDim.x] = make_uint4(y[4 * q], y[4 * q + 1], y[4 * q + 2], y[4 * q + 3]);
        BQuads<1> src; src.stride = blockDim.x; src.p[0] = sl;
        for (int it = 0; it < iters; ++it) {
            uint32_t aa[1][NLIMB];
#pragma unroll
            for (int i = 0; i < NLIMB; ++i) aa[0][i] = x[i];
            fq_dot<M, 1>(x, aa, src);
        }
    } else if (MODE == 7) {     // CIOS, warps of a block deliberately out of phase
        const long long t0 = clock64();
        while (clock64() - t0 < (long long)(threadIdx.x >> 5) * 431) { }
        for (int it = 0; it < iters; ++it) fq_mul<M>(x, x, y);
    } else {
        if (warp % 3 == 0) { for (int it = 0; it < iters; ++it) fq_mul<M>(x, x, y); }
        else { for (int it = 0; it < iters; ++it) fq_mul_fp<M>(x, x, y); }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < NLIMB; ++i) s ^= x[i];
    if (s == 0x12345678u) out[0] = 1;
}

template <class M>
__global__ void k_check(uint32_t *bad, int iters) {
    fq_t x, y, r0, r1;
    seed<M>(x, y);
    for (int it = 0; it < iters; ++it) {
        fq_mul<M>(r0, x, y);
        fq_mul_fp<M>(r1, x, y);
        bool ok = true;
#pragma unroll
        for (int i = 0; i < NLIMB; ++i) ok = ok && (r0[i] == r1[i]);
        if (!ok) atomicAdd(bad, 1u);
#pragma unroll
        for (int i = 0; i < NLIMB; ++i) { y[i] = x[i]; x[i] = r0[i]; }
    }
}

template <class M, int MODE>
void run(const char *name, int sms, int bps, int iters, double clock_ghz, int threads = 128) {
    uint32_t *d;
    CK(cudaMalloc(&d, 256));
    const int smem = (227 * 1024) / bps - 1024;   // dynamic shared memory pins the number of resident blocks
    CK(cudaFuncSetAttribute(k_mul<M, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0));
        k_mul<M, MODE><<<sms * bps, threads, smem>>>(d, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
    }
    CK(cudaGetLastError());
    const double muls = double(sms) * bps * threads * iters;
    const double gps = muls / (ms * 1e6);
    // cycles of one scheduler per warp-level product
    const double cyc = (ms * 1e-3 * clock_ghz * 1e9) * (sms * 4.0) / (muls / 32.0);
    printf("%-18s %3d thr x %2d blk = warps/SM %2d  %8.3f G modmul/s  %7.0f scheduler-cycles per warp product (at %.3f GHz)\n", name, threads, bps, bps * threads / 32, gps, cyc, clock_ghz);
    cudaFree(d);
}

int main(int argc, char **argv) {
    int dev = 0;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    const int sms = prop.multiProcessorCount;
    const double ghz = prop.clockRate * 1e-6;
    printf("%s, %d SMs, %.3f GHz max\n", prop.name, sms, ghz);
    uint32_t *bad;
    CK(cudaMalloc(&bad, 4));
    CK(cudaMemset(bad, 0, 4));
    k_check<ModA><<<64, 128>>>(bad, 64);
    k_check<ModB><<<64, 128>>>(bad, 64);
    uint32_t hbad = 1;
    CK(cudaMemcpy(&hbad, bad, 4, cudaMemcpyDeviceToHost));
    printf("device check: %u mismatches in %d products\n", hbad, 2 * 64 * 128 * 64);
    const int iters = argc > 1 ? atoi(argv[1]) : 2000;
    const int bpss[] = {1, 2, 3, 4, 6, 8};
    const int combos[][2] = {{384, 1}, {128, 3}, {128, 2}, {256, 1}};
    for (auto &cb : combos) run<ModA, 0>("CIOS (IMAD.WIDE)", sms, cb[1], iters, ghz, cb[0]);
    const int c2[][2] = {{384, 1}, {128, 3}, {128, 2}, {256, 1}};
    for (auto &cb : c2) run<ModA, 6>("CIOS, b in smem", sms, cb[1], iters, ghz, cb[0]);
    for (auto &cb : c2) run<ModA, 4>("rolled x4, b smem", sms, cb[1], iters, ghz, cb[0]);
    for (auto &cb : c2) run<ModA, 5>("rolled x8, b smem", sms, cb[1], iters, ghz, cb[0]);
    for (auto &cb : c2) run<ModA, 7>("CIOS dephased", sms, cb[1], iters, ghz, cb[0]);
    for (int b : bpss) run<ModA, 1>("FP64 (DFMA)", sms, b, iters, ghz);
    for (int b : bpss) run<ModA, 2>("alternating", sms, b, iters, ghz);
    for (int b : bpss) run<ModA, 3>("1/3 CIOS, 2/3 FP64", sms, b, iters, ghz);
    return hbad ? 2 : 0;
}
