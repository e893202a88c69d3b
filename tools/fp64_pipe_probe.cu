// Which pipes can the 52-bit FP64 multiplier (fq_fp64.cuh) use side by side on B200?  Independent chains of
//   0: DFMA (rz)                        1: DFMA, DADD, DFMA (the exact split of one limb product)
//   2: the split + 3-input 64-bit add   3: IMAD.WIDE.U32 alone
//   4: DFMA and IMAD.WIDE interleaved   5: 64-bit 3-input integer add alone (IADD3 + IADD3.X)
//   6: DFMA and 64-bit adds interleaved 7: DADD alone
// nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/fp64_pipe_probe tools/fp64_pipe_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int CH = 12;

__global__ void __launch_bounds__(256) p0(uint32_t *out, int iters, double b) {
    double acc[CH];
    for (int j = 0; j < CH; ++j) acc[j] = 1.0 + threadIdx.x * 1e-3 + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CH; ++j) acc[j] = __fma_rz(acc[j], b, 1e-9);
    }
    double s = 0; for (int j = 0; j < CH; ++j) s += acc[j];
    if (s == 0.12345) out[0] = 1;
}
__global__ void __launch_bounds__(256) p1(uint32_t *out, int iters, double b) {
    double a[CH];
    for (int j = 0; j < CH; ++j) a[j] = 4503599627370.0 + threadIdx.x * 1000 + j;
    const double C1 = __longlong_as_double(0x4670000000000000ll), C2 = __longlong_as_double(0x4670000000000001ll);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const double h = __fma_rz(a[j], b, C1);
            const double t = __dsub_rn(C2, h);
            a[j] = __fma_rz(a[j], b, t);   // 2^52 + low part: still an integer below 2^53
        }
    }
    double s = 0; for (int j = 0; j < CH; ++j) s += a[j];
    if (s == 0.12345) out[0] = 1;
}
__global__ void __launch_bounds__(256) p2(uint32_t *out, int iters, double b) {
    double a[CH]; uint64_t c[CH];
    for (int j = 0; j < CH; ++j) { a[j] = 4503599627370.0 + threadIdx.x * 1000 + j; c[j] = j; }
    const double C1 = __longlong_as_double(0x4670000000000000ll), C2 = __longlong_as_double(0x4670000000000001ll);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const double h = __fma_rz(a[j], b, C1);
            const double t = __dsub_rn(C2, h);
            const double l = __fma_rz(a[j], b, t);
            c[j] += (uint64_t)__double_as_longlong(h) + (uint64_t)__double_as_longlong(l);
            a[j] = l;
        }
    }
    uint64_t s = 0; for (int j = 0; j < CH; ++j) s ^= c[j];
    if (s == 0x123456789abcdefull) out[0] = 1;
}
__global__ void __launch_bounds__(256) p3(uint32_t *out, int iters, uint32_t b) {
    uint32_t acc[CH];
    for (int j = 0; j < CH; ++j) acc[j] = (threadIdx.x + 1) * (j + 3);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CH; ++j) { uint64_t p; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(acc[j]), "r"(b)); acc[j] = (uint32_t)p ^ (uint32_t)(p >> 32) ^ j; }
    }
    uint32_t s = 0; for (int j = 0; j < CH; ++j) s ^= acc[j];
    if (s == 0x12345678u) out[0] = 1;
}
__global__ void __launch_bounds__(256) p4(uint32_t *out, int iters, uint32_t b, double bd) {
    uint32_t acc[CH]; double d[CH];
    for (int j = 0; j < CH; ++j) { acc[j] = (threadIdx.x + 1) * (j + 3); d[j] = 1.0 + threadIdx.x * 1e-3 + j; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            uint64_t p; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(acc[j]), "r"(b)); acc[j] = (uint32_t)p ^ (uint32_t)(p >> 32) ^ j;
            d[j] = __fma_rz(d[j], bd, 1e-9);
        }
    }
    uint32_t s = 0; double sd = 0; for (int j = 0; j < CH; ++j) { s ^= acc[j]; sd += d[j]; }
    if (s == 0x12345678u || sd == 0.12345) out[0] = 1;
}
__global__ void __launch_bounds__(256) p5(uint32_t *out, int iters, uint64_t b) {
    uint64_t c[CH];
    for (int j = 0; j < CH; ++j) c[j] = (uint64_t)(threadIdx.x + 1) * (j + 3);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CH; ++j) c[j] = c[j] + c[(j + 1) % CH] + b;
    }
    uint64_t s = 0; for (int j = 0; j < CH; ++j) s ^= c[j];
    if (s == 0x123456789abcdefull) out[0] = 1;
}
__global__ void __launch_bounds__(256) p6(uint32_t *out, int iters, uint64_t b, double bd) {
    uint64_t c[CH]; double d[CH];
    for (int j = 0; j < CH; ++j) { c[j] = (uint64_t)(threadIdx.x + 1) * (j + 3); d[j] = 1.0 + threadIdx.x * 1e-3 + j; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CH; ++j) { c[j] = c[j] + c[(j + 1) % CH] + b; d[j] = __fma_rz(d[j], bd, 1e-9); }
    }
    uint64_t s = 0; double sd = 0; for (int j = 0; j < CH; ++j) { s ^= c[j]; sd += d[j]; }
    if (s == 0x123456789abcdefull || sd == 0.12345) out[0] = 1;
}
__global__ void __launch_bounds__(256) p7(uint32_t *out, int iters, double b) {
    double acc[CH];
    for (int j = 0; j < CH; ++j) acc[j] = 1.0 + threadIdx.x * 1e-3 + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CH; ++j) acc[j] = __dadd_rn(acc[j], b);
    }
    double s = 0; for (int j = 0; j < CH; ++j) s += acc[j];
    if (s == 0.12345) out[0] = 1;
}

// split + a chosen way of folding h and l into the integer accumulators
//   SINK 0: two 2-input 64-bit adds   1: 3-input add of the low words only   2: 3-input LOP3 (xor) on both words
//   3: 3-input 64-bit add with the PREVIOUS step's h (as the multiplier does)   4: one 2-input 64-bit add (l only)
template <int SINK>
__global__ void __launch_bounds__(256) p2x(uint32_t *out, int iters, double b) {
    double a[CH]; uint64_t c[CH];
    for (int j = 0; j < CH; ++j) { a[j] = 4503599627370.0 + threadIdx.x * 1000 + j; c[j] = j; }
    const double C1 = __longlong_as_double(0x4670000000000000ll), C2 = __longlong_as_double(0x4670000000000001ll);
    uint64_t hprev = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const double h = __fma_rz(a[j], b, C1);
            const double t = __dsub_rn(C2, h);
            const double l = __fma_rz(a[j], b, t);
            const uint64_t hb = (uint64_t)__double_as_longlong(h), lb = (uint64_t)__double_as_longlong(l);
            if (SINK == 0) { c[j] += hb; c[(j + 1) % CH] += lb; }
            else if (SINK == 1) { uint32_t lo = (uint32_t)c[j] + (uint32_t)hb + (uint32_t)lb; c[j] = (c[j] & 0xffffffff00000000ull) | lo; }
            else if (SINK == 2) { c[j] ^= hb ^ lb; }
            else if (SINK == 3) { c[j] += lb + hprev; hprev = hb; }
            else { c[j] += lb; }
            a[j] = l;
        }
    }
    uint64_t s = hprev; for (int j = 0; j < CH; ++j) s ^= c[j];
    if (s == 0x123456789abcdefull) out[0] = 1;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    uint32_t *d; cudaMalloc(&d, 256);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int sms = prop.multiProcessorCount, iters = 4096, threads = 256, blocks = sms * 8;
    const char *names[] = {"DFMA.RZ", "DFMA,DADD,DFMA (split)", "split + 64-bit add3", "IMAD.WIDE", "IMAD.WIDE + DFMA", "64-bit add3", "64-bit add3 + DFMA", "DADD"};
    const double per_iter[] = {1, 3, 3, 1, 1, 1, 1, 1};   // counted FP64 (or IMAD / add) instructions per chain step
    printf("SMs %d  clock %.0f MHz   (warp-instruction issue interval per scheduler = 32 / (rate / 4))\n", sms, clk_khz / 1e3);
    for (int k = 0; k < 8; ++k) {
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            switch (k) {
                case 0: p0<<<blocks, threads>>>(d, iters, 1.0000001); break;
                case 1: p1<<<blocks, threads>>>(d, iters, 4503599627370.0); break;
                case 2: p2<<<blocks, threads>>>(d, iters, 4503599627370.0); break;
                case 3: p3<<<blocks, threads>>>(d, iters, 0x9e3779b9u); break;
                case 4: p4<<<blocks, threads>>>(d, iters, 0x9e3779b9u, 1.0000001); break;
                case 5: p5<<<blocks, threads>>>(d, iters, 0x9e3779b97f4a7c15ull); break;
                case 6: p6<<<blocks, threads>>>(d, iters, 0x9e3779b97f4a7c15ull, 1.0000001); break;
                case 7: p7<<<blocks, threads>>>(d, iters, 1.0000001); break;
            }
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double steps = double(blocks) * threads * iters * CH;
        printf("%-26s %8.3f ms  %7.2f chain steps/clk/SM  (%.0f counted instr per step: %6.2f instr/clk/SM)\n", names[k], best,
               steps / (best * 1e-3) / sms / (clk_khz * 1e3), per_iter[k], per_iter[k] * steps / (best * 1e-3) / sms / (clk_khz * 1e3));
    }
    const char *xn[] = {"split + two 64-bit add2", "split + low-word add3", "split + 64-bit xor3", "split + add3 (prev h)", "split + one 64-bit add2"};
    for (int k = 0; k < 5; ++k) {
        float best = 1e30f;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            switch (k) {
                case 0: p2x<0><<<blocks, threads>>>(d, iters, 4503599627370.0); break;
                case 1: p2x<1><<<blocks, threads>>>(d, iters, 4503599627370.0); break;
                case 2: p2x<2><<<blocks, threads>>>(d, iters, 4503599627370.0); break;
                case 3: p2x<3><<<blocks, threads>>>(d, iters, 4503599627370.0); break;
                case 4: p2x<4><<<blocks, threads>>>(d, iters, 4503599627370.0); break;
            }
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        const double steps = double(blocks) * threads * iters * CH;
        const double rate = steps / (best * 1e-3) / sms / (clk_khz * 1e3);
        printf("%-26s %8.3f ms  %7.2f splits/clk/SM = %5.2f scheduler cycles per split\n", xn[k], best, rate, 128.0 / rate);
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
