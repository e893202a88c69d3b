// 32-bit multiply-add / carry-chain primitives for sm_100a.
//
// Each wrapper is exactly one PTX instruction.  ptxas fuses an adjacent
// `mad{c}.lo.cc.u32 / madc.hi.cc.u32` pair on the same operands into a single
// IMAD.WIDE.U32.X (32x32+64 -> 64 with predicate carry-in/out), which is the instruction the
// whole engine is built on; `asm volatile` keeps NVVM from reordering the carry chains.
//
// MNT753_HOST_EMU: the same names are implemented in plain C++ with an explicit carry flag so that
// the limb-level algorithms in fq.cuh / fe.cuh / ec.cuh can be exercised by the CPU test-suite
// (tests/host_emu) where no GPU exists.  The emulation is never part of the shipped library.
#pragma once
#include <cstdint>

#ifdef MNT753_HOST_EMU
#define MSM_HD
#define MSM_DEVICE inline
struct uint4 { uint32_t x, y, z, w; };
namespace prim {
static thread_local uint32_t cf = 0;
inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; cf = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + cf; cf = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + cf; }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; cf = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - cf; cf = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - cf; }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_lo(a, b), c); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_lo(a, b), c); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return mul_hi(a, b) + c + cf; }
inline uint64_t mad_wide(uint32_t a, uint32_t b, uint64_t c) { return (uint64_t)a * b + c; }
inline uint32_t opaque(uint32_t x) { return x; }
}  // namespace prim
#else
#define MSM_HD __host__ __device__
#define MSM_DEVICE __device__ __forceinline__
namespace prim {
#define MSM_ASM asm volatile
MSM_DEVICE uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; MSM_ASM("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MSM_DEVICE uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; MSM_ASM("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MSM_DEVICE uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; MSM_ASM("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MSM_DEVICE uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; MSM_ASM("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MSM_DEVICE uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; MSM_ASM("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MSM_DEVICE uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; MSM_ASM("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
MSM_DEVICE uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
MSM_DEVICE uint32_t mul_hi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
MSM_DEVICE uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; MSM_ASM("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
MSM_DEVICE uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; MSM_ASM("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
MSM_DEVICE uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; MSM_ASM("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
MSM_DEVICE uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; MSM_ASM("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
MSM_DEVICE uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; MSM_ASM("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
// carry-free 32x32+64 -> 64: one plain IMAD.WIDE.U32 (full-rate, unlike the .X carry form).  Not volatile:
// it is pure arithmetic and ptxas is free to schedule it.
MSM_DEVICE uint64_t mad_wide(uint32_t a, uint32_t b, uint64_t c) { uint64_t r; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r; }
// hides a compile-time constant from NVVM (which would turn u32 x const into a 64-bit multiply whose
// high-word fix-up ptxas does not fold away); ptxas still propagates it into the immediate field.
MSM_DEVICE uint32_t opaque(uint32_t x) { asm("" : "+r"(x)); return x; }
}  // namespace prim
#endif
