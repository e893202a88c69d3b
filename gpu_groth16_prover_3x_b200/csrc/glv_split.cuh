// The split of a G2 scalar  k = k0 + k1 (q mod r)  (glv.cuh): constants and the multi-limb arithmetic, one scalar per
// thread.  In a header of its own so that the CPU test-suite can run exactly this code (tests/host_emu/plan_emu.cpp).
#pragma once
#include "fq.cuh"

namespace mnt753 {

template <int CURVE> struct Glv;
#define MNT753_GLV_TABLE(NAME, N, MACRO) \
    MSM_HD static constexpr uint32_t NAME(int j) { constexpr uint32_t t[N] = MACRO; return t[j]; }
template <> struct Glv<0> {
    MNT753_GLV_TABLE(G1, MNT753_GLV_G_LIMBS, MNT753_GLV_G1_C0_U32)
    MNT753_GLV_TABLE(G2, MNT753_GLV_G_LIMBS, MNT753_GLV_G2_C0_U32)
    MNT753_GLV_TABLE(NA1, MNT753_GLV_LIMBS, MNT753_GLV_NA1_C0_U32)
    MNT753_GLV_TABLE(NA2, MNT753_GLV_LIMBS, MNT753_GLV_NA2_C0_U32)
    MNT753_GLV_TABLE(NB1, MNT753_GLV_LIMBS, MNT753_GLV_NB1_C0_U32)
    MNT753_GLV_TABLE(NB2, MNT753_GLV_LIMBS, MNT753_GLV_NB2_C0_U32)
    MNT753_GLV_TABLE(TWX, NLIMB, MNT753_TWIST_Q_X_C0_U32)
    MNT753_GLV_TABLE(TWY, NLIMB, MNT753_TWIST_Q_Y_C0_U32)
};
template <> struct Glv<1> {
    MNT753_GLV_TABLE(G1, MNT753_GLV_G_LIMBS, MNT753_GLV_G1_C1_U32)
    MNT753_GLV_TABLE(G2, MNT753_GLV_G_LIMBS, MNT753_GLV_G2_C1_U32)
    MNT753_GLV_TABLE(NA1, MNT753_GLV_LIMBS, MNT753_GLV_NA1_C1_U32)
    MNT753_GLV_TABLE(NA2, MNT753_GLV_LIMBS, MNT753_GLV_NA2_C1_U32)
    MNT753_GLV_TABLE(NB1, MNT753_GLV_LIMBS, MNT753_GLV_NB1_C1_U32)
    MNT753_GLV_TABLE(NB2, MNT753_GLV_LIMBS, MNT753_GLV_NB2_C1_U32)
    MNT753_GLV_TABLE(TWX, NLIMB, MNT753_TWIST_Q_X_C1_U32)
    MNT753_GLV_TABLE(TWY, NLIMB, MNT753_TWIST_Q_Y_C1_U32)
};

constexpr int GLV_L = MNT753_GLV_LIMBS, GLV_GL = MNT753_GLV_G_LIMBS;
constexpr int GLV_HALF_LIMBS = 12;       // limbs of a half scalar as stored (|k| < 2^379, sign in bit 383)

// k: the scalar as a plain integer in [0, r).  r[h]: |k_h| in GLV_L limbs with the sign in bit 31 of limb GLV_HALF_LIMBS - 1.
template <int CURVE>
MSM_DEVICE void glv_split(const uint32_t (&k)[NLIMB], uint32_t (&r)[2][GLV_L]) {
    typedef Glv<CURVE> C;
    // c_i = (k * G_i) >> 768
    uint32_t c[2][GLV_GL];
#pragma unroll
    for (int which = 0; which < 2; ++which) {
        uint32_t acc[NLIMB + GLV_GL];
#pragma unroll
        for (int i = 0; i < NLIMB + GLV_GL; ++i) acc[i] = 0u;
#pragma unroll
        for (int j = 0; j < GLV_GL; ++j) {
            const uint32_t g = which == 0 ? C::G1(j) : C::G2(j);
            uint32_t carry = 0u;
#pragma unroll
            for (int i = 0; i < NLIMB; ++i) {
                const unsigned long long t = (unsigned long long)k[i] * g + acc[i + j] + carry;
                acc[i + j] = (uint32_t)t;
                carry = (uint32_t)(t >> 32);
            }
            acc[NLIMB + j] = carry;
        }
#pragma unroll
        for (int j = 0; j < GLV_GL; ++j) c[which][j] = acc[NLIMB + j];
    }
    // k0 = k + c1 NA1 + c2 NA2,  k1 = c1 NB1 + c2 NB2   modulo 2^(32 GLV_L)
#pragma unroll
    for (int i = 0; i < GLV_L; ++i) { r[0][i] = k[i]; r[1][i] = 0u; }
#pragma unroll
    for (int half = 0; half < 2; ++half)
#pragma unroll
        for (int which = 0; which < 2; ++which)
#pragma unroll
            for (int j = 0; j < GLV_GL; ++j) {
                uint32_t carry = 0u;
#pragma unroll
                for (int i = 0; i + j < GLV_L; ++i) {
                    const uint32_t m = half == 0 ? (which == 0 ? C::NA1(i) : C::NA2(i)) : (which == 0 ? C::NB1(i) : C::NB2(i));
                    const unsigned long long t = (unsigned long long)c[which][j] * m + r[half][i + j] + carry;
                    r[half][i + j] = (uint32_t)t;
                    carry = (uint32_t)(t >> 32);
                }
            }
    // magnitude and sign
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const uint32_t neg = r[half][GLV_L - 1] >> 31;
        uint32_t carry = neg;
#pragma unroll
        for (int i = 0; i < GLV_L; ++i) {
            const unsigned long long t = (unsigned long long)(neg ? ~r[half][i] : r[half][i]) + carry;
            r[half][i] = (uint32_t)t;
            carry = (uint32_t)(t >> 32);
        }
        r[half][GLV_HALF_LIMBS - 1] |= neg << 31;
    }
}

}  // namespace mnt753
