// Jacobian point arithmetic (X/Z^2, Y/Z^3; infinity <=> Z == 0) on slab slots, executed by a Team.
//
// Replaces multiexp/curves.cu:148-335 (ec_jac::mixed_add / add / dbl) of the reference.  Formulas:
//   madd : classic 8M+3S Jacobian+affine addition (Z2 = 1)
//   add  : classic 12M+4S Jacobian+Jacobian addition
//   dbl  : 2M+6S(+mul_by_a) Jacobian doubling for general a
// all with complete handling of infinity, P+P (falls through to dbl) and P+(-P) (-> infinity),
// selected per lane through store predicates so a warp never diverges on the common path.
#pragma once
#include "fe.cuh"

namespace mnt753 {

#ifdef MNT753_HOST_EMU
inline bool team_any(bool p) { return p; }
#else
__device__ __forceinline__ bool team_any(bool p) { return __any_sync(0xffffffffu, p); }
#endif

// Slot triples / scratch used by the point routines.
struct PtSlots {
    int X1, Y1, Z1;  // accumulator (result in place)
    int X2, Y2, Z2;  // second operand (clobbered by add; Z2 unused by madd)
    int T0, T1, T2;  // scratch
};

template <class F>
struct Ec {
    typedef Team<F> TeamT;

    // acc = 2 * acc on lanes with `pred`; scratch T0..T2.
    static MSM_DEVICE void dbl(const TeamT &T, const PtSlots &s, bool pred) {
        T.sqr(s.T1, s.Z1);            // ZZ
        T.sqr(s.T1, s.T1);            // ZZ^2
        T.mul_by_a(s.T1, s.T1);       // a*ZZ^2
        T.sqr(s.T0, s.Y1);            // YY
        T.mul(s.Z1, s.Y1, s.Z1, pred);
        T.dbl(s.Z1, s.Z1, pred);      // Z3 = 2*Y*Z
        T.sqr(s.T2, s.X1);            // XX
        T.add(s.T1, s.T1, s.T2);
        T.dbl(s.T2, s.T2);
        T.add(s.T1, s.T1, s.T2);      // M = 3*XX + a*ZZ^2
        T.mul(s.T2, s.X1, s.T0);
        T.dbl(s.T2, s.T2);
        T.dbl(s.T2, s.T2);            // S = 4*X*YY
        T.sqr(s.T0, s.T0);
        T.dbl(s.T0, s.T0);
        T.dbl(s.T0, s.T0);
        T.dbl(s.T0, s.T0);            // 8*YYYY
        T.sqr(s.X1, s.T1, pred);
        T.sub(s.X1, s.X1, s.T2, pred);
        T.sub(s.X1, s.X1, s.T2, pred);  // X3 = M^2 - 2S
        T.sub(s.T2, s.T2, s.X1);
        T.mul(s.T2, s.T1, s.T2);
        T.sub(s.Y1, s.T2, s.T0, pred);  // Y3 = M*(S - X3) - 8*YYYY
    }

    // First half of acc += (X2, neg ? -Y2 : Y2): everything that reads the affine operand.
    // After it returns, slots X2/Y2 are dead (the caller may start fetching the next point).
    // `active`: lane has a point this step.  `acc_inf`: lane's accumulator is infinity.
    static MSM_DEVICE void madd_head(const TeamT &T, const PtSlots &s, bool neg, bool active, bool acc_inf) {
        const bool init = active && acc_inf;
        T.copy(s.X1, s.X2, init);
        T.neg_if(s.Y1, s.Y2, neg, init);
        T.set_one(s.Z1, init);
        T.sqr(s.T0, s.Z1);            // Z1Z1
        T.mul(s.T1, s.X2, s.T0);      // U2
        T.mul(s.T2, s.Z1, s.T0);      // Z1^3
        T.mul(s.T2, s.Y2, s.T2);      // S2 (sign applied in the tail)
    }
    // Second half; updates acc_inf.
    static MSM_DEVICE void madd_tail(const TeamT &T, const PtSlots &s, bool neg, bool active, bool &acc_inf) {
        const bool init = active && acc_inf;
        const bool norm = active && !acc_inf;
        T.neg_if(s.T2, s.T2, neg);
        T.sub(s.T1, s.T1, s.X1);      // H
        T.sub(s.T2, s.T2, s.Y1);      // r
        const bool hz = T.is_zero(s.T1);
        bool rz = false;
        if (team_any(norm && hz)) rz = T.is_zero(s.T2);
        const bool wr = norm && !hz;
        T.mul(s.Z1, s.Z1, s.T1, wr);  // Z3 = Z1*H
        T.sqr(s.T0, s.T1);            // HH
        T.mul(s.T1, s.T1, s.T0);      // HHH
        T.mul(s.T0, s.X1, s.T0);      // V = X1*HH
        T.sqr(s.X1, s.T2, wr);        // r^2
        T.sub(s.X1, s.X1, s.T1, wr);
        T.sub(s.X1, s.X1, s.T0, wr);
        T.sub(s.X1, s.X1, s.T0, wr);  // X3 = r^2 - HHH - 2V
        T.sub(s.T0, s.T0, s.X1);
        T.mul(s.T0, s.T2, s.T0);      // r*(V - X3)
        T.mul(s.T1, s.Y1, s.T1);      // Y1*HHH
        T.sub(s.Y1, s.T0, s.T1, wr);  // Y3
        const bool dbl_case = norm && hz && rz;
        if (team_any(dbl_case)) dbl(T, s, dbl_case);
        if (init) acc_inf = false;
        else if (norm && hz && !rz) acc_inf = true;
    }
    static MSM_DEVICE void madd(const TeamT &T, const PtSlots &s, bool neg, bool active, bool &acc_inf) {
        madd_head(T, s, neg, active, acc_inf);
        madd_tail(T, s, neg, active, acc_inf);
    }

    // acc += Q (both Jacobian, infinity <=> Z == 0) on lanes with `active`.  Clobbers Q and T0..T2.
    static MSM_DEVICE void add(const TeamT &T, const PtSlots &s, bool active) {
        const bool inf1 = T.is_zero(s.Z1);
        const bool inf2 = T.is_zero(s.Z2);
        const bool take2 = active && inf1;
        const bool norm = active && !inf1 && !inf2;
        T.copy(s.X1, s.X2, take2);
        T.copy(s.Y1, s.Y2, take2);
        T.copy(s.Z1, s.Z2, take2);
        T.sqr(s.T0, s.Z2);            // Z2Z2
        T.mul(s.T1, s.X1, s.T0);      // U1
        T.mul(s.T0, s.T0, s.Z2);      // Z2^3
        T.mul(s.T0, s.Y1, s.T0);      // S1
        T.sqr(s.T2, s.Z1);            // Z1Z1
        T.mul(s.X2, s.X2, s.T2);      // U2
        T.mul(s.T2, s.T2, s.Z1);      // Z1^3
        T.mul(s.Y2, s.Y2, s.T2);      // S2
        T.sub(s.X2, s.X2, s.T1);      // H
        T.sub(s.Y2, s.Y2, s.T0);      // r
        const bool hz = T.is_zero(s.X2);
        bool rz = false;
        if (team_any(norm && hz)) rz = T.is_zero(s.Y2);
        const bool wr = norm && !hz;
        T.mul(s.Z2, s.Z1, s.Z2);      // Z1*Z2
        T.mul(s.Z1, s.Z2, s.X2, wr);  // Z3 = Z1*Z2*H
        T.sqr(s.T2, s.X2);            // HH
        T.mul(s.X2, s.X2, s.T2);      // HHH
        T.mul(s.T2, s.T1, s.T2);      // V = U1*HH
        T.sqr(s.X1, s.Y2, wr);        // r^2
        T.sub(s.X1, s.X1, s.X2, wr);
        T.sub(s.X1, s.X1, s.T2, wr);
        T.sub(s.X1, s.X1, s.T2, wr);  // X3
        T.sub(s.T2, s.T2, s.X1);
        T.mul(s.T2, s.Y2, s.T2);      // r*(V - X3)
        T.mul(s.T0, s.T0, s.X2);      // S1*HHH
        T.sub(s.Y1, s.T2, s.T0, wr);  // Y3
        const bool dbl_case = norm && hz && rz;
        if (team_any(dbl_case)) dbl(T, s, dbl_case);
        T.set_zero(s.Z1, norm && hz && !rz);
    }
};

}  // namespace mnt753
