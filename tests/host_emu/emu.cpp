// CPU emulation harness for the limb-level device code (fq.cuh / fe.cuh / ec.cuh).
// Compiled by tests/conftest.py with g++ -DMNT753_HOST_EMU: the PTX carry-chain primitives are
// replaced by C++ equivalents (prim.cuh) and a Team runs its DEG coefficients sequentially, so the
// exact same templates that the kernels instantiate are checked against the oracle on machines
// without a GPU.  Test-only; never shipped.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../gpu_groth16_prover_3x_b200/csrc/curves.cuh"
#include "../../gpu_groth16_prover_3x_b200/csrc/batch_affine.cuh"
#include "../../gpu_groth16_prover_3x_b200/csrc/glv_split.cuh"
#include "../../tools/experiments/fq_fp64.cuh"
#include "../../tools/experiments/fq_experiments.cuh"

using namespace mnt753;

namespace {

template <class G>
struct Emu {
    typedef typename G::F F;
    static constexpr int DEG = F::DEG;
    static constexpr int NE = 12;
    std::vector<uint4> slab;
    Team<F> T;
    Emu() : slab((size_t)NE * DEG * QUADS * LANES) {
        memset(slab.data(), 0xA5, slab.size() * sizeof(uint4));
        T.slab = slab.data() + 7;  // arbitrary lane
        T.flags = nullptr;
        T.comp = 0;
        T.bar_id = 0;
    }
    void put(int e, const uint64_t *w) {  // DEG*12 u64 wire element
        for (int c = 0; c < DEG; ++c) {
            fq_t x;
            memcpy(x, w + 12 * c, 96);
            T.st(e, c, x);
        }
    }
    void get(int e, uint64_t *w) {
        for (int c = 0; c < DEG; ++c) {
            fq_t x;
            T.ld(x, e, c);
            memcpy(w + 12 * c, x, 96);
        }
    }
};

const PtSlots SL = {0, 1, 2, 3, 4, 5, 6, 7, 8};

template <class G>
int field_op(int op, size_t n, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    Emu<G> E;
    const size_t st = 12 * Emu<G>::DEG;
    for (size_t i = 0; i < n; ++i) {
        E.put(0, a + i * st);
        if (b) E.put(1, b + i * st);
        switch (op) {
            case 0: E.T.mul(2, 0, 1); break;
            case 1: E.T.add(2, 0, 1); break;
            case 2: E.T.sub(2, 0, 1); break;
            case 3: E.T.sqr(2, 0); break;
            case 5: E.T.neg_if(2, 0, true); break;
            case 6: E.T.mul_by_a(2, 0); break;
            case 7: E.T.mul(0, 0, 1); E.T.copy(2, 0); break;  // in place
            case 8: E.T.dbl(2, 0); break;
            default: return -1;
        }
        E.get(2, out + i * st);
    }
    return 0;
}

// op 0: madd (acc jacobian, p affine x||y, flags: bit0 neg, bit1 acc_inf) -> acc jacobian, returns acc_inf
// op 1: add  (acc jacobian, q jacobian)
// op 2: dbl
template <class G>
int point_op(int op, const uint64_t *acc, const uint64_t *q, int flags, uint64_t *out) {
    Emu<G> E;
    const size_t st = 12 * Emu<G>::DEG;
    E.put(0, acc); E.put(1, acc + st); E.put(2, acc + 2 * st);
    bool acc_inf = flags & 2;
    int ret = 0;
    if (op == 0) {
        E.put(3, q); E.put(4, q + st);
        Ec<typename G::F>::madd(E.T, SL, flags & 1, true, acc_inf);
        ret = acc_inf;
        if (acc_inf) E.T.set_zero(2);
    } else if (op == 1) {
        E.put(3, q); E.put(4, q + st); E.put(5, q + 2 * st);
        Ec<typename G::F>::add(E.T, SL, true);
    } else if (op == 2) {
        Ec<typename G::F>::dbl(E.T, SL, true);
    } else return -1;
    E.get(0, out); E.get(1, out + st); E.get(2, out + 2 * st);
    return ret;
}

// one batched-affine addition with a batch of one: forward, inversion of the lone denominator, backward.
// has2 = 0: copy of p1.  Returns 1 when the result is infinity.
template <class G>
int affine_add(const uint64_t *p1, const uint64_t *p2, int has2, uint64_t *out) {
    typedef typename G::F F;
    Emu<G> E;
    const size_t st = 12 * Emu<G>::DEG;
    const BaSlots s = {0, 1, 2, 3, 4, 5};
    E.put(s.X1, p1); E.put(s.Y1, p1 + st);
    if (has2) { E.put(s.X2, p2); E.put(s.Y2, p2 + st); }
    E.T.set_one(s.INV);
    const uint32_t code = pair_forward(E.T, s, true, has2 != 0, [](bool) {});
    E.T.copy(s.PRE, s.INV);                 // exclusive prefix of a batch of one
    E.T.mul(s.INV, s.INV, s.X2);
    tile_inverse(E.T, s.INV, 8, 9, 10, 11);
    E.put(s.X1, p1); E.put(s.Y1, p1 + st);
    if (has2) { E.put(s.X2, p2); E.put(s.Y2, p2 + st); }
    const bool inf = pair_backward(E.T, s, code);
    E.get(s.X2, out); E.get(s.Y2, out + st);
    return inf ? 1 : 0;
}

// The UNCLASSIFIED passes of a tile (ba_tile, batch_affine.cuh) as one lane runs them: B generic additions that share
// one inversion.  flags[i] bit 0 / 1: operand 1 / 2 is referenced with the negation flag (its ordinate is negated after
// the load, as the kernel does).  Returns 1 when the product of the denominators is zero (the kernel then runs the
// classified passes instead) and leaves out untouched.
template <class G>
int batch_add_generic(size_t B, const uint64_t *p1, const uint64_t *p2, const int *flags, uint64_t *out) {
    Emu<G> E;
    const size_t st = 12 * Emu<G>::DEG;
    const BaSlots s = {0, 1, 2, 3, 4, 5};
    const int PARK = 6;
    std::vector<uint64_t> prefix(B * st);
    E.T.set_one(s.INV);
    for (size_t i = 0; i < B; ++i) {
        E.put(s.X1, p1 + 2 * i * st);
        E.put(s.X2, p2 + 2 * i * st);
        E.T.copy(PARK, s.INV);
        E.get(PARK, prefix.data() + i * st);                  // exclusive prefix, parked in the addition's output slot
        pair_generic_forward(E.T, s, true);
    }
    if (E.T.is_zero(s.INV)) return 1;
    tile_inverse(E.T, s.INV, s.X1, s.Y1, s.X2, s.Y2);
    for (size_t k = B; k-- > 0;) {
        E.put(s.X1, p1 + 2 * k * st); E.put(s.Y1, p1 + 2 * k * st + st);
        E.put(s.X2, p2 + 2 * k * st); E.put(s.Y2, p2 + 2 * k * st + st);
        E.put(s.PRE, prefix.data() + k * st);
        if (flags[k] & 1) E.T.neg_if(s.Y1, s.Y1, true, true);
        if (flags[k] & 2) E.T.neg_if(s.Y2, s.Y2, true, true);
        pair_generic_backward(E.T, s, true);
        E.get(s.X2, out + 2 * k * st);
        E.get(s.Y2, out + 2 * k * st + st);
    }
    return 0;
}

// psi(x, y) = (cX Frob(x), cY Frob(y)) of an affine G2 point, the steps of k_psi_many (glv.cuh)
template <class G>
int psi_point(const uint64_t *p, uint64_t *out) {
    typedef Glv<G::CURVE> K;
    Emu<G> E;
    const size_t st = 12 * Emu<G>::DEG;
    E.put(0, p); E.put(1, p + st);
    E.T.frob(2, 0, 1);
    E.T.frob(3, 1, 1);
    fq_t cx, cy;
    for (int i = 0; i < NLIMB; ++i) { cx[i] = K::TWX(i); cy[i] = K::TWY(i); }
    E.T.scale_fq(2, 2, cx);
    E.T.scale_fq(3, 3, cy);
    E.get(2, out); E.get(3, out + st);
    return 0;
}

template <class G>
int field_inv(size_t n, const uint64_t *a, uint64_t *out) {
    Emu<G> E;
    const size_t st = 12 * Emu<G>::DEG;
    for (size_t i = 0; i < n; ++i) {
        E.put(0, a + i * st);
        E.T.inv_lane0(1, 0, 2, 3);
        E.get(1, out + i * st);
    }
    return 0;
}

}  // namespace

extern "C" {
int emu_affine_add(int curve, int group, const uint64_t *p1, const uint64_t *p2, int has2, uint64_t *out) {
    if (curve == 0 && group == 1) return affine_add<Mnt4G1>(p1, p2, has2, out);
    if (curve == 0 && group == 2) return affine_add<Mnt4G2>(p1, p2, has2, out);
    if (curve == 1 && group == 1) return affine_add<Mnt6G1>(p1, p2, has2, out);
    if (curve == 1 && group == 2) return affine_add<Mnt6G2>(p1, p2, has2, out);
    return -1;
}
int emu_batch_add_generic(int curve, int group, size_t B, const uint64_t *p1, const uint64_t *p2, const int *flags, uint64_t *out) {
    if (curve == 0 && group == 1) return batch_add_generic<Mnt4G1>(B, p1, p2, flags, out);
    if (curve == 0 && group == 2) return batch_add_generic<Mnt4G2>(B, p1, p2, flags, out);
    if (curve == 1 && group == 1) return batch_add_generic<Mnt6G1>(B, p1, p2, flags, out);
    if (curve == 1 && group == 2) return batch_add_generic<Mnt6G2>(B, p1, p2, flags, out);
    return -1;
}
int emu_psi(int curve, const uint64_t *p, uint64_t *out) {
    return curve == 0 ? psi_point<Mnt4G2>(p, out) : psi_point<Mnt6G2>(p, out);
}
int emu_field_inv(int curve, int group, size_t n, const uint64_t *a, uint64_t *out) {
    if (curve == 0 && group == 1) return field_inv<Mnt4G1>(n, a, out);
    if (curve == 0 && group == 2) return field_inv<Mnt4G2>(n, a, out);
    if (curve == 1 && group == 1) return field_inv<Mnt6G1>(n, a, out);
    if (curve == 1 && group == 2) return field_inv<Mnt6G2>(n, a, out);
    return -1;
}
int emu_field_op(int curve, int group, int op, size_t n, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    if (curve == 0 && group == 1) return field_op<Mnt4G1>(op, n, a, b, out);
    if (curve == 0 && group == 2) return field_op<Mnt4G2>(op, n, a, b, out);
    if (curve == 1 && group == 1) return field_op<Mnt6G1>(op, n, a, b, out);
    if (curve == 1 && group == 2) return field_op<Mnt6G2>(op, n, a, b, out);
    return -1;
}
int emu_point_op(int curve, int group, int op, const uint64_t *acc, const uint64_t *q, int flags, uint64_t *out) {
    if (curve == 0 && group == 1) return point_op<Mnt4G1>(op, acc, q, flags, out);
    if (curve == 0 && group == 2) return point_op<Mnt4G2>(op, acc, q, flags, out);
    if (curve == 1 && group == 1) return point_op<Mnt6G1>(op, acc, q, flags, out);
    if (curve == 1 && group == 2) return point_op<Mnt6G2>(op, acc, q, flags, out);
    return -1;
}
// scalar-field helper used by the digit kernel: Montgomery -> integer
// the multiplier implementations on register operands: which = 0 CIOS (engine), 1 reduced-radix (experiment),
// 2 FP64-pipe form (fq_fp64.cuh, experiment)
int emu_fq_mul(int modulus, int which, size_t n, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    for (size_t i = 0; i < n; ++i) {
        fq_t x, y, r;
        memcpy(x, a + 12 * i, 96);
        memcpy(y, b + 12 * i, 96);
        if (modulus == 0) { if (which == 2) fq_mul_fp<ModA>(r, x, y); else if (which) fq_mul_rr<ModA>(r, x, y); else fq_mul<ModA>(r, x, y); }
        else { if (which == 2) fq_mul_fp<ModB>(r, x, y); else if (which) fq_mul_rr<ModB>(r, x, y); else fq_mul<ModB>(r, x, y); }
        memcpy(out + 12 * i, r, 96);
    }
    return 0;
}
// which: 0 = fq_inv (fast path with fallback), 1 = plain binary gcd only, 2 = fast path only (returns the
// number of inputs on which the fast path reported failure); the plain-integer variants are wrapped the same
// way as fq_inv so that all three return the Montgomery-form inverse
int emu_fq_inv(int modulus, int which, size_t n, const uint64_t *a, uint64_t *out) {
    int failed = 0;
    for (size_t i = 0; i < n; ++i) {
        fq_t x, r, t, r2;
        memcpy(x, a + 12 * i, 96);
        if (which == 0) {
            if (modulus == 0) fq_inv<ModA>(r, x); else fq_inv<ModB>(r, x);
        } else {
            bool ok;
            if (modulus == 0) {
                ok = which == 1 ? fq_inv_plain<ModA>(t, x) : fq_inv_plain_fast<ModA>(t, x);
                for (int k = 0; k < NLIMB; ++k) r2[k] = ModA::R2(k);
                fq_mul<ModA>(t, t, r2); fq_mul<ModA>(r, t, r2);
            } else {
                ok = which == 1 ? fq_inv_plain<ModB>(t, x) : fq_inv_plain_fast<ModB>(t, x);
                for (int k = 0; k < NLIMB; ++k) r2[k] = ModB::R2(k);
                fq_mul<ModB>(t, t, r2); fq_mul<ModB>(r, t, r2);
            }
            if (!ok) ++failed;
        }
        memcpy(out + 12 * i, r, 96);
    }
    return failed;
}
int emu_fr_from_mont(int curve, size_t n, const uint64_t *in, uint64_t *out) {
    for (size_t i = 0; i < n; ++i) {
        fq_t x, r;
        memcpy(x, in + 12 * i, 96);
        if (curve == 0) fq_from_mont<ModB>(r, x); else fq_from_mont<ModA>(r, x);
        memcpy(out + 12 * i, r, 96);
    }
    return 0;
}
}
