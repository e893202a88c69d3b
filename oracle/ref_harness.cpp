// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product path.
//
// C-ABI harness around the *unmodified* reference CPU implementation (libff as vendored under
// /root/reference/depends/libff).  oracle/Makefile compiles this file together with the
// reference's own sources, where they lie, into oracle/_ref/libref.so.  It is used to
//   (1) pin the plain-C restatement in oracle/mnt753_oracle.c (tests/golden/ is generated from it),
//   (2) serve as the "reference" CPU baseline timed by bench.py (kind = "reference").
//
// Every entry point takes/returns the reference's wire format (libsnark/serialization.hpp:24-121):
//   Fq/Fr element  = 12 x u64 little-endian limbs, Montgomery form (R = 2^768), canonical (< p)
//   Fqe element    = DEG consecutive Fq (c0, c1[, c2])
//   affine point   = x || y ; infinity is the all-zero encoding and is recognised by y == 0
//   jacobian point = X || Y || Z as written by the reference GPU kernels (multiexp/curves.cu:104-114),
//                    converted exactly like prover_reference_functions.cpp:106-115 (x*z, y, z^3).
//
// curve: 0 = MNT4753, 1 = MNT6753.   group: 1 = G1, 2 = G2.
#include <cstdint>
#include <cstring>
#include <vector>

#include <fcntl.h>
#include <omp.h>
#include <unistd.h>

#include <libff/algebra/curves/mnt753/mnt4753/mnt4753_pp.hpp>
#include <libff/algebra/curves/mnt753/mnt6753/mnt6753_pp.hpp>
#include <libff/algebra/scalar_multiplication/multiexp.hpp>
#include <libff/common/profiling.hpp>
#include <libff/common/rng.hpp>
#include <libfqfft/evaluation_domain/get_evaluation_domain.hpp>

using namespace libff;

namespace {

constexpr size_t L = 12;  // u64 limbs per base-field element

template <typename F>
F rd_fp(const uint64_t *&p) {
    F x;
    memcpy(x.mont_repr.data, p, L * 8);
    p += L;
    return x;
}
template <typename F>
void wr_fp(uint64_t *&p, const F &x) {
    memcpy(p, x.mont_repr.data, L * 8);
    p += L;
}

// generic extension-field codec through all_base_field_elements()/vector ctor, as
// libsnark/serialization.hpp:35-41,92-100 does.
template <typename ppT>
struct codec {
    typedef Fq<ppT> FqT;
    typedef Fqe<ppT> FqeT;
    typedef Fr<ppT> FrT;
    typedef G1<ppT> G1T;
    typedef G2<ppT> G2T;
    static constexpr size_t DEG = sizeof(FqeT) / sizeof(FqT);

    static FqeT rd_fqe(const uint64_t *&p) {
        std::vector<FqT> v;
        for (size_t i = 0; i < DEG; ++i) v.emplace_back(rd_fp<FqT>(p));
        return FqeT(v);
    }
    static void wr_fqe(uint64_t *&p, FqeT x) {
        std::vector<FqT> v = x.all_base_field_elements();  // non-const member in libff
        for (size_t i = 0; i < DEG; ++i) wr_fp<FqT>(p, v[i]);
    }
    static FqT rd(const uint64_t *&p, FqT *) { return rd_fp<FqT>(p); }
    static FqeT rd(const uint64_t *&p, FqeT *) { return rd_fqe(p); }
    static void wr(uint64_t *&p, const FqT &x) { wr_fp<FqT>(p, x); }
    static void wr(uint64_t *&p, const FqeT &x) { wr_fqe(p, x); }
};

template <typename ppT, typename G>
struct grp {
    typedef decltype(G::zero().X()) FF;
    typedef codec<ppT> C;
    static constexpr size_t DEG = sizeof(FF) / sizeof(Fq<ppT>);
    static constexpr size_t AFF = 2 * DEG * L;  // u64 per affine point
    static constexpr size_t JAC = 3 * DEG * L;

    // serialization.hpp read_g1/read_g2: y == 0 => zero
    static G rd_affine(const uint64_t *p) {
        FF x = C::rd(p, (FF *)nullptr);
        FF y = C::rd(p, (FF *)nullptr);
        if (y == FF::zero()) return G::zero();
        return G(x, y, FF::one());
    }
    // serialization.hpp write_g1/write_g2
    static void wr_affine(uint64_t *p, G g) {
        if (g.is_zero()) {
            memset(p, 0, AFF * 8);
            return;
        }
        g.to_affine_coordinates();
        C::wr(p, g.X());
        C::wr(p, g.Y());
    }
    // prover_reference_functions.cpp:106-115 read_pt (Jacobian -> libff projective)
    static G rd_jacobian(const uint64_t *p) {
        FF x = C::rd(p, (FF *)nullptr);
        FF y = C::rd(p, (FF *)nullptr);
        FF z = C::rd(p, (FF *)nullptr);
        return G(x * z, y, z * z * z);
    }
};

template <typename ppT, typename G>
double msm_impl(size_t n, const uint64_t *bases, const uint64_t *scalars, uint64_t *out, int method,
                int chunks, int prefilter) {
    typedef grp<ppT, G> GG;
    typedef Fr<ppT> FrT;
    std::vector<G> g(n);
    std::vector<FrT> s(n);
#pragma omp parallel for
    for (size_t i = 0; i < n; ++i) {
        g[i] = GG::rd_affine(bases + i * GG::AFF);
        const uint64_t *p = scalars + i * L;
        s[i] = rd_fp<FrT>(p);
    }
    if (chunks <= 0) chunks = omp_get_max_threads();
    G res;
    // multi_exp_with_mixed_addition printf()s three statistics lines unconditionally
    // (multiexp.tcc:489-491); park stdout on /dev/null for the duration of the call.
    fflush(stdout);
    int saved = dup(1), nul = open("/dev/null", O_WRONLY);
    dup2(nul, 1);
    long long t0 = get_nsec_time();
#define RUN(M)                                                                                              \
    res = prefilter ? multi_exp_with_mixed_addition<G, FrT, M>(g.begin(), g.end(), s.begin(), s.end(), chunks) \
                    : multi_exp<G, FrT, M>(g.begin(), g.end(), s.begin(), s.end(), chunks)
    if (method == 0) {
        RUN(multi_exp_method_naive);
    } else if (method == 1) {
        RUN(multi_exp_method_BDLO12);
    } else {
        RUN(multi_exp_method_bos_coster);
    }
#undef RUN
    long long t1 = get_nsec_time();
    fflush(stdout);
    dup2(saved, 1);
    close(saved);
    close(nul);
    GG::wr_affine(out, res);
    return (double)(t1 - t0) * 1e-9;
}

template <typename ppT, typename G>
void gen_bases_impl(size_t n, uint64_t seed_p0, uint64_t seed_q, uint64_t *out) {
    typedef grp<ppT, G> GG;
    typedef Fr<ppT> FrT;
    const G P0 = SHA512_rng<FrT>(seed_p0) * G::one();
    G Q = SHA512_rng<FrT>(seed_q) * G::one();
    Q.to_special();
    int T = omp_get_max_threads();
    size_t per = (n + T - 1) / T;
#pragma omp parallel for
    for (int t = 0; t < T; ++t) {
        size_t lo = (size_t)t * per, hi = std::min(n, lo + per);
        if (lo >= hi) continue;
        std::vector<G> v(hi - lo);
        G cur = P0 + FrT((long)lo) * Q;
        for (size_t i = lo; i < hi; ++i) {
            v[i - lo] = cur;
            cur = cur + Q;
        }
        batch_to_special<G>(v);
        for (size_t i = lo; i < hi; ++i) GG::wr_affine(out + i * GG::AFF, v[i - lo]);
    }
}

template <typename ppT, typename G>
void point_op_impl(int op, const uint64_t *a, const uint64_t *b, const uint64_t *k, uint64_t *out) {
    typedef grp<ppT, G> GG;
    typedef Fr<ppT> FrT;
    G A = GG::rd_affine(a);
    G R;
    switch (op) {
        case 0: R = A + GG::rd_affine(b); break;                 // operator+
        case 1: R = A.dbl(); break;
        case 2: R = A.mixed_add(GG::rd_affine(b)); break;
        case 3: { const uint64_t *p = k; R = rd_fp<FrT>(p) * A; } break;  // scalar (Montgomery Fr) * A
        case 4: R = -A; break;
        case 5: R = A.add(GG::rd_affine(b)); break;
        default: R = G::zero();
    }
    GG::wr_affine(out, R);
}

template <typename F, typename Cd>
void field_op_impl(int op, size_t n, size_t stride, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    for (size_t i = 0; i < n; ++i) {
        const uint64_t *pa = a + i * stride, *pb = b ? b + i * stride : nullptr;
        uint64_t *po = out + i * stride;
        F x = Cd::rd(pa, (F *)nullptr);
        F y = pb ? Cd::rd(pb, (F *)nullptr) : F::zero();
        F r;
        switch (op) {
            case 0: r = x * y; break;
            case 1: r = x + y; break;
            case 2: r = x - y; break;
            case 3: r = x.squared(); break;
            case 4: r = x.is_zero() ? F::zero() : x.inverse(); break;
            case 5: r = -x; break;
            default: r = F::zero();
        }
        Cd::wr(po, r);
    }
}

template <typename ppT>
int dispatch_field(int which, int op, size_t n, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    typedef codec<ppT> C;
    if (which == 0) field_op_impl<Fq<ppT>, C>(op, n, L, a, b, out);
    else if (which == 1) field_op_impl<Fqe<ppT>, C>(op, n, L * C::DEG, a, b, out);
    else return -1;
    return 0;
}

bool g_init = false;

}  // namespace

extern "C" {

void ref_init(void) {
    if (g_init) return;
    inhibit_profiling_info = true;
    inhibit_profiling_counters = true;
    mnt4753_pp::init_public_params();
    mnt6753_pp::init_public_params();
    g_init = true;
}

int ref_num_threads(void) { return omp_get_max_threads(); }
void ref_set_num_threads(int t) { omp_set_num_threads(t); }

// which: 0 = Fq modulus, 1 = Fr modulus, 2 = Fq R^2, 3 = Fr R^2, 4 = G1 coeff_a (Montgomery),
//        5 = G1 generator affine (2*12), 6 = G2 coeff_a (DEG*12), 7 = G2 generator affine (2*DEG*12),
//        8 = Fq inv (1 limb), 9 = Fr inv (1 limb), 10 = Fqe non_residue (Montgomery, 12)
#define CONSTS(ppT)                                                                                \
    {                                                                                              \
        typedef codec<ppT> C;                                                                      \
        uint64_t *p = out;                                                                         \
        switch (which) {                                                                           \
            case 0: memcpy(out, Fq<ppT>::mod.data, 96); return 12;                                 \
            case 1: memcpy(out, Fr<ppT>::mod.data, 96); return 12;                                 \
            case 2: memcpy(out, Fq<ppT>::Rsquared.data, 96); return 12;                            \
            case 3: memcpy(out, Fr<ppT>::Rsquared.data, 96); return 12;                            \
            case 4: C::wr(p, G1<ppT>::coeff_a); return 12;                                         \
            case 5: grp<ppT, G1<ppT>>::wr_affine(out, G1<ppT>::one()); return 24;                  \
            case 6: C::wr(p, G2<ppT>::coeff_a); return (int)(12 * C::DEG);                         \
            case 7: grp<ppT, G2<ppT>>::wr_affine(out, G2<ppT>::one()); return (int)(24 * C::DEG);  \
            case 8: out[0] = Fq<ppT>::inv; return 1;                                               \
            case 9: out[0] = Fr<ppT>::inv; return 1;                                               \
            case 10: C::wr(p, Fqe<ppT>::non_residue); return 12;                                   \
        }                                                                                          \
        return -1;                                                                                 \
    }

int ref_get_constant(int curve, int which, uint64_t *out) {
    ref_init();
    if (curve == 0) CONSTS(mnt4753_pp) else CONSTS(mnt6753_pp)
}

// field: 0 = Fq, 1 = Fqe (G2 coordinate field).  op: 0 mul, 1 add, 2 sub, 3 sqr, 4 inv (0 -> 0), 5 neg.
int ref_field_op(int curve, int field, int op, size_t n, const uint64_t *a, const uint64_t *b, uint64_t *out) {
    ref_init();
    return curve == 0 ? dispatch_field<mnt4753_pp>(field, op, n, a, b, out)
                      : dispatch_field<mnt6753_pp>(field, op, n, a, b, out);
}

// scalar-field helpers: Montgomery <-> plain integer limbs (Fp_model::as_bigint / Fp_model(bigint))
int ref_fr_from_mont(int curve, size_t n, const uint64_t *in, uint64_t *out) {
    ref_init();
    for (size_t i = 0; i < n; ++i) {
        const uint64_t *p = in + i * L;
        if (curve == 0) { auto b = rd_fp<Fr<mnt4753_pp>>(p).as_bigint(); memcpy(out + i * L, b.data, 96); }
        else { auto b = rd_fp<Fr<mnt6753_pp>>(p).as_bigint(); memcpy(out + i * L, b.data, 96); }
    }
    return 0;
}
int ref_fr_to_mont(int curve, size_t n, const uint64_t *in, uint64_t *out) {
    ref_init();
    for (size_t i = 0; i < n; ++i) {
        uint64_t *o = out + i * L;
        if (curve == 0) { bigint<12> b; memcpy(b.data, in + i * L, 96); wr_fp(o, Fr<mnt4753_pp>(b)); }
        else { bigint<12> b; memcpy(b.data, in + i * L, 96); wr_fp(o, Fr<mnt6753_pp>(b)); }
    }
    return 0;
}

// scalars[i] = SHA512_rng<Fr>(seed * 2^32 + i), Montgomery limbs (rng.tcc:26-80)
int ref_gen_scalars(int curve, size_t n, uint64_t seed, uint64_t *out) {
    ref_init();
#pragma omp parallel for
    for (size_t i = 0; i < n; ++i) {
        uint64_t *o = out + i * L;
        if (curve == 0) wr_fp(o, SHA512_rng<Fr<mnt4753_pp>>((seed << 32) + i));
        else wr_fp(o, SHA512_rng<Fr<mnt6753_pp>>((seed << 32) + i));
    }
    return 0;
}

// bases[i] = P0 + i*Q, P0 = SHA512_rng(seed_p0)*G, Q = SHA512_rng(seed_q)*G; affine wire format.
int ref_gen_bases(int curve, int group, size_t n, uint64_t seed_p0, uint64_t seed_q, uint64_t *out) {
    ref_init();
    if (curve == 0 && group == 1) gen_bases_impl<mnt4753_pp, G1<mnt4753_pp>>(n, seed_p0, seed_q, out);
    else if (curve == 0 && group == 2) gen_bases_impl<mnt4753_pp, G2<mnt4753_pp>>(n, seed_p0, seed_q, out);
    else if (curve == 1 && group == 1) gen_bases_impl<mnt6753_pp, G1<mnt6753_pp>>(n, seed_p0, seed_q, out);
    else if (curve == 1 && group == 2) gen_bases_impl<mnt6753_pp, G2<mnt6753_pp>>(n, seed_p0, seed_q, out);
    else return -1;
    return 0;
}

// method: 0 naive, 1 BDLO12 (what ./main uses), 2 bos_coster (what prover_reference_functions.cpp uses).
// prefilter != 0 -> multi_exp_with_mixed_addition (the call ./main makes, main.cpp:150-170).
// chunks <= 0 -> omp_get_max_threads().  Returns seconds spent inside the libff call; result affine in out.
double ref_msm(int curve, int group, size_t n, const uint64_t *bases, const uint64_t *scalars, uint64_t *out,
               int method, int chunks, int prefilter) {
    ref_init();
    if (curve == 0 && group == 1) return msm_impl<mnt4753_pp, G1<mnt4753_pp>>(n, bases, scalars, out, method, chunks, prefilter);
    if (curve == 0 && group == 2) return msm_impl<mnt4753_pp, G2<mnt4753_pp>>(n, bases, scalars, out, method, chunks, prefilter);
    if (curve == 1 && group == 1) return msm_impl<mnt6753_pp, G1<mnt6753_pp>>(n, bases, scalars, out, method, chunks, prefilter);
    if (curve == 1 && group == 2) return msm_impl<mnt6753_pp, G2<mnt6753_pp>>(n, bases, scalars, out, method, chunks, prefilter);
    return -1.0;
}

// op: 0 a+b, 1 dbl(a), 2 a.mixed_add(b), 3 k*a (k = Montgomery Fr), 4 -a, 5 a.add(b).  All affine in/out.
int ref_point_op(int curve, int group, int op, const uint64_t *a, const uint64_t *b, const uint64_t *k, uint64_t *out) {
    ref_init();
    if (curve == 0 && group == 1) point_op_impl<mnt4753_pp, G1<mnt4753_pp>>(op, a, b, k, out);
    else if (curve == 0 && group == 2) point_op_impl<mnt4753_pp, G2<mnt4753_pp>>(op, a, b, k, out);
    else if (curve == 1 && group == 1) point_op_impl<mnt6753_pp, G1<mnt6753_pp>>(op, a, b, k, out);
    else if (curve == 1 && group == 2) point_op_impl<mnt6753_pp, G2<mnt6753_pp>>(op, a, b, k, out);
    else return -1;
    return 0;
}

// The reference's device-result import: (X,Y,Z) Jacobian -> libff projective -> affine wire bytes,
// i.e. read_pt_ECp/ECpe followed by write_g1/write_g2.
int ref_jacobian_to_affine(int curve, int group, const uint64_t *xyz, uint64_t *out) {
    ref_init();
    if (curve == 0 && group == 1) grp<mnt4753_pp, G1<mnt4753_pp>>::wr_affine(out, grp<mnt4753_pp, G1<mnt4753_pp>>::rd_jacobian(xyz));
    else if (curve == 0 && group == 2) grp<mnt4753_pp, G2<mnt4753_pp>>::wr_affine(out, grp<mnt4753_pp, G2<mnt4753_pp>>::rd_jacobian(xyz));
    else if (curve == 1 && group == 1) grp<mnt6753_pp, G1<mnt6753_pp>>::wr_affine(out, grp<mnt6753_pp, G1<mnt6753_pp>>::rd_jacobian(xyz));
    else if (curve == 1 && group == 2) grp<mnt6753_pp, G2<mnt6753_pp>>::wr_affine(out, grp<mnt6753_pp, G2<mnt6753_pp>>::rd_jacobian(xyz));
    else return -1;
    return 0;
}

}  // extern "C"

// compute_H exactly as the reference's GPU prover driver does it (cuda_prover_piecewise.cu:14-49), through
// libfqfft's own domain object (prover_reference_functions.cpp: domain_iFFT / domain_cosetFFT / ...).
namespace {
template <typename ppT>
int compute_h_ref(size_t d, const uint64_t *ca, const uint64_t *cb, const uint64_t *cc, uint64_t *out) {
    typedef Fr<ppT> F;
    const size_t m = d + 1;
    std::vector<F> A(m), B(m), C(m);
    for (size_t i = 0; i < m; ++i) {
        const uint64_t *p = ca + i * L; A[i] = rd_fp<F>(p);
        p = cb + i * L; B[i] = rd_fp<F>(p);
        p = cc + i * L; C[i] = rd_fp<F>(p);
    }
    auto domain = libfqfft::get_evaluation_domain<F>(d + 1);
    if (domain->m != m) return -2;
    domain->iFFT(A);
    domain->iFFT(B);
    domain->cosetFFT(A, F::multiplicative_generator);
    domain->cosetFFT(B, F::multiplicative_generator);
    for (size_t i = 0; i < m; ++i) A[i] *= B[i];
    domain->iFFT(C);
    domain->cosetFFT(C, F::multiplicative_generator);
    for (size_t i = 0; i < m; ++i) A[i] -= C[i];
    domain->divide_by_Z_on_coset(A);
    domain->icosetFFT(A, F::multiplicative_generator);
    for (size_t i = 0; i < m; ++i) { uint64_t *o = out + i * L; wr_fp(o, A[i]); }
    memset(out + m * L, 0, L * 8);
    return 0;
}
}  // namespace

extern "C" int ref_compute_h(int curve, size_t d, const uint64_t *ca, const uint64_t *cb, const uint64_t *cc, uint64_t *out) {
    ref_init();
    return curve == 0 ? compute_h_ref<mnt4753_pp>(d, ca, cb, cc, out) : compute_h_ref<mnt6753_pp>(d, ca, cb, cc, out);
}

