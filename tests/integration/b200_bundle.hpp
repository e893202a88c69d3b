// INTEGRATION HARNESS (test infrastructure) -- the libff-side adaptor of INTEGRATION.md section 3, compiled.
//
// b200_bundle<Base> is a drop-in for the reference's plugin bundles `mnt4753_libsnark` / `mnt6753_libsnark`
// (libsnark/prover_reference_include/prover_reference_functions.hpp:7-91,92-176): it inherits every static
// function of Base unchanged and replaces exactly two,
//
//     static G1 *multiexp_G1(vector_Fr *scalar_start, vector_G1 *g_start, size_t length);   hpp:59-60
//     static G2 *multiexp_G2(vector_Fr *scalar_start, vector_G2 *g_start, size_t length);   hpp:61-62
//
// (CPU: prover_reference_functions.cpp:350-368,690-708 -> libff::multi_exp_with_mixed_addition) by the engine
// behind include/b200_msm.h.  Any template of the reference that is written against a bundle type --
// run_prover<B> of cuda_prover_piecewise.cu:96-230, compute_H<B> -- works with b200_bundle<Base> as its B.
//
// What the adaptor does (the marshalling SURVEY.md 8(b) describes):
//   * scalars: std::vector<Fr> is a contiguous array of 96-byte Montgomery limbs (fp.hpp:40-42), so the
//     engine reads &(*scalar_start->data)[scalar_start->offset] in place -- no copy;
//   * bases: std::vector<libff::G1/G2> is a 288 / 576 / 864-byte-stride array of projective (X, Y, Z); the
//     engine's wire format is affine x || y with infinity encoded as y == 0 (serialization.hpp:87-89), and
//     libff's infinity is (0, 1, 0), so the adaptor tests Z == 0 (not y == 0) and emits zeros; a point with
//     Z != 1 is normalised first.  This happens ONCE per base vector: the uploaded base set (with its window
//     tables) is cached under the vector's storage address and reused by every later multiexp on it;
//   * result: the engine returns Jacobian X || Y || Z, imported through the reference's own
//     B::read_pt_ECp / read_pt_ECpe (prover_reference_functions.cpp:795-817).
// With several GPUs every base vector is sharded by point range and the partial points are folded
// on GPU 0.  There is no CPU fallback: an engine error aborts, as the reference does on a failed allocation.
//
// The reference translation unit is #included (not copied) because the handle types vector_Fr / vector_G1 / G1
// are only *declared* in the header (hpp:9-28); their definitions live in prover_reference_functions.cpp:208-234.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <utility>
#include <vector>

#include "libsnark/prover_reference_functions.cpp"  // the reference TU itself (see above)

#include "b200_msm.h"

// the only CUDA runtime symbol the adaptor needs (declared by hand: no CUDA headers in a libff translation unit)
extern "C" int cudaGetDeviceCount(int *count);

namespace b200 {

template <class Base> struct CurveOf;
template <> struct CurveOf<mnt4753_libsnark> { static constexpr int id = B200MSM_MNT4753, deg2 = 2; };
template <> struct CurveOf<mnt6753_libsnark> { static constexpr int id = B200MSM_MNT6753, deg2 = 3; };

// A libff coordinate (Fq, Fq2 = {c0, c1}, Fq3 = {c0, c1, c2}) is a plain aggregate of 96-byte Montgomery limb
// arrays (fp.hpp:40-42, fp2.hpp:49, fp3.hpp:49), i.e. already the wire format of one coordinate
// (extension components contiguous, serialization.hpp:35-67).
template <class F> inline void put(uint64_t *&o, const F &x) {
    static_assert(sizeof(F) % 96 == 0 && sizeof(F) <= 288, "coordinate = 1..3 Fq elements of 12 x u64");
    memcpy(o, &x, sizeof(F));
    o += sizeof(F) / 8;
}

// one engine context per GPU and the cache of uploaded base vectors, per curve
template <class Base>
struct Engine {
    struct Shard { int slot; size_t off, len; };
    struct Entry { size_t n; std::vector<Shard> shards; };
    std::vector<b200msm_ctx *> ctx;
    std::map<const void *, Entry> cache;   // key: address of the vector's first element
    size_t uploads = 0, hits = 0;

    static Engine &get() { static Engine e; return e; }

    void die(b200msm_ctx *c, const char *what, int rc) {
        fprintf(stderr, "b200_bundle: %s failed (%d): %s\n", what, rc, c ? b200msm_last_error(c) : "no context");
        abort();
    }
    void open(int n_gpus) {
        if (!ctx.empty()) return;
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != 0 || ndev < 1) die(nullptr, "cudaGetDeviceCount", 2);
        for (int g = 0; g < n_gpus; ++g) {
            b200msm_ctx *c = nullptr;
            // more shards than devices (tests on a one-GPU box): contexts share a device, the sharding is the same
            int rc = b200msm_create(CurveOf<Base>::id, g % ndev, &c);
            if (rc) die(nullptr, "b200msm_create", rc);
            ctx.push_back(c);
        }
    }
    void close() {
        for (auto c : ctx) b200msm_destroy(c);
        ctx.clear();
        cache.clear();
    }

    // std::vector<libff::G> (projective, 3 coordinates) -> affine wire format, uploaded and sharded once
    template <class G>
    const Entry &bases(const std::vector<G> &v, int group) {
        if (ctx.empty()) open(1);
        auto it = cache.find(v.data());
        if (it != cache.end() && it->second.n == v.size()) { ++hits; return it->second; }
        const size_t pw = 24 * (group == B200MSM_G1 ? 1 : CurveOf<Base>::deg2);
        std::vector<uint64_t> wire(v.size() * pw, 0);
        for (size_t i = 0; i < v.size(); ++i) {
            const auto Z = v[i].Z();
            if (Z.is_zero()) continue;                        // libff's infinity (0, 1, 0) -> x = y = 0
            uint64_t *o = wire.data() + i * pw;
            if (Z == decltype(Z)::one()) { put(o, v[i].X()); put(o, v[i].Y()); }
            else { G t = v[i]; t.to_affine_coordinates(); put(o, t.X()); put(o, t.Y()); }
        }
        Entry e;
        e.n = v.size();
        const int G_ = (int)ctx.size();
        for (int g = 0; g < G_; ++g) {
            // the rule of b200msm_shard_range (include/b200_msm.h): [n g / G, n (g + 1) / G)
            size_t lo, len;
            b200msm_shard_range(e.n, g, G_, &lo, &len);
            Shard s{-1, lo, len};
            int rc = b200msm_bases_upload(ctx[g], group, wire.data() + lo * pw, len, &s.slot);
            if (rc) die(ctx[g], "b200msm_bases_upload", rc);
            e.shards.push_back(s);
        }
        ++uploads;
        return cache[v.data()] = e;
    }

    // sum_{i < length} scalars[i] * v[i]  ->  Jacobian X || Y || Z in out (36 * deg words)
    template <class G>
    void msm(const uint64_t *scalars, const std::vector<G> &v, size_t length, int group, uint64_t *out) {
        if (length > v.size()) { fprintf(stderr, "b200_bundle: multiexp of %zu terms over %zu bases\n", length, v.size()); abort(); }
        const Entry &e = bases(v, group);
        const size_t jw = 36 * (group == B200MSM_G1 ? 1 : CurveOf<Base>::deg2);
        std::vector<uint64_t> partial(e.shards.size() * jw, 0);
        for (size_t g = 0; g < e.shards.size(); ++g) {
            const Shard &s = e.shards[g];
            const size_t n = length > s.off ? std::min(s.len, length - s.off) : 0;
            int rc = b200msm_msm_async(ctx[g], 0, s.slot, 0, scalars + s.off * 12, n, partial.data() + g * jw);
            if (rc) die(ctx[g], "b200msm_msm_async", rc);
        }
        for (size_t g = 0; g < e.shards.size(); ++g) {
            int rc = b200msm_wait(ctx[g], 0);
            if (rc) die(ctx[g], "b200msm_wait", rc);
        }
        if (e.shards.size() == 1) { memcpy(out, partial.data(), jw * 8); return; }
        int rc = b200msm_fold(ctx[0], group, partial.data(), e.shards.size(), out);
        if (rc) die(ctx[0], "b200msm_fold", rc);
    }
};

// the seven FFTs of compute_H<B> (cuda_prover_piecewise.cu:14-49; a template inside the reference's .cu file, which
// a C++ harness cannot include), through the bundle's own functions
template <class B>
typename B::vector_Fr *compute_H(size_t d, typename B::vector_Fr *ca, typename B::vector_Fr *cb, typename B::vector_Fr *cc) {
    auto domain = B::get_evaluation_domain(d + 1);
    B::domain_iFFT(domain, ca);
    B::domain_iFFT(domain, cb);
    B::domain_cosetFFT(domain, ca);
    B::domain_cosetFFT(domain, cb);
    size_t m = B::domain_get_m(domain);
    B::vector_Fr_muleq(ca, cb, m);
    B::domain_iFFT(domain, cc);
    B::domain_cosetFFT(domain, cc);
    B::vector_Fr_subeq(ca, cc, m);
    B::domain_divide_by_Z_on_coset(domain, ca);
    B::domain_icosetFFT(domain, ca);
    typename B::vector_Fr *res = B::vector_Fr_zeros(m + 1);
    B::vector_Fr_copy_into(ca, res, m);
    return res;
}

}  // namespace b200

// The bundle: Base with its two multiexps on the engine.  Everything else is inherited from the reference.
template <class Base>
class b200_bundle : public Base {
public:
    typedef typename Base::G1 G1;
    typedef typename Base::G2 G2;
    typedef typename Base::vector_Fr vector_Fr;
    typedef typename Base::vector_G1 vector_G1;
    typedef typename Base::vector_G2 vector_G2;

    // number of GPUs the base vectors are sharded over (call before the first multiexp; default 1)
    static void engine_open(int n_gpus) { b200::Engine<Base>::get().open(n_gpus); }
    static void engine_close() { b200::Engine<Base>::get().close(); }
    static size_t engine_uploads() { return b200::Engine<Base>::get().uploads; }
    static size_t engine_cache_hits() { return b200::Engine<Base>::get().hits; }

    // prover_reference_functions.hpp:59-60 / .cpp:350-358 (mnt4753), :690-698 (mnt6753)
    static G1 *multiexp_G1(vector_Fr *scalar_start, vector_G1 *g_start, size_t length) {
        uint64_t xyz[36];
        const uint64_t *s = reinterpret_cast<const uint64_t *>(scalar_start->data->data() + scalar_start->offset);
        b200::Engine<Base>::get().msm(s, *g_start->data, length, B200MSM_G1, xyz);
        return Base::read_pt_ECp(xyz);        // .cpp:795-817: Jacobian -> libff projective
    }
    // prover_reference_functions.hpp:61-62 / .cpp:359-368, :699-708
    static G2 *multiexp_G2(vector_Fr *scalar_start, vector_G2 *g_start, size_t length) {
        uint64_t xyz[108];
        const uint64_t *s = reinterpret_cast<const uint64_t *>(scalar_start->data->data() + scalar_start->offset);
        b200::Engine<Base>::get().msm(s, *g_start->data, length, B200MSM_G2, xyz);
        return Base::read_pt_ECpe(xyz);
    }
};
