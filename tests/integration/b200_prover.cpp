// INTEGRATION HARNESS (test infrastructure) -- the reference's GPU prover driver with its five
// multi-exponentiations routed through the C ABI of include/b200_msm.h.
//
// It is the patch INTEGRATION.md describes, made runnable: everything that is not an MSM -- parameter
// semantics, input parsing, the seven FFTs of compute_H, the final r*Bt1 + Lt + Ht, the affine
// conversion and the proof writer -- is the UNMODIFIED reference, compiled from /root/reference by
// oracle/Makefile (target `prover`); nothing of it is copied here.  The reference translation unit
// libsnark/prover_reference_functions.cpp is #included (not copied) because its vector_Fr / vector_G1
// handles are opaque outside it (prover_reference_include/prover_reference_functions.hpp:9-28) and the
// H-query scalars produced by compute_H must be handed to the engine as raw limbs.
//
//   b200_prover <MNT4753|MNT6753> compute <params> <input> <output> [n_gpus] [cpu-h|gpu-h] [repeats]
//
// By default the H polynomial is computed on the device too (b200msm_compute_h) and its
// coefficients never leave HBM before the H-query MSM; `cpu-h` keeps the reference's libfqfft path
// (compute_H, cuda_prover_piecewise.cu:14-49) for comparison.
//
// mirrors `cuda_prover_piecewise <curve> compute <params> <input> <output>` (cuda_prover_piecewise.cu:232-263)
// minus the preprocessed-table argument, which no longer exists.  tests/test_gpu_prover.py checks that the
// proof file is byte-identical (sha256) to the one written by the reference's CPU prover `main` on the same
// freshly generated parameters and input -- the acceptance procedure of the reference's README.md.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "b200_bundle.hpp"  // #includes the reference TU itself (see above) and include/b200_msm.h

// pinned host memory for the witness file (only cudaHostAlloc / cudaFreeHost of the CUDA runtime are used here;
// declared by hand so that the harness does not need the CUDA headers)
extern "C" int cudaHostAlloc(void **ptr, size_t size, unsigned int flags);
extern "C" int cudaFreeHost(void *ptr);

namespace {

typedef std::chrono::high_resolution_clock Clock;
double ms_since(Clock::time_point t) { return std::chrono::duration<double, std::milli>(Clock::now() - t).count(); }

using b200::CurveOf;

std::vector<char> slurp(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<char> buf((size_t)n);
    if (fread(buf.data(), 1, (size_t)n, f) != (size_t)n) { fprintf(stderr, "short read on %s\n", path); exit(2); }
    fclose(f);
    return buf;
}

#define CHECK(ctx, call)                                                                 \
    do {                                                                                 \
        int rc_ = (call);                                                                \
        if (rc_) { fprintf(stderr, "%s: %s\n", #call, b200msm_last_error(ctx)); abort(); } \
    } while (0)

using b200::compute_H;   // the seven FFTs of cuda_prover_piecewise.cu:14-49 through the bundle (b200_bundle.hpp)

struct Query {
    int group;
    const uint64_t *bases;  // affine wire format inside the parameter file image
    size_t n, point_words;
    std::vector<int> slot;              // per GPU
    std::vector<std::pair<size_t, size_t>> range;  // per GPU (offset, length)
    std::vector<uint64_t> partial;      // per GPU Jacobian X||Y||Z
    uint64_t result[108];
};

template <class B>
int run(const char *params_path, const char *input_path, const char *output_path, int n_gpus, bool cpu_h, int repeats) {
    typedef CurveOf<B> C;
    B::init_public_params();
    auto t_all = Clock::now();

    std::vector<char> pfile = slurp(params_path);
    const size_t d = ((const uint64_t *)pfile.data())[0], m = ((const uint64_t *)pfile.data())[1];
    printf("d = %zu, m = %zu, gpus = %d\n", d, m, n_gpus);
    const size_t g1w = 24, g2w = 24 * C::deg2;  // u64 words per affine point
    const uint64_t *p = (const uint64_t *)pfile.data() + 2;
    Query q[5];  // A, B1, B2, L, H   (layout: generate_parameters.cpp:59-108)
    q[0] = {B200MSM_G1, p, m + 1, g1w}; p += (m + 1) * g1w;
    q[1] = {B200MSM_G1, p, m + 1, g1w}; p += (m + 1) * g1w;
    q[2] = {B200MSM_G2, p, m + 1, g2w}; p += (m + 1) * g2w;
    q[3] = {B200MSM_G1, p, m - 1, g1w}; p += (m - 1) * g1w;
    q[4] = {B200MSM_G1, p, d, g1w};     p += d * g1w;
    if ((const char *)p != pfile.data() + pfile.size()) { fprintf(stderr, "parameter file size does not match d, m\n"); return 2; }

    // resident base sets, point-range sharded over the GPUs
    auto t = Clock::now();
    std::vector<b200msm_ctx *> ctx(n_gpus, nullptr);
    for (int g = 0; g < n_gpus; ++g)
        if (b200msm_create(C::id, g, &ctx[g])) { fprintf(stderr, "no usable CUDA device %d\n", g); return 3; }
    for (auto &Q : q) {
        Q.slot.resize(n_gpus);
        Q.range.resize(n_gpus);
        Q.partial.assign((size_t)n_gpus * 36 * (Q.group == B200MSM_G1 ? 1 : C::deg2), 0);
        for (int g = 0; g < n_gpus; ++g) {
            size_t off = 0, len = 0;
            b200msm_shard_range(Q.n, g, n_gpus, &off, &len);   // the one sharding rule of the ABI
            Q.range[g] = {off, len};
            CHECK(ctx[g], b200msm_bases_upload(ctx[g], Q.group, Q.bases + off * Q.point_words, len, &Q.slot[g]));
        }
    }
    printf("upload + window tables: %.1f ms\n", ms_since(t));

  // witness file image: w[0..m], ca, cb, cc (d + 1 each), r -- all Fr, Montgomery limbs (main.cpp:35-85)
  const size_t input_bytes = ((m + 1) + 3 * (d + 1) + 1) * 96;
  char *ibuf = nullptr;
  if (cudaHostAlloc((void **)&ibuf, input_bytes, 0) != 0) { fprintf(stderr, "cudaHostAlloc failed\n"); return 3; }
  for (int rep = 0; rep < repeats; ++rep) {   // a resident prover: later proofs reuse contexts, tables and buffers
    auto t_main = Clock::now();
    {
        FILE *f = fopen(input_path, "rb");
        if (!f || fread(ibuf, 1, input_bytes, f) != input_bytes) { fprintf(stderr, "cannot read %s\n", input_path); return 2; }
        fclose(f);
    }
    const uint64_t *w = (const uint64_t *)ibuf;
    typename B::groth16_input *inputs = nullptr;
    typename B::field r_field;
    if (cpu_h) {                               // the reference's parser is only needed for its own compute_H
        FILE *inputs_file = fopen(input_path, "r");
        inputs = B::read_input(inputs_file, d, m);
        fclose(inputs_file);
    }
    memcpy(r_field.data.mont_repr.data, ibuf + input_bytes - 96, 96);
    printf("load inputs: %.1f ms\n", ms_since(t_main));

    auto t_gpu = Clock::now();
    const size_t primary_input_size = 1;
    auto launch = [&](int qi, int lane, const uint64_t *scalars) {
        Query &Q = q[qi];
        const size_t jw = 36 * (Q.group == B200MSM_G1 ? 1 : C::deg2);
        for (int g = 0; g < n_gpus; ++g)
            CHECK(ctx[g], b200msm_msm_async(ctx[g], lane, Q.slot[g], 0, scalars + Q.range[g].first * 12, Q.range[g].second,
                                            Q.partial.data() + (size_t)g * jw));
    };
    auto finish = [&](int qi, int lane) {
        Query &Q = q[qi];
        for (int g = 0; g < n_gpus; ++g) CHECK(ctx[g], b200msm_wait(ctx[g], lane));
        CHECK(ctx[0], b200msm_fold(ctx[0], Q.group, Q.partial.data(), (size_t)n_gpus, Q.result));
    };
    launch(0, 0, w);                                    // A:  sum w_i A_i,            i in [0, m]
    launch(1, 1, w);                                    // B1
    launch(2, 2, w);                                    // B2 (G2)
    launch(3, 3, w + (primary_input_size + 1) * 12);    // L:  w[2..m]   (cuda_prover_piecewise.cu:167)
    const uint64_t *h_scalars = nullptr;
    std::vector<uint64_t> h_copy;
    if (cpu_h) {
        // CPU work overlaps the four MSMs, as in the reference (cuda_prover_piecewise.cu:174-179)
        auto coefficients_for_H = compute_H<B>(d, B::input_ca(inputs), B::input_cb(inputs), B::input_cc(inputs));
        printf("compute_H (CPU, libfqfft): %.1f ms\n", ms_since(t_gpu));
        h_scalars = (const uint64_t *)(coefficients_for_H->data->data() + coefficients_for_H->offset);
    } else {
        // ca, cb, cc follow w in the input file (main.cpp:35-85): d + 1 Fr elements each
        const uint64_t *ca = w + (m + 1) * 12, *cb = ca + (d + 1) * 12, *cc = cb + (d + 1) * 12;
        auto t_h = Clock::now();
        const uint64_t *h_dev = nullptr;
        CHECK(ctx[0], b200msm_compute_h(ctx[0], d, ca, cb, cc, n_gpus > 1 ? (h_copy.resize((d + 2) * 12), h_copy.data()) : nullptr, &h_dev));
        float hms[2];
        b200msm_compute_h_timings(ctx[0], hms);
        printf("compute_H (device): %.1f ms wall, %.2f ms device, domain tables %.2f ms\n", ms_since(t_h), hms[0], hms[1]);
        h_scalars = n_gpus > 1 ? h_copy.data() : h_dev;   // one GPU: the coefficients stay in HBM
    }
    finish(0, 0);
    launch(4, 0, h_scalars);                            // H:  d coefficients
    finish(1, 1);
    finish(2, 2);
    finish(3, 3);
    finish(4, 0);
    printf("gpu e2e (5 MSMs incl. compute_H overlap): %.1f ms\n", ms_since(t_gpu));

    auto evaluation_At = B::read_pt_ECp(q[0].result);   // prover_reference_functions.cpp:795-817
    auto evaluation_Bt1 = B::read_pt_ECp(q[1].result);
    auto evaluation_Bt2 = B::read_pt_ECpe(q[2].result);
    auto evaluation_Lt = B::read_pt_ECp(q[3].result);
    auto evaluation_Ht = B::read_pt_ECp(q[4].result);
    auto scaled_Bt1 = B::G1_scale(&r_field, evaluation_Bt1);                    // cuda_prover_piecewise.cu:198-204
    auto Lt1_plus_scaled_Bt1 = B::G1_add(evaluation_Lt, scaled_Bt1);
    auto final_C = B::G1_add(evaluation_Ht, Lt1_plus_scaled_Bt1);
    B::groth16_output_write(evaluation_At, evaluation_Bt2, final_C, output_path);
    printf("Total time from input to output: %.1f ms\n", ms_since(t_main));
    B::delete_G1(evaluation_At); B::delete_G1(evaluation_Bt1); B::delete_G2(evaluation_Bt2); B::delete_G1(evaluation_Lt);
    B::delete_G1(evaluation_Ht); B::delete_G1(scaled_Bt1); B::delete_G1(Lt1_plus_scaled_Bt1); B::delete_G1(final_C);
    if (inputs) B::delete_groth16_input(inputs);
  }
  cudaFreeHost(ibuf);
    printf("Total runtime (incl. file reads, uploads): %.1f ms\n", ms_since(t_all));
    for (auto c : ctx) b200msm_destroy(c);
    return 0;
}

}  // namespace

int main(int argc, char **argv) {
    setbuf(stdout, NULL);
    if (argc < 6 || std::string(argv[2]) != "compute") {
        fprintf(stderr, "usage: %s <MNT4753|MNT6753> compute <params> <input> <output> [n_gpus] [cpu-h|gpu-h] [repeats]\n", argv[0]);
        return 1;
    }
    const int n_gpus = argc > 6 ? atoi(argv[6]) : 1;
    const bool cpu_h = argc > 7 && std::string(argv[7]) == "cpu-h";
    const int repeats = argc > 8 ? atoi(argv[8]) : 1;
    const std::string curve(argv[1]);
    if (curve == "MNT4753") return run<mnt4753_libsnark>(argv[3], argv[4], argv[5], n_gpus, cpu_h, repeats);
    if (curve == "MNT6753") return run<mnt6753_libsnark>(argv[3], argv[4], argv[5], n_gpus, cpu_h, repeats);
    fprintf(stderr, "unknown curve %s\n", argv[1]);
    return 1;
}
