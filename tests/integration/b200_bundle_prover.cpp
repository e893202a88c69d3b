// INTEGRATION HARNESS (test infrastructure) -- the reference's prover call sequence, written ONLY against the
// plugin-bundle API (prover_reference_include/prover_reference_functions.hpp), instantiated with
// B = b200_bundle<mnt4753_libsnark / mnt6753_libsnark> (b200_bundle.hpp).
//
// run_prover<B> below is the call sequence of run_prover<B> in cuda_prover_piecewise.cu:96-230 with the three
// ec_reduce_straus launches (:165-167) written in the bundle form the reference itself keeps next to them
// (:170-171, commented out: `B::multiexp_G1(B::input_w(inputs), B::params_B1(params), m + 1)`), so that ALL five
// multi-exponentiations go through B::multiexp_G1 / B::multiexp_G2 -- the SURVEY.md 8(a) row a11 boundary.
// Everything else -- B::read_params, B::read_input, the FFTs of compute_H, G1_scale, G1_add, groth16_output_write --
// is the unmodified reference.  With B = mnt4753_libsnark the very same function is the reference CPU prover; with
// B = b200_bundle<...> the multiexps run on the GPU and tests/test_gpu_prover.py requires byte-identical proofs.
//
//   b200_bundle_prover <MNT4753|MNT6753> compute <params> <input> <output> [n_gpus] [cpu|b200] [repeats]
#include <chrono>
#include <string>

#include "b200_bundle.hpp"

namespace {
typedef std::chrono::high_resolution_clock Clock;
double ms_since(Clock::time_point t) { return std::chrono::duration<double, std::milli>(Clock::now() - t).count(); }

template <class B>
void run_prover(const char *params_path, const char *input_path, const char *output_path, int repeats) {
    B::init_public_params();
    const size_t primary_input_size = 1;

    FILE *params_file = fopen(params_path, "r");
    if (!params_file) { fprintf(stderr, "cannot open %s\n", params_path); exit(2); }
    size_t d = 0, m = 0;
    if (fread(&d, 8, 1, params_file) != 1 || fread(&m, 8, 1, params_file) != 1) { fprintf(stderr, "short read\n"); exit(2); }
    rewind(params_file);
    printf("d = %zu, m = %zu\n", d, m);
    auto t = Clock::now();
    auto params = B::read_params(params_file, d, m);
    fclose(params_file);
    printf("load params: %.1f ms\n", ms_since(t));

    for (int rep = 0; rep < repeats; ++rep) {
        auto t_main = Clock::now();
        FILE *inputs_file = fopen(input_path, "r");
        if (!inputs_file) { fprintf(stderr, "cannot open %s\n", input_path); exit(2); }
        auto inputs = B::read_input(inputs_file, d, m);
        fclose(inputs_file);

        auto t_msm = Clock::now();
        auto A = B::params_A(params);
        auto B1 = B::params_B1(params);
        auto B2 = B::params_B2(params);
        auto L = B::params_L(params);
        auto H = B::params_H(params);
        typename B::G1 *evaluation_At = B::multiexp_G1(B::input_w(inputs), A, m + 1);
        typename B::G1 *evaluation_Bt1 = B::multiexp_G1(B::input_w(inputs), B1, m + 1);
        typename B::G2 *evaluation_Bt2 = B::multiexp_G2(B::input_w(inputs), B2, m + 1);
        auto w_L = B::vector_Fr_offset(B::input_w(inputs), primary_input_size + 1);
        typename B::G1 *evaluation_Lt = B::multiexp_G1(w_L, L, m - 1);
        double ms_msm = ms_since(t_msm);

        auto t_h = Clock::now();
        auto coefficients_for_H = b200::compute_H<B>(d, B::input_ca(inputs), B::input_cb(inputs), B::input_cc(inputs));
        printf("compute_H (CPU, libfqfft): %.1f ms\n", ms_since(t_h));
        t_msm = Clock::now();
        typename B::G1 *evaluation_Ht = B::multiexp_G1(coefficients_for_H, H, d);
        ms_msm += ms_since(t_msm);
        printf("five multiexps through B::multiexp_G1/G2: %.1f ms\n", ms_msm);

        auto scaled_Bt1 = B::G1_scale(B::input_r(inputs), evaluation_Bt1);
        auto Lt1_plus_scaled_Bt1 = B::G1_add(evaluation_Lt, scaled_Bt1);
        auto final_C = B::G1_add(evaluation_Ht, Lt1_plus_scaled_Bt1);
        B::groth16_output_write(evaluation_At, evaluation_Bt2, final_C, output_path);
        printf("Total time from input to output: %.1f ms\n", ms_since(t_main));

        B::delete_vector_G1(A); B::delete_vector_G1(B1); B::delete_vector_G2(B2); B::delete_vector_G1(L); B::delete_vector_G1(H);
        B::delete_G1(evaluation_At); B::delete_G1(evaluation_Bt1); B::delete_G2(evaluation_Bt2);
        B::delete_G1(evaluation_Ht); B::delete_G1(evaluation_Lt); B::delete_G1(scaled_Bt1);
        B::delete_G1(Lt1_plus_scaled_Bt1); B::delete_G1(final_C);
        B::delete_vector_Fr(w_L);
        B::delete_vector_Fr(coefficients_for_H);
        B::delete_groth16_input(inputs);
    }
    B::delete_groth16_params(params);
}

template <class Base>
int run(const char *params, const char *input, const char *output, int n_gpus, bool cpu, int repeats) {
    if (cpu) { run_prover<Base>(params, input, output, repeats); return 0; }   // the reference bundle itself
    typedef b200_bundle<Base> B;
    B::engine_open(n_gpus);
    run_prover<B>(params, input, output, repeats);
    // five base vectors are uploaded once; every later multiexp on them is a cache hit
    printf("base-set uploads: %zu, cache hits: %zu\n", B::engine_uploads(), B::engine_cache_hits());
    B::engine_close();
    return 0;
}
}  // namespace

int main(int argc, char **argv) {
    setbuf(stdout, NULL);
    if (argc < 6 || std::string(argv[2]) != "compute") {
        fprintf(stderr, "usage: %s <MNT4753|MNT6753> compute <params> <input> <output> [n_gpus] [cpu|b200] [repeats]\n", argv[0]);
        return 1;
    }
    const int n_gpus = argc > 6 ? atoi(argv[6]) : 1;
    const bool cpu = argc > 7 && std::string(argv[7]) == "cpu";
    const int repeats = argc > 8 ? atoi(argv[8]) : 1;
    const std::string curve(argv[1]);
    if (curve == "MNT4753") return run<mnt4753_libsnark>(argv[3], argv[4], argv[5], n_gpus, cpu, repeats);
    if (curve == "MNT6753") return run<mnt6753_libsnark>(argv[3], argv[4], argv[5], n_gpus, cpu, repeats);
    fprintf(stderr, "unknown curve %s\n", argv[1]);
    return 1;
}
