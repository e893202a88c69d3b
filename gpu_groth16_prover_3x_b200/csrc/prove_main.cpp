// b200_prove -- command-line twin of the reference's GPU prover for its `compute` mode
// (cuda_prover_piecewise.cu:232-263):   b200_prove <MNT4753|MNT6753> compute <params> <input> <output> [repeats]
// No preprocessing file and no libff: the parameter file is loaded into HBM once (b200msm_key_load_file), every
// proof is one b200msm_prove call, the output file holds the same bytes as the reference provers write.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200_msm.h"

static double ms_since(std::chrono::high_resolution_clock::time_point t) {
    return std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t).count();
}

int main(int argc, char **argv) {
    setbuf(stdout, NULL);
    if (argc < 6 || std::string(argv[2]) != "compute") {
        fprintf(stderr, "usage: %s <MNT4753|MNT6753> compute <params> <input> <output> [repeats]\n", argv[0]);
        return 1;
    }
    const std::string curve(argv[1]);
    if (curve != "MNT4753" && curve != "MNT6753") { fprintf(stderr, "unknown curve %s\n", argv[1]); return 1; }
    const int repeats = argc > 6 ? atoi(argv[6]) : 1;
    auto t_all = std::chrono::high_resolution_clock::now();
    b200msm_ctx *ctx = nullptr;
    if (b200msm_create(curve == "MNT4753" ? B200MSM_MNT4753 : B200MSM_MNT6753, 0, &ctx)) { fprintf(stderr, "no usable sm_100 device\n"); return 3; }
    b200msm_key *key = nullptr;
    auto t = std::chrono::high_resolution_clock::now();
    if (b200msm_key_load_file(ctx, argv[3], &key)) { fprintf(stderr, "%s\n", b200msm_last_error(ctx)); return 2; }
    uint64_t info[2];
    b200msm_key_info(key, info);
    printf("d = %llu, m = %llu; key load + window tables: %.1f ms\n", (unsigned long long)info[0], (unsigned long long)info[1], ms_since(t));
    std::vector<uint8_t> proof(b200msm_proof_bytes(ctx));
    const size_t input_bytes = b200msm_input_bytes(key);
    char *input = static_cast<char *>(b200msm_pinned_alloc(input_bytes));
    if (!input) { fprintf(stderr, "cannot allocate %zu bytes of pinned memory\n", input_bytes); return 3; }
    for (int rep = 0; rep < repeats; ++rep) {
        t = std::chrono::high_resolution_clock::now();
        if (b200msm_prove_file(ctx, key, argv[4], input, proof.data())) { fprintf(stderr, "%s\n", b200msm_last_error(ctx)); return 4; }
        FILE *f;
        f = fopen(argv[5], "wb");
        if (!f || fwrite(proof.data(), 1, proof.size(), f) != proof.size()) { fprintf(stderr, "cannot write %s\n", argv[5]); return 2; }
        fclose(f);
        printf("Total time from input to output: %.1f ms\n", ms_since(t));
    }
    b200msm_pinned_free(input);
    b200msm_key_free(ctx, key);
    b200msm_destroy(ctx);
    printf("Total runtime (incl. key load): %.1f ms\n", ms_since(t_all));
    return 0;
}
