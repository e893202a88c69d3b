// b200_prove -- command-line twin of the reference's GPU prover for its `compute` mode
// (cuda_prover_piecewise.cu:232-263):   b200_prove <MNT4753|MNT6753> compute <params> <input> <output> [repeats [gpus]]
// No preprocessing file and no libff: the parameter file is loaded into HBM once (b200msm_key_load_sharded_file), every
// proof is one b200msm_prove_sharded_file call, the output file holds the same bytes as the reference provers write.
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
        fprintf(stderr, "usage: %s <MNT4753|MNT6753> compute <params> <input> <output> [repeats [gpus]]\n", argv[0]);
        return 1;
    }
    const std::string curve(argv[1]);
    if (curve != "MNT4753" && curve != "MNT6753") { fprintf(stderr, "unknown curve %s\n", argv[1]); return 1; }
    const int repeats = argc > 6 ? atoi(argv[6]) : 1;
    const int gpus = argc > 7 ? atoi(argv[7]) : 1;
    if (gpus < 1 || gpus > 16) { fprintf(stderr, "gpus: 1..16\n"); return 1; }
    auto t_all = std::chrono::high_resolution_clock::now();
    const int curve_id = curve == "MNT4753" ? B200MSM_MNT4753 : B200MSM_MNT6753;
    std::vector<b200msm_ctx *> ctxs(gpus, nullptr);
    std::vector<b200msm_key *> keys(gpus, nullptr);
    for (int g = 0; g < gpus; ++g)
        if (b200msm_create(curve_id, g, &ctxs[g])) { fprintf(stderr, "no usable sm_100 device %d\n", g); return 3; }
    b200msm_ctx *ctx = ctxs[0];
    auto t = std::chrono::high_resolution_clock::now();
    // every GPU takes its point range of every query out of one image of the parameter file, loaded by its own host thread
    if (b200msm_key_load_sharded_file(ctxs.data(), gpus, argv[3], keys.data())) { fprintf(stderr, "%s\n", b200msm_last_error(ctx)); return 2; }
    b200msm_key *key = keys[0];
    uint64_t info[2];
    b200msm_key_info(key, info);
    printf("d = %llu, m = %llu, gpus = %d; key load + window tables: %.1f ms\n", (unsigned long long)info[0], (unsigned long long)info[1], gpus, ms_since(t));
    std::vector<uint8_t> proof(b200msm_proof_bytes(ctx));
    const size_t input_bytes = b200msm_input_bytes(key);
    char *input = static_cast<char *>(b200msm_pinned_alloc(input_bytes));
    if (!input) { fprintf(stderr, "cannot allocate %zu bytes of pinned memory\n", input_bytes); return 3; }
    for (int rep = 0; rep < repeats; ++rep) {
        t = std::chrono::high_resolution_clock::now();
        if (b200msm_prove_sharded_file(ctxs.data(), keys.data(), gpus, argv[4], input, proof.data())) { fprintf(stderr, "%s\n", b200msm_last_error(ctx)); return 4; }
        FILE *f;
        f = fopen(argv[5], "wb");
        if (!f || fwrite(proof.data(), 1, proof.size(), f) != proof.size()) { fprintf(stderr, "cannot write %s\n", argv[5]); return 2; }
        fclose(f);
        printf("Total time from input to output: %.1f ms\n", ms_since(t));
    }
    b200msm_pinned_free(input);
    for (int g = 0; g < gpus; ++g) { b200msm_key_free(ctxs[g], keys[g]); b200msm_destroy(ctxs[g]); }
    printf("Total runtime (incl. key load): %.1f ms\n", ms_since(t_all));
    return 0;
}
