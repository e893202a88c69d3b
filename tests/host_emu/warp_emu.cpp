// CPU harness for the warp-cooperative device code: csrc/fq_inv_coop.cuh run on a simulated warp (warp_sim.hpp),
// limb i on lane i exactly as on the device.  Compiled by tests/conftest.py with g++ -DMNT753_HOST_EMU; test-only.
#include "warp_sim.hpp"

#include "../../gpu_groth16_prover_3x_b200/csrc/fq_inv_coop.cuh"

using namespace mnt753;

namespace {
template <class M>
int inv_coop(size_t n, const uint64_t *a, uint64_t *out) {
    int failed = 0;
    for (size_t i = 0; i < n; ++i) {
        uint32_t in[NLIMB], res[warpsim::LANES];
        bool ok[warpsim::LANES];
        memcpy(in, a + 12 * i, 96);
        warpsim::run_warp([&](int lane) {
            uint32_t x = lane < NLIMB ? in[lane] : 0xdeadbeefu;      // lanes >= 24 are ignored by contract
            ok[lane] = fq_inv_coop<M>(x);
            res[lane] = x;
        });
        for (int l = 1; l < warpsim::LANES; ++l)
            if (ok[l] != ok[0]) return -1;                            // the return value must be warp-uniform
        if (!ok[0]) ++failed;
        memcpy(out + 12 * i, res, 96);
    }
    return failed;
}
}  // namespace

extern "C" {
// Montgomery-form inverse of n elements by the warp-cooperative inversion; returns the number of inputs on which it
// reported failure (the caller's cue to fall back to fq_inv), -1 if the lanes disagreed on that.
int emu_fq_inv_coop(int modulus, size_t n, const uint64_t *a, uint64_t *out) {
    return modulus == 0 ? inv_coop<ModA>(n, a, out) : inv_coop<ModB>(n, a, out);
}
}
