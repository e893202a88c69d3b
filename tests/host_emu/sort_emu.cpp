// CPU run of the first phase of an MSM (csrc/sort_kernels.cuh, the shipped kernels): k_count, the three scan kernels and
// k_scatter on a simulated thread block (block_sim.hpp), in the launch sequence of enqueue_msm (csrc/group_ops.cuh).
// Test-only.
#include "block_sim.hpp"

#include <cstring>

#undef __shared__
#define __shared__ static          // one array per kernel, seen by all the simulated threads of the block
static inline uint32_t atomicAdd(uint32_t *p, uint32_t v) { const uint32_t o = *p; *p = o + v; return o; }   // threads run one at a time

#include "../../gpu_groth16_prover_3x_b200/csrc/sort_kernels.cuh"

using namespace mnt753;

extern "C" {
// scalars: n x 24 words (plain integers, or the two stored halves of split scalars when glv); c window bits, Wd digits per
// scalar (Wh per half when glv), sets = bucket sets.  Out: count[K], offs[K + 1], entries[offs[K]] (K = sets * 2^(c-1)).
int emu_sort(uint32_t n, int c, int Wd, int sets, uint32_t tab_stride, int glv, int Wh, const uint32_t *scalars, const uint8_t *base_inf,
             uint32_t *count, uint32_t *offs, uint32_t *entries) {
    MsmArgs a;
    memset(&a, 0, sizeof a);
    a.n = n;
    a.c = c;
    a.Wd = Wd;
    a.W = sets;
    a.tab_stride = tab_stride;
    a.glv = glv;
    a.Wh = Wh;
    a.NB = 1u << (c - 1);
    a.K = (uint32_t)sets * a.NB;
    std::vector<uint32_t> sc(scalars, scalars + (size_t)n * NLIMB), cursor(a.K), bsum;
    a.scalars = sc.data();
    a.base_inf = base_inf;
    a.count = count;
    a.offs = offs;
    a.cursor = cursor.data();
    a.entries = entries;
    memset(count, 0, (size_t)a.K * 4);
    const unsigned nscan = (a.K + SCAN_B - 1) / SCAN_B;
    bsum.resize(nscan);
    a.i0 = 0;
    a.i1 = n;
    blocksim::launch((n + 255) / 256, 256, [&] { k_count(a); });
    blocksim::launch(nscan, SCAN_T, [&] { k_scan_local(a.count, a.offs, bsum.data(), a.K); });
    blocksim::launch(1, SCAN_T, [&] { k_scan_bsum(bsum.data(), nscan, a.offs + a.K); });
    blocksim::launch(nscan, SCAN_T, [&] { k_scan_add(a.offs, a.cursor, bsum.data(), a.K); });
    blocksim::launch((n + 255) / 256, 256, [&] { k_scatter(a); });
    return 0;
}
}
