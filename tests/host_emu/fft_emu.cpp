// CPU run of the H-polynomial kernels (csrc/fft_kernels.cuh, the shipped code: domain constants, power tables, the tiled
// and the one-stage-per-launch NTTs, the pointwise steps) on a simulated thread block (block_sim.hpp).  The launch
// sequence is restated from csrc/fft.cu (fft_prepare, compute_h_stage x 3, compute_h_finish).  Test-only.
#include "block_sim.hpp"

#include <cstring>

#include "../../gpu_groth16_prover_3x_b200/csrc/fq.cuh"
namespace mnt753 { uint4 sm[6 << 10]; }      // the kernels' dynamic shared memory: 96 KB, a tile of 2^10 elements
#include "../../gpu_groth16_prover_3x_b200/csrc/fft_kernels.cuh"

using namespace mnt753;

namespace {
typedef std::vector<uint32_t> Vec;

template <class M, bool DIF>
void ntt_tiled(uint32_t *x, int logm, const uint32_t *tw, int tile_bits) {
    const int ngroups = (logm + tile_bits - 1) / tile_bits;
    int bits[32], start[32];
    for (int g = 0, s = 0; g < ngroups; ++g) { bits[g] = logm / ngroups + (g < logm % ngroups ? 1 : 0); start[g] = s; s += bits[g]; }
    for (int k = 0; k < ngroups; ++k) {
        const int g = DIF ? ngroups - 1 - k : k;
        blocksim::launch(1u << (logm - bits[g]), NTT_TILE_THREADS, [&] { k_ntt_tile<M, DIF>(x, tw, logm, start[g], bits[g]); });
    }
}
template <class M>
void ifft_dif(uint32_t *x, uint32_t m, int logm, const uint32_t *twi, int tile_bits) {
    if (tile_bits) { ntt_tiled<M, true>(x, logm, twi, tile_bits); return; }
    for (uint32_t len = m; len >= 2; len >>= 1) blocksim::launch((m / 2 + 127) / 128, 128, [&] { k_ntt_dif<M>(x, twi, m, len); });
}
template <class M>
void fft_dit(uint32_t *x, uint32_t m, int logm, const uint32_t *tw, int tile_bits) {
    if (tile_bits) { ntt_tiled<M, false>(x, logm, tw, tile_bits); return; }
    for (uint32_t len = 2; len <= m && len; len <<= 1) blocksim::launch((m / 2 + 127) / 128, 128, [&] { k_ntt_dit<M>(x, tw, m, len); });
}

// tile_bits = 0: the one-launch-per-stage kernels; else the tiled kernel with groups of at most tile_bits index bits
template <class M>
int compute_h(int logm, int tile_bits, const uint64_t *ca, const uint64_t *cb, const uint64_t *cc, uint64_t *out) {
    const size_t m = size_t(1) << logm;
    if (logm > M::TWO_ADICITY || tile_bits > NTT_TILE_BITS) return 1;
    Vec consts(FC_COUNT * NLIMB), tw(std::max<size_t>(m / 2, 1) * NLIMB), twi(tw.size()), cg(m * NLIMB), cgi(m * NLIMB), res((m + 1) * NLIMB);
    Vec a(m * NLIMB), b(m * NLIMB), c(m * NLIMB);
    const uint32_t *C = consts.data();
    // fft_prepare
    blocksim::launch(1, 32, [&] { k_fft_consts<M>(consts.data(), logm); });
    const unsigned gh = (unsigned)((m / 2 + 127) / 128), gm = (unsigned)((m + 127) / 128), gm1 = (unsigned)((m + 1 + 127) / 128);
    if (m >= 2) {
        blocksim::launch(gh, 128, [&] { k_powers<M>(tw.data(), (uint32_t)(m / 2), C + FC_OMEGA * NLIMB, C + FC_ONE * NLIMB, 0); });
        blocksim::launch(gh, 128, [&] { k_powers<M>(twi.data(), (uint32_t)(m / 2), C + FC_OMEGA_INV * NLIMB, C + FC_ONE * NLIMB, 0); });
    }
    blocksim::launch(gm, 128, [&] { k_powers<M>(cg.data(), (uint32_t)m, C + FC_G * NLIMB, C + FC_M_INV * NLIMB, logm); });
    blocksim::launch(gm, 128, [&] { k_powers<M>(cgi.data(), (uint32_t)m, C + FC_G_INV * NLIMB, C + FC_M_INV * NLIMB, logm); });
    // compute_h_stage x 3
    memcpy(a.data(), ca, m * 96);
    memcpy(b.data(), cb, m * 96);
    memcpy(c.data(), cc, m * 96);
    for (uint32_t *x : {a.data(), b.data(), c.data()}) {
        ifft_dif<M>(x, (uint32_t)m, logm, twi.data(), tile_bits);
        blocksim::launch(gm, 128, [&] { k_pointwise_mul<M>(x, cg.data(), (uint32_t)m); });
        fft_dit<M>(x, (uint32_t)m, logm, tw.data(), tile_bits);
    }
    // compute_h_finish
    blocksim::launch(gm, 128, [&] { k_h_pointwise<M>(a.data(), b.data(), c.data(), C + FC_Z_INV * NLIMB, (uint32_t)m); });
    ifft_dif<M>(a.data(), (uint32_t)m, logm, twi.data(), tile_bits);
    blocksim::launch(gm1, 128, [&] { k_h_final<M>(res.data(), a.data(), cgi.data(), (uint32_t)m, logm); });
    memcpy(out, res.data(), (m + 1) * 96);
    return 0;
}
}  // namespace

extern "C" {
// coefficients_for_H (d + 2 = 2^logm + 1 elements) over Fr of `curve` from ca, cb, cc (2^logm elements each)
int emu_compute_h(int curve, int logm, int tile_bits, const uint64_t *ca, const uint64_t *cb, const uint64_t *cc, uint64_t *out) {
    return curve == 0 ? compute_h<ModB>(logm, tile_bits, ca, cb, cc, out) : compute_h<ModA>(logm, tile_bits, ca, cb, cc, out);
}
}
