/* b200_msm.h -- C ABI of the B200-native MNT4753/MNT6753 multi-scalar-multiplication engine.
 *
 * This is the drop-in boundary for the MSM path of vezenovm/gpu-groth16-prover-3x.  Every entry
 * point takes plain pointers and sizes in the reference's own wire format
 * (libsnark/serialization.hpp:24-121, multiexp/reduce.cu:131-152):
 *
 *   scalar   12 x u64 little-endian limbs, Montgomery form of Fr (R = 2^768)
 *   base     affine x || y, each coordinate DEG x 12 limbs Montgomery; infinity <=> y == 0
 *            (DEG = 1 for G1, 2 for MNT4753 G2, 3 for MNT6753 G2)
 *   result   Jacobian X || Y || Z (3 * DEG * 12 limbs), infinity <=> Z == 0, reported as (1,1,0);
 *            it is what B::read_pt_ECp / read_pt_ECpe (prover_reference_functions.cpp:795-817)
 *            already consume.
 *
 * There is no CPU fallback behind any of these symbols: without a usable CUDA device they fail
 * with B200MSM_ERR_CUDA and never produce a point.
 */
#ifndef B200_MSM_H
#define B200_MSM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200MSM_MNT4753 0
#define B200MSM_MNT6753 1
#define B200MSM_G1 1
#define B200MSM_G2 2

#define B200MSM_OK 0
#define B200MSM_ERR_ARG 1
#define B200MSM_ERR_CUDA 2
#define B200MSM_ERR_OOM 3

typedef struct b200msm_ctx b200msm_ctx;

/* One context per (curve, GPU).  Replaces the implicit global state of cuda_prover_piecewise.cu
 * (cudaMallocManaged buffers + streams created inside ec_reduce_straus, reduce.cu:135,198-209). */
int b200msm_create(int curve, int device, b200msm_ctx **out);
void b200msm_destroy(b200msm_ctx *ctx);
/* Human-readable description of the last failure on this context (never NULL). */
const char *b200msm_last_error(const b200msm_ctx *ctx);

/* Upload a base set (A, B1, L, H queries: group G1; B2 query: group G2) and keep it resident in
 * HBM.  Replaces load_points_affine<EC>(31 * n, preprocessed_file) (reduce.cu:254-271,
 * cuda_prover_piecewise.cu:132-139) AND the reference's `main preprocess` step (main.cpp:311-339):
 * instead of reading a 31x multiples table from a 25 GB file, the upload builds "window tables"
 * 2^(c*G*t) * P_i (t < NT) on the device, inside the per-set byte budget, so that an MSM needs only
 * G = ceil(windows / NT) bucket sets (one, with a full set of tables).  Sets of fewer than 256 points
 * get no tables.  `affine` may be a host or device pointer.  Returns a slot id >= 0 in *slot. */
int b200msm_bases_upload(b200msm_ctx *ctx, int group, const uint64_t *affine, size_t n, int *slot);
int b200msm_bases_free(b200msm_ctx *ctx, int slot);
/* Byte budget for the window tables of each base set uploaded afterwards (default 32 GiB; 0 = never
 * build tables, every MSM then runs the plain one-bucket-set-per-window Pippenger).
 * G2 base sets with at least two tables use the twist Frobenius psi(P) = [q mod r] P (README.md:73 of the
 * reference; libff G2::mul_by_q, mnt4753_g2.cpp:364-368): scalars are split k = k0 + k1 (q mod r) on the
 * device, the tables come in two halves of equal size (an EVEN number of tables), and the upper half is psi of the
 * lower half -- two multiplications by Fq constants per point instead of doublings and a normalisation. */
int b200msm_set_table_budget(b200msm_ctx *ctx, size_t max_bytes_per_set);
/* info[0] = points, [1] = window bits the tables were built for (0: none), [2] = tables NT,
 * [3] = bucket sets G, [4] = bytes resident, [5] = table build time in microseconds. */
int b200msm_bases_info(b200msm_ctx *ctx, int slot, uint64_t info[6]);

/* result = sum_{i<n} scalars[i] * bases[offset + i] over the resident base set `slot`.
 * Replaces ec_reduce_straus<EC,C,R>(strm, out, multiples, scalars, N) (reduce.cu:131-152) and,
 * through the host adaptor in INTEGRATION.md, B::multiexp_G1 / multiexp_G2
 * (prover_reference_functions.cpp:350-368, 690-708).
 * `scalars_mont` may be host (pageable or pinned) or device memory; `out_xyz` is host memory.
 * The _async form enqueues on internal stream `lane` (0..4, the reference used one CUDA stream per
 * MSM as a future: cuda_prover_piecewise.cu:162-167,187-193) and returns immediately; the result
 * is valid after b200msm_wait(ctx, lane). */
int b200msm_msm(b200msm_ctx *ctx, int slot, size_t offset, const uint64_t *scalars_mont, size_t n, uint64_t *out_xyz);
int b200msm_msm_async(b200msm_ctx *ctx, int lane, int slot, size_t offset, const uint64_t *scalars_mont, size_t n,
                      uint64_t *out_xyz);
int b200msm_wait(b200msm_ctx *ctx, int lane);

/* Literal counterpart of ec_reduce_straus: bases are passed with the call (host or device
 * pointer, n affine points -- NOT the 31n-point multiples table) and uploaded inside the call. */
int b200msm_ec_reduce(b200msm_ctx *ctx, int group, const uint64_t *bases_affine, const uint64_t *scalars_mont,
                      size_t n, uint64_t *out_xyz);

/* Synthetic base set of SURVEY.md 8(d), generated in HBM: bases[i] = P0 + i*Q with P0 = k_p0 * G and
 * Q = k_q * G (G = G1_one / G2_one of the curve; scalars in Montgomery form).  Its MSM has the closed
 * form (sum s_i) P0 + (sum i s_i) Q.  b200msm_bases_download copies points of a resident set to
 * host (or device) memory in the affine wire format. */
int b200msm_bases_synthetic(b200msm_ctx *ctx, int group, size_t n, const uint64_t *k_p0_mont, const uint64_t *k_q_mont,
                            int *slot);
int b200msm_bases_download(b200msm_ctx *ctx, int slot, size_t offset, size_t n, uint64_t *out_affine);

/* Affine normalisation of n Jacobian points on the device: (X/Z^2, Y/Z^3), infinity -> all-zero
 * (the writer convention of libsnark/serialization.hpp:43-67).  The reference does this on the host
 * with libff after read_pt_* (cuda_prover_piecewise.cu:187-204); its kernels have no inversion. */
int b200msm_to_affine(b200msm_ctx *ctx, int group, size_t n, const uint64_t *xyz, uint64_t *out_affine);

/* Multi-GPU: an MSM shards by point range (one context per GPU, each returns one partial Jacobian
 * point); this folds n partials into one point on this context's GPU.  The reference has no
 * multi-GPU path; inside the reference the same fold is G - 1 libff additions after read_pt_*. */
int b200msm_fold(b200msm_ctx *ctx, int group, const uint64_t *partials_xyz, size_t n, uint64_t *out_xyz);
/* THE point-range sharding rule, used by b200msm_key_load_shard and to be used by every caller that
 * shards on its own: shard g of G of an n-point query owns [n g / G, n (g + 1) / G) (integer division).  Pure host
 * arithmetic, no context.  Returns B200MSM_ERR_ARG for nshards < 1 or shard outside [0, nshards). */
int b200msm_shard_range(size_t n, int shard, int nshards, size_t *offset, size_t *length);

/* The H-polynomial of the prover on the device: coefficients_for_H = compute_H(d, ca, cb, cc)
 * (cuda_prover_piecewise.cu:14-49), i.e. the three inverse FFTs, three coset FFTs, the pointwise
 * (ca * cb - cc) / Z and the inverse coset FFT that the reference runs on the CPU through libfqfft
 * (basic_radix2_domain.tcc:63-126), over the scalar field Fr of the context's curve.  ca, cb, cc: d + 1
 * Montgomery-form Fr elements each (host or device); d + 1 must be a power of two <= 2^s (s = 30 for
 * MNT4753, 15 for MNT6753).  Result: d + 2 elements (the last one zero, like vector_Fr_zeros(m + 1)),
 * copied to out_host when it is not NULL and left in device memory owned by the context -- *out_dev, valid
 * until the next call -- so that the H-query MSM can take its scalars without a round trip over PCIe.
 * Synchronous.  The domain tables (twiddles, coset powers: 3 (d+1) elements) are cached per size. */
int b200msm_compute_h(b200msm_ctx *ctx, size_t d, const uint64_t *ca, const uint64_t *cb, const uint64_t *cc,
                      uint64_t *out_host, const uint64_t **out_dev);
/* ms[0] = device time of the last b200msm_compute_h (copies included), ms[1] = one-off table build. */
int b200msm_compute_h_timings(b200msm_ctx *ctx, float ms[2]);
/* Free the FFT tables and work vectors (also done by b200msm_destroy). */
void b200msm_fft_release(b200msm_ctx *ctx);

/* k * P for one affine point and one Montgomery-form scalar of Fr -> Jacobian (the r * Bt1 of the proof
 * assembly, cuda_prover_piecewise.cu:198, done by libff on the host in the reference). */
int b200msm_scalar_mul(b200msm_ctx *ctx, int group, const uint64_t *affine, const uint64_t *k_mont, uint64_t *out_xyz);

/* ---- the whole `compute` step --------------------------------------------------
 * A proving key resident in HBM and one call per proof: replaces run_prover (cuda_prover_piecewise.cu:96-230)
 * including B::read_params / load_points_affine (no preprocessed file), the five MSMs, compute_H, the assembly
 * C = Ht + Lt + r * Bt1 and groth16_output_write.  `params_image` / `input_image` are the bytes of the reference's
 * <curve>-parameters and <curve>-input files (generate_parameters.cpp:59-108, main.cpp:35-85; host memory);
 * `proof` receives b200msm_proof_bytes() bytes, identical to the file the reference's provers write
 * (A || B || C affine: 768 bytes for MNT4753, 960 for MNT6753).  Single GPU; uses lanes 0-4 of the context
 * (A, B1, B2, L, H). */
typedef struct b200msm_key b200msm_key;
int b200msm_key_load(b200msm_ctx *ctx, const void *params_image, size_t bytes, b200msm_key **key);
int b200msm_key_load_file(b200msm_ctx *ctx, const char *path, b200msm_key **key);
int b200msm_key_info(const b200msm_key *key, uint64_t info[2]); /* d, m */
void b200msm_key_free(b200msm_ctx *ctx, b200msm_key *key);
size_t b200msm_proof_bytes(const b200msm_ctx *ctx);
size_t b200msm_input_bytes(const b200msm_key *key);
int b200msm_prove(b200msm_ctx *ctx, const b200msm_key *key, const void *input_image, size_t bytes, uint8_t *proof);
/* The same from the <curve>-input FILE (what run_prover does with load_scalars / B::read_input,
 * cuda_prover_piecewise.cu:151-155): r and the witness are read first, the four witness MSMs start, and the
 * coefficient vectors of the H polynomial (three quarters of the file) are read while the GPU works.
 * `buffer`: host scratch of b200msm_input_bytes() bytes (b200msm_pinned_alloc for full-rate uploads). */
int b200msm_prove_file(b200msm_ctx *ctx, const b200msm_key *key, const char *input_path, void *buffer, uint8_t *proof);
/* The same over several GPUs of one box (the reference has no multi-GPU path): every query is split by
 * point range, shard g of n (a context on its own GPU and the key shard loaded into it) runs its five MSMs on its
 * slice of the witness, the H polynomial is computed on shard 0's GPU and its coefficients reach the other shards by
 * peer copy, the 5 n partial points are folded on shard 0's GPU.  No collective; ctxs[g] / keys[g] must be shard g of n
 * of the same parameter image.  Contexts may share a device (tests). */
int b200msm_key_load_shard(b200msm_ctx *ctx, const void *params_image, size_t bytes, int shard, int nshards, b200msm_key **key);
/* All shards from the <curve>-parameters FILE: read once, shard g loaded into ctxs[g] by its own host thread (uploads and
 * window-table builds of the GPUs run side by side).  keys: nshards entries; on failure none is left allocated. */
int b200msm_key_load_sharded_file(b200msm_ctx *const *ctxs, int nshards, const char *path, b200msm_key **keys);
int b200msm_prove_sharded(b200msm_ctx *const *ctxs, b200msm_key *const *keys, int nshards, const void *input_image, size_t bytes,
                          uint8_t *proof);
int b200msm_prove_sharded_file(b200msm_ctx *const *ctxs, b200msm_key *const *keys, int nshards, const char *input_path, void *buffer,
                               uint8_t *proof);
/* Page-locked host memory for witness / scalar buffers (H2D at full PCIe rate, truly asynchronous uploads);
 * plain malloc'ed memory works everywhere too, only slower.  NULL on failure. */
void *b200msm_pinned_alloc(size_t bytes);
void b200msm_pinned_free(void *p);

/* Enqueue lane `lane` on a caller-owned CUDA stream (a cudaStream_t passed as void*; NULL restores the
 * internal stream).  The reference hands a cudaStream_t& back to its caller for the same purpose
 * (reduce.cu:131-135): ordering the MSM against the caller's own work and timing it with events. */
int b200msm_set_stream(b200msm_ctx *ctx, int lane, void *cuda_stream);

/* MSMs enqueued on `lane` afterwards occupy at most `sms` SMs (0, or anything outside (0, SM count): all of them).
 * The accumulation and the first reduction rounds of an MSM are persistent kernels of one block per SM; with the
 * default every MSM takes the whole GPU and MSMs on different lanes run one after the other.  Small MSMs are bound by
 * the LATENCY of their rounds, not by throughput: giving the lanes disjoint parts of the GPU lets them run side by
 * side (b200msm_prove does this on its own for the four witness MSMs of a small proof, B200MSM_LANE_SPLIT=0 turns
 * it off).  The result does not depend on the setting.  The reference has no counterpart: its four kernels shared
 * the GPU through the hardware scheduler (cuda_prover_piecewise.cu:162-167). */
int b200msm_set_lane_sms(b200msm_ctx *ctx, int lane, int sms);

/* Tuning / introspection. */
/* Window width for MSMs and for the tables of base sets uploaded afterwards; 0 = automatic.  An MSM
 * whose forced width differs from the one its base set's tables were built for ignores the tables. */
int b200msm_set_window_bits(b200msm_ctx *ctx, int c);
/* Device time of the phases of the most recent completed MSM on `lane`, milliseconds, measured with
 * CUDA events on the launching stream: [0] total, [1] H2D scalars, [2] recode+sort,
 * [3] bucket accumulation (k_batch_add + k_ba_fixup), [4] bucket reduction + window combine, [5] D2H result.
 * info[0] = window bits c, [1] = signed digits (windows) per scalar, [2] = sorted entries,
 * [3] = shares the sorted list was cut into (teams of the accumulation kernel), [4] = total kernel launches of the
 * MSM, [5] = bucket sets G, [6] = window tables used NT, [7] = 0. */
int b200msm_last_timings(b200msm_ctx *ctx, int lane, float ms[6], uint64_t info[8]);

/* Batched-affine rounds of the most recent completed MSM on `lane`: info[0] = rounds executed (the largest number
 * over the shares: every team runs its own buckets through their rounds), [1] = shares, [2] = largest bucket (or
 * piece of a split bucket) after the sort, [3] = affine additions performed; pairs_per_round (may be NULL) receives
 * the additions of each round summed over the shares. */
int b200msm_last_rounds(b200msm_ctx *ctx, int lane, uint64_t info[4], uint32_t *pairs_per_round, size_t max_rounds);

/* Synthetic microbenchmarks used by bench.py for the roofline denominator: runs `iters`
 * dependent-chain iterations of the named instruction mix on every SM and returns the achieved
 * rate in 10^9 operations per second (a 32x32->64 multiply-accumulate counts as one operation).
 * kind: 0 = IMAD.WIDE.U32 (the 32x32->64 product the MSM is made of; half the rate of a 32-bit IMAD
 *           on B200: 32 per clock per SM), 1 = IMAD (mad.lo, 64 per clock per SM),
 *       2 = the engine's own Fq Montgomery multiplication (32-bit CIOS on IMAD.WIDE.U32.X carry
 *           chains; returns 10^9 modmul/s),
 *       (3 was the reduced-radix experiment, now tools/experiments/fq_experiments.cuh; rejected);
 *       4, 5, 6 = latency of one Fq inversion by a single thread, in MICROSECONDS (not a rate): the engine's
 *           fq_inv, the plain binary gcd, the approximation-based fast path alone (fails if it ever needs
 *           the fallback);
 *       7..10 (G1) / 11..14 (G2) = the slab multiplier Team::mul (operands in shared memory, as used by
 *           every curve operation) with 1..4 resident blocks per SM; returns 10^9 tower products/s;
 *       15, 16, 17 = kind 2 (register operands) with only 4, 8, 12 warps per SM. */
int b200msm_microbench(b200msm_ctx *ctx, int kind, int iters, double *gops);

/* Self-test hooks (used by tests/ only): the device field / point layer applied elementwise.
 * field op: 0 mul, 1 add, 2 sub, 3 sqr, 5 neg, 6 mul_by_curve_a, 7 in-place mul, 8 dbl; elements are
 * Fq for G1 and Fqe for G2 (DEG x 12 limbs).  point op: 0 acc += affine q (flags bit0: negate q,
 * bit1: acc is infinity), 1 acc += jacobian q, 2 acc = 2*acc; acc / out are Jacobian X||Y||Z. */
int b200msm_selftest_field(b200msm_ctx *ctx, int group, int op, size_t n, const uint64_t *a, const uint64_t *b,
                           uint64_t *out);
int b200msm_selftest_point(b200msm_ctx *ctx, int group, int op, size_t n, const uint64_t *acc, const uint64_t *q,
                           const uint32_t *flags, uint64_t *out);

#ifdef __cplusplus
}
#endif
#endif
