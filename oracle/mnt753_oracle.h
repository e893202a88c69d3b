/* TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C) of the reference's MSM path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product (gpu_groth16_prover_3x_b200/) never does.
 *
 * Parity status: PINNED.  Every function here is checked against the reference's own libff
 * (oracle/_ref/libref.so, built from /root/reference by oracle/Makefile) by
 * tools/gen_golden.py, whose outputs are committed under tests/golden/ and replayed by
 * tests/test_oracle.py on machines where the reference is absent.
 *
 * Wire formats are the reference's (libsnark/serialization.hpp:24-121):
 *   Fq/Fr   12 x u64 little-endian limbs, Montgomery form with R = 2^768, canonical
 *   Fqe     DEG consecutive Fq (c0, c1[, c2]); DEG = 2 (MNT4753 G2), 3 (MNT6753 G2)
 *   affine  x || y, infinity = all-zero, recognised on input by y == 0
 *   jacobian X || Y || Z (what the reference's GPU kernels emit, multiexp/curves.cu:104-114)
 * curve: 0 = MNT4753, 1 = MNT6753.  group: 1 = G1, 2 = G2.
 */
#ifndef MNT753_ORACLE_H
#define MNT753_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* field: 0 = Fq, 1 = Fqe.  op: 0 mul, 1 add, 2 sub, 3 sqr, 4 inv (0 -> 0), 5 neg. */
int orc_field_op(int curve, int field, int op, size_t n, const uint64_t *a, const uint64_t *b, uint64_t *out);
int orc_fr_from_mont(int curve, size_t n, const uint64_t *in, uint64_t *out);
int orc_fr_to_mont(int curve, size_t n, const uint64_t *in, uint64_t *out);

/* op: 0 a+b, 1 dbl(a), 2 a.mixed_add(b), 3 k*a (k = Montgomery Fr), 4 -a, 5 a.add(b).  Affine in/out. */
int orc_point_op(int curve, int group, int op, const uint64_t *a, const uint64_t *b, const uint64_t *k, uint64_t *out);

/* bases[i] = P0 + i*Q (P0, Q affine), written in affine wire format. */
int orc_gen_bases(int curve, int group, size_t n, const uint64_t *p0, const uint64_t *q, uint64_t *out);

/* method: 0 naive, 1 BDLO12 (multiexp.tcc:165-282).  prefilter: multiexp.tcc:443-496.
 * chunks <= 0 -> omp_get_max_threads() (main.cpp:150-170).  Returns seconds; affine result in out. */
double orc_msm(int curve, int group, size_t n, const uint64_t *bases, const uint64_t *scalars, uint64_t *out,
               int method, int chunks, int prefilter);

/* (sum s_i) * P0 + (sum i*s_i) * Q : closed form of an MSM over bases P0 + i*Q. */
int orc_msm_closed_form(int curve, int group, size_t n, const uint64_t *p0, const uint64_t *q,
                        const uint64_t *scalars, uint64_t *out);

/* read_pt_ECp/ECpe (prover_reference_functions.cpp:106-115) then write_g1/g2. */
int orc_jacobian_to_affine(int curve, int group, const uint64_t *xyz, uint64_t *out);

/* sum of n Jacobian partials (multi-GPU host fold), result affine. */
int orc_fold_jacobian(int curve, int group, size_t n, const uint64_t *xyz, uint64_t *out);

/* coefficients_for_H = compute_H(d, ca, cb, cc) (cuda_prover_piecewise.cu:14-49) over Fr: ca, cb, cc hold
 * d + 1 Montgomery elements each (d + 1 a power of two), out receives d + 2 (the last one zero). */
int orc_compute_h(int curve, size_t d, const uint64_t *ca, const uint64_t *cb, const uint64_t *cc, uint64_t *out);

int orc_num_threads(void);
void orc_set_num_threads(int t);

#ifdef __cplusplus
}
#endif
#endif
