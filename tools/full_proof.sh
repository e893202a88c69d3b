#!/bin/bash
# Acceptance run of the reference's README at the default size: generate_parameters (MNT4753 d = 2^20 - 1,
# MNT6753 d = 2^15 - 1), the reference CPU prover, the prover with MSMs + H on the engine, sha256 of the proofs;
# then (second baseline) the reference's own GPU prover, unmodified, compiled for sm_100a, where its 31x table
# is affordable (always for MNT6753; for MNT4753 only at the `fast` size).
# Usage: tools/full_proof.sh [fast]      (run on the GPU box; ~7 minutes at the default size)
set -u
REPO=$(cd "$(dirname "$0")/.." && pwd)
REF=$REPO/oracle/_ref
W=${TMPDIR:-/tmp}/g16_$$
mkdir -p "$W" && cd "$W"
now() { date +%s.%N; }
el() { python3 -c "print('%.2f' % ($(now) - $1))"; }
echo "host cores: $(nproc)"
t0=$(now); $REF/generate_parameters ${1:-} > gen.log 2>&1; echo "generate_parameters ${1:-default}: $(el $t0) s"
ls -la | grep -E "parameters|input"
for curve in MNT4753 MNT6753; do
  echo "=== $curve"
  t0=$(now); $REF/main $curve compute $curve-parameters $curve-input $curve-output-ref > main_$curve.log 2>&1
  echo "reference CPU prover (main $curve compute, $(nproc) threads): $(el $t0) s wall"
  echo "-- product CLI (b200msm_key_load_file + b200msm_prove, no libff):"
  $REPO/gpu_groth16_prover_3x_b200/b200_prove $curve compute $curve-parameters $curve-input $curve-output-cli 3
  echo "-- reference driver with MSMs + H on the engine (tests/integration/b200_prover.cpp):"
  $REF/b200_prover $curve compute $curve-parameters $curve-input $curve-output-b200 1 gpu-h 3
  $REF/b200_prover $curve compute $curve-parameters $curve-input $curve-output-b200-cpuh 1 cpu-h 2 | grep -E "compute_H|Total time"
  sha256sum $curve-output-ref $curve-output-cli $curve-output-b200 $curve-output-b200-cpuh
  if [ -x $REF/cuda_prover_piecewise ] && { [ $curve = MNT6753 ] || [ "${1:-}" = fast ]; }; then
    t0=$(now); $REF/main $curve preprocess $curve-parameters > pre_$curve.log 2>&1; echo "reference preprocess (31x table): $(el $t0) s, $(du -h ${curve}_preprocessed | cut -f1)"
    t0=$(now); $REF/cuda_prover_piecewise $curve compute $curve-parameters $curve-input $curve-output-refgpu > refgpu_$curve.log 2>&1
    echo "reference GPU prover (cuda_prover_piecewise, sm_100a build): $(el $t0) s wall"
    grep -E "gpu e2e|Total time from input|cpu 1|load preprocessing" refgpu_$curve.log
    sha256sum $curve-output-refgpu
  fi
done
rm -rf "$W"
