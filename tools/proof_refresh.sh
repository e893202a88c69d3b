#!/bin/bash
# Proof latency at the default size on the current build, without the 2-minute reference CPU prover (tools/full_proof.sh
# does the full acceptance run): generate_parameters, then the product CLI and the reference driver on the engine.
set -u
REPO=$(cd "$(dirname "$0")/.." && pwd)
REF=$REPO/oracle/_ref
W=${TMPDIR:-/tmp}/g16r_$$
mkdir -p "$W" && cd "$W"
t0=$(date +%s); $REF/generate_parameters > gen.log 2>&1; echo "generate_parameters default: $(( $(date +%s) - t0 )) s"
for curve in MNT4753 MNT6753; do
  echo "=== $curve"
  echo "-- product CLI (b200msm_key_load_file + b200msm_prove, no libff):"
  $REPO/gpu_groth16_prover_3x_b200/b200_prove $curve compute $curve-parameters $curve-input $curve-output-cli 3
  echo "-- reference driver with MSMs + H on the engine (tests/integration/b200_prover.cpp):"
  $REF/b200_prover $curve compute $curve-parameters $curve-input $curve-output-b200 1 gpu-h ${1:-3} | grep -E "upload|compute_H|gpu e2e|Total time"
  sha256sum $curve-output-cli $curve-output-b200
done
rm -rf "$W"
