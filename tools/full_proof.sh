#!/bin/bash
# Default-size acceptance run (README procedure of the reference): generate_parameters (MNT4753 d = 2^20 - 1,
# MNT6753 d = 2^15 - 1), the reference CPU prover, the prover with MSMs + H on the engine; sha256 of the proofs.
# Usage: tools/full_proof.sh [fast]     (run on the GPU box; takes ~15 minutes at the default size)
set -u
REPO=$(cd "$(dirname "$0")/.." && pwd)
REF=$REPO/oracle/_ref
W=${TMPDIR:-/tmp}/g16_$$
mkdir -p "$W" && cd "$W"
echo "host cores: $(nproc)"
t0=$(date +%s.%N)
$REF/generate_parameters ${1:-} > gen.log 2>&1
echo "generate_parameters ${1:-default}: $(echo "$(date +%s.%N) - $t0" | bc) s"
ls -la
for curve in MNT4753 MNT6753; do
  t0=$(date +%s.%N)
  $REF/main $curve compute $curve-parameters $curve-input $curve-output-ref > main_$curve.log 2>&1
  echo "reference CPU prover ($curve): $(echo "$(date +%s.%N) - $t0" | bc) s wall"
  grep -iE "total|time" main_$curve.log | tail -5
  $REF/b200_prover $curve compute $curve-parameters $curve-input $curve-output-b200 1 gpu-h 3
  $REF/b200_prover $curve compute $curve-parameters $curve-input $curve-output-b200-cpuh 1 cpu-h 2 | grep -E "compute_H|Total time"
  sha256sum $curve-output-ref $curve-output-b200 $curve-output-b200-cpuh
done
rm -rf "$W"
