#!/bin/bash
# Proof latency of the default MNT4753 instance (d = 2^20 - 1) on 1, 2, 4, ... GPUs of one box: every query sharded by
# point range (tests/integration/b200_prover.cpp), H polynomial on GPU 0.  Usage: tools/proof_scaling.sh "1 2 4"
set -u
REPO=$(cd "$(dirname "$0")/.." && pwd)
REF=$REPO/oracle/_ref
W=${TMPDIR:-/tmp}/g16s_$$
mkdir -p "$W" && cd "$W"
$REF/generate_parameters > gen.log 2>&1
for g in ${1:-1 2}; do
  echo "=== $g GPU(s)"
  $REF/b200_prover MNT4753 compute MNT4753-parameters MNT4753-input out-$g $g gpu-h 3 | grep -E "upload|compute_H|gpu e2e|Total time"
  sha256sum out-$g
done
rm -rf "$W"
