#!/bin/bash
# Proof latency through the product's command-line prover on N GPUs of one box, witness MSMs one after the other
# (B200MSM_LANE_SPLIT=0) against the automatic lane split (1), on synthetic instances of the reference's default sizes;
# every run must write the same proof bytes.   tools/proof_ngpu.sh <gpus> ["<curve>:<log2(d+1)> ..."]
set -u
REPO=$(cd "$(dirname "$0")/.." && pwd)
GPUS=${1:-1}
W=${TMPDIR:-/tmp}/pngpu_$$
mkdir -p "$W"
for spec in ${2:-MNT4753:20 MNT6753:15}; do
  curve=${spec%%:*}; k=${spec##*:}
  dir=$W/$curve-$k; mkdir -p "$dir"
  (cd "$REPO" && python - "$curve" "$k" "$dir") <<'PY'
import sys
sys.path.insert(0, ".")
import gpu_groth16_prover_3x_b200 as pkg
from gpu_groth16_prover_3x_b200 import synthetic
curve, k, d = sys.argv[1], int(sys.argv[2]), sys.argv[3]
synthetic.write_instance(pkg.MNT4753 if curve == "MNT4753" else pkg.MNT6753, (1 << k) - 1, d, 0)
PY
  for mode in ${MODES:-0 1}; do
    out=$(B200MSM_TRACE=${TRACE:-} B200MSM_LANE_SPLIT=$mode "$REPO/gpu_groth16_prover_3x_b200/b200_prove" "$curve" compute "$dir/$curve-parameters" "$dir/$curve-input" "$dir/out$mode" ${REPEATS:-5} "$GPUS" 2>"$dir/trace$mode")
    times=$(echo "$out" | grep -oE "input to output: [0-9.]+" | grep -oE "[0-9.]+$" | tr '\n' ' ')
    load=$(echo "$out" | grep -oE "window tables: [0-9.]+" | grep -oE "[0-9.]+$")
    echo "$curve d+1=2^$k gpus=$GPUS split=$mode  key load $load ms  proofs ms: $times  sha $(sha256sum "$dir/out$mode" | cut -c1-16)"
    if [ -n "${TRACE:-}" ]; then grep -E "shard [01] " "$dir/trace$mode" | tail -13; fi
  done
  rm -rf "$dir"
done
rm -rf "$W"
