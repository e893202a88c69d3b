#!/bin/bash
# A/B of the prover's lane split (csrc/prover.cu lane_split_plan) on one GPU: synthetic instances of several sizes,
# b200_prove with B200MSM_LANE_SPLIT = 0 (witness MSMs one after the other), 1 (automatic), 2 (always side by side);
# prints the proof times and checks that every mode writes the same proof bytes.
#   tools/lane_split_ab.sh "<curve>:<log2(d+1)> ..."     default: MNT6753:15 MNT4753:14 MNT4753:17 MNT4753:18
set -u
REPO=$(cd "$(dirname "$0")/.." && pwd)
W=${TMPDIR:-/tmp}/lsab_$$
mkdir -p "$W"
for spec in ${1:-MNT6753:15 MNT4753:14 MNT4753:17 MNT4753:18}; do
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
  for mode in ${MODES:-0 1 2}; do
    out=$(B200MSM_LANE_SPLIT=$mode "$REPO/gpu_groth16_prover_3x_b200/b200_prove" "$curve" compute "$dir/$curve-parameters" "$dir/$curve-input" "$dir/out$mode" ${REPEATS:-6} 1)
    times=$(echo "$out" | grep -oE "input to output: [0-9.]+" | grep -oE "[0-9.]+$" | tr '\n' ' ')
    echo "$curve d+1=2^$k split=$mode  ms: $times  sha $(sha256sum "$dir/out$mode" | cut -c1-16)"
    if [ -n "${TRACE:-}" ]; then B200MSM_TRACE=1 B200MSM_LANE_SPLIT=$mode "$REPO/gpu_groth16_prover_3x_b200/b200_prove" "$curve" compute "$dir/$curve-parameters" "$dir/$curve-input" "$dir/out$mode" 2 1 2>&1 >/dev/null | tail -6; fi
  done
  rm -rf "$dir"
done
rm -rf "$W"
