#!/bin/bash
# A/B timing of library variants on the GPU box: tools/ab.sh "<lib1> <lib2> ..." (file names inside the package dir)
cd "$(dirname "$0")/.."
for lib in $1; do
  export B200MSM_LIB=$PWD/gpu_groth16_prover_3x_b200/$lib
  echo "=== $lib"
  timeout 120 python tools/ncu_one_msm.py 20 0 1 4 2>&1 | grep -E "msm [123]|rounds"
  timeout 120 python tools/ncu_one_msm.py 19 0 2 3 2>&1 | grep -E "msm [12]"
  timeout 120 python tools/ncu_one_msm.py 18 1 2 3 2>&1 | grep -E "msm [12]"
  timeout 120 python tools/ncu_one_msm.py 16 0 1 4 2>&1 | grep -E "msm [23]"
done
