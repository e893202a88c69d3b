"""B200-native MSM engine for the MNT4753 / MNT6753 Groth16 prover (drop-in for the multiexp path of
vezenovm/gpu-groth16-prover-3x: cuda_prover_piecewise.cu + prover_reference_functions multiexp_G1/G2).

The compute path is hand-written CUDA for sm_100a behind the C ABI of include/b200_msm.h
(csrc/msm.cu -> libb200msm.so).  This package is the thin host mirror of that ABI; it has no CPU
fallback and raises if the CUDA library is missing.
"""
from .engine import (G1, G2, MNT4753, MNT6753, MsmContext, MsmError, degree, library_path, load_key_sharded_file,
                     load_library, prove_sharded, shard_ranges)

__all__ = ["G1", "G2", "MNT4753", "MNT6753", "MsmContext", "MsmError", "degree", "library_path", "load_key_sharded_file",
           "load_library", "prove_sharded", "shard_ranges"]
