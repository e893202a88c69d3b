"""MSM size sweep with the automatic window width (BASELINE.json configs[4]).
   python tools/size_sweep.py curve group log_lo log_hi [step]   -> one line per size (best of 3 after a warm-up)"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gpu_groth16_prover_3x_b200 as pkg
from gpu_groth16_prover_3x_b200 import synthetic
curve, group, lo, hi = (int(x) for x in sys.argv[1:5])
step = int(sys.argv[5]) if len(sys.argv) > 5 else 2
ctx = pkg.MsmContext(curve, 0)
k0, k1 = synthetic.base_seed_scalars(curve)
names = {(0, 1): "MNT4753 G1", (0, 2): "MNT4753 G2", (1, 1): "MNT6753 G1", (1, 2): "MNT6753 G2"}
for log_n in range(lo, hi + 1, step):
    n = 1 << log_n
    sc = torch.from_numpy(synthetic.random_scalars(curve, n, 5).view(np.int64)).cuda()
    slot = ctx.synthetic_bases(group, n, k0, k1)
    info = ctx.bases_info(slot)
    best = None
    for i in range(4):
        ctx.msm(slot, sc, n)
        t = ctx.last_timings()
        if i and (best is None or t["total"] < best["total"]):
            best = t
    r = ctx.last_rounds()
    print("%s 2^%-2d | %8.2f ms  %6.2f M points/s | c %2d tables %2d sets %d rounds %2d | sort %5.2f acc %7.2f red %5.2f | tables %5.1f GB built in %5.2f s" % (
        names[(curve, group)], log_n, best["total"], n / best["total"] / 1e3, best["window_bits"], best["tables"], best["bucket_sets"], r["rounds"],
        best["recode_sort"], best["accumulate"], best["reduce_combine"], info["bytes"] / 1e9, info["table_build_ms"] / 1e3), flush=True)
    ctx.free_bases(slot)
    del sc
ctx.close()
