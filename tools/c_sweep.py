"""Window-width sweep: MSM time vs c (tables rebuilt for every c).  python tools/c_sweep.py log_n curve group c_lo c_hi"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gpu_groth16_prover_3x_b200 as pkg
from gpu_groth16_prover_3x_b200 import synthetic
log_n, curve, group, c_lo, c_hi = (int(x) for x in sys.argv[1:6])
n = 1 << log_n
ctx = pkg.MsmContext(curve, 0)
k0, k1 = synthetic.base_seed_scalars(curve)
sc = torch.from_numpy(synthetic.random_scalars(curve, n, 5).view(np.int64)).cuda()
for c in range(c_lo, c_hi + 1):
    ctx.set_window_bits(c)
    slot = ctx.synthetic_bases(group, n, k0, k1)
    info = ctx.bases_info(slot)
    best = None
    for i in range(4):
        ctx.msm(slot, sc, n)
        t = ctx.last_timings()
        if best is None or t["total"] < best["total"]:
            best = t
    print("n=2^%d c=%2d W=%3d tables=%3d sets=%d | total %7.3f ms  sort %6.3f acc %7.3f red %6.3f | build %.0f ms" % (
        log_n, c, best["windows"], best["tables"], best["bucket_sets"], best["total"], best["recode_sort"], best["accumulate"],
        best["reduce_combine"], info["table_build_ms"]), flush=True)
    ctx.free_bases(slot)
ctx.close()
