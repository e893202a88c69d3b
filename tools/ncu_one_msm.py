"""Profiling driver: one base set, a few MSMs (ncu attaches to this; see profiles/).  Usage:
   python tools/ncu_one_msm.py [log_n] [curve] [group] [reps]"""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gpu_groth16_prover_3x_b200 as pkg
from gpu_groth16_prover_3x_b200 import synthetic

log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
curve = int(sys.argv[2]) if len(sys.argv) > 2 else 0
group = int(sys.argv[3]) if len(sys.argv) > 3 else 1
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
n = 1 << log_n
ctx = pkg.MsmContext(curve, 0)
if os.environ.get("B200MSM_C"):
    ctx.set_window_bits(int(os.environ["B200MSM_C"]))
k0, k1 = synthetic.base_seed_scalars(curve)
slot = ctx.synthetic_bases(group, n, k0, k1)
print(ctx.bases_info(slot))
sc = torch.from_numpy(synthetic.random_scalars(curve, n, 5).view(np.int64)).cuda()
for i in range(reps):
    t0 = time.perf_counter()
    ctx.msm(slot, sc, n)
    print("msm %d: wall %.2f ms" % (i, (time.perf_counter() - t0) * 1e3), ctx.last_timings())
print(ctx.last_rounds())
ctx.close()
