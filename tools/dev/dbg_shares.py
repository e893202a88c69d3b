"""Development: one failing configuration under different share counts (B200MSM_SHARES is read per MSM)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import gpu_groth16_prover_3x_b200 as pkg
from oracle import pyoracle as po
orc = po.load_oracle()
curve, group = 0, 1
c = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100
shares = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1, 2, 3, 4, 8, 16, 58]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
bases = orc.gen_bases(curve, group, n)
sc = po.gen_scalars(curve, n, 3)
ctx = pkg.MsmContext(curve, 0)
ctx.set_table_budget(0)
slot = ctx.upload_bases(group, bases)
ctx.set_window_bits(c)
want, _ = orc.msm(curve, group, bases, sc)
for s in shares:
    os.environ["B200MSM_SHARES"] = str(s)
    res = []
    for rep in range(reps):
        got = orc.jacobian_to_affine(curve, group, ctx.msm(slot, sc, n))
        res.append(bool((got == want).all()))
    r = ctx.last_rounds()
    print("c=%d n=%d shares=%d" % (c, n, s), res, r["rounds"], r["pairs_per_round"], flush=True)
