"""Development: smallest failing n at one window width, one share."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import gpu_groth16_prover_3x_b200 as pkg
from oracle import pyoracle as po
orc = po.load_oracle()
curve, group = 0, 1
c = int(sys.argv[1]) if len(sys.argv) > 1 else 5
ns = [int(x) for x in sys.argv[2].split(",")]
os.environ["B200MSM_SHARES"] = sys.argv[3] if len(sys.argv) > 3 else "1"
nmax = max(ns)
bases = orc.gen_bases(curve, group, nmax)
sc = po.gen_scalars(curve, nmax, 3)
ctx = pkg.MsmContext(curve, 0)
ctx.set_table_budget(0)
slot = ctx.upload_bases(group, bases)
ctx.set_window_bits(c)
for n in ns:
    want, _ = orc.msm(curve, group, bases[:n * 24], sc[:n * 12])
    got = orc.jacobian_to_affine(curve, group, ctx.msm(slot, sc[:n * 12], n))
    r = ctx.last_rounds()
    print("c=%d n=%d" % (c, n), bool((got == want).all()), r["rounds"], r["max_bucket_occupancy"], r["pairs_per_round"], flush=True)
