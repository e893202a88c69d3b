"""Development: which bucket occupancy breaks?  Equal scalars put all n points of a window into ONE bucket."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import gpu_groth16_prover_3x_b200 as pkg
from oracle import pyoracle as po
orc = po.load_oracle()
curve, group = 0, 1
c = int(sys.argv[1]) if len(sys.argv) > 1 else 9
nmax = int(sys.argv[2]) if len(sys.argv) > 2 else 70
bases = orc.gen_bases(curve, group, nmax)
one = po.gen_scalars(curve, 1, 3)
sc = np.tile(one, nmax)
ctx = pkg.MsmContext(curve, 0)
ctx.set_table_budget(0)
slot = ctx.upload_bases(group, bases)
ctx.set_window_bits(c)
bad = []
for n in range(1, nmax + 1):
    got = orc.jacobian_to_affine(curve, group, ctx.msm(slot, sc[:n * 12], n))
    want, _ = orc.msm(curve, group, bases[:n * 24], sc[:n * 12])
    ok = (got == want).all()
    r = ctx.last_rounds()
    if not ok:
        bad.append(n)
    print(n, "ok" if ok else "FAIL", r["rounds"], r["shares"], r["pairs_per_round"], flush=True)
print("failing n:", bad)
