"""Development: smallest random MSMs that fail."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import gpu_groth16_prover_3x_b200 as pkg
from oracle import pyoracle as po
orc = po.load_oracle()
curve, group = 0, 1
c = int(sys.argv[1]) if len(sys.argv) > 1 else 5
nmax = 40
bases = orc.gen_bases(curve, group, nmax)
ctx = pkg.MsmContext(curve, 0)
ctx.set_table_budget(0)
slot = ctx.upload_bases(group, bases)
ctx.set_window_bits(c)
r = po.fr_modulus(curve)
for mode in ("random", "positive-small", "two-windows"):
    for n in (2, 3, 4, 5, 6, 8, 12, 16, 24, 40):
        fails = 0
        for seed in range(4):
            if mode == "random":
                sc = po.gen_scalars(curve, n, 100 + seed)
            elif mode == "positive-small":
                rng = np.random.default_rng(seed)
                sc = po.ints_to_array([int(rng.integers(1, 1 << (c - 1))) * po.R % r for _ in range(n)])      # one window, digits > 0
            else:
                rng = np.random.default_rng(seed)
                sc = po.ints_to_array([int(rng.integers(1, 1 << (2 * c))) * po.R % r for _ in range(n)])      # two windows, signed digits
            got = orc.jacobian_to_affine(curve, group, ctx.msm(slot, sc, n))
            want, _ = orc.msm(curve, group, bases[:n * 24], sc)
            fails += not (got == want).all()
        rr = ctx.last_rounds()
        print(mode, "n=%d" % n, "fails %d/4" % fails, rr["rounds"], rr["pairs_per_round"], flush=True)
