"""Development: parity of small MSMs over a grid of sizes / window widths (prints a table instead of asserting)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import gpu_groth16_prover_3x_b200 as pkg
from oracle import pyoracle as po
orc = po.load_oracle()
curve = int(sys.argv[1]) if len(sys.argv) > 1 else 0
group = int(sys.argv[2]) if len(sys.argv) > 2 else 1
sizes = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [100, 150, 200, 256, 257, 300, 513, 1000, 3000]
cs = [int(x) for x in sys.argv[4].split(",")] if len(sys.argv) > 4 else [0, 5, 9, 14]
nmax = max(sizes)
bases = orc.gen_bases(curve, group, nmax)
sc = po.gen_scalars(curve, nmax, 3)
deg = po.degree(curve, group)
ctx = pkg.MsmContext(curve, 0)
for tables in (True, False):
    if not tables:
        ctx.set_table_budget(0)
    slot = ctx.upload_bases(group, bases)
    print("tables" if tables else "no tables", ctx.bases_info(slot))
    for c in cs:
        ctx.set_window_bits(c)
        for n in sizes:
            got = orc.jacobian_to_affine(curve, group, ctx.msm(slot, sc[:n * 12], n))
            want, _ = orc.msm(curve, group, bases[:n * 24 * deg], sc[:n * 12])
            t = ctx.last_timings(); r = ctx.last_rounds()
            print("c=%2d n=%5d %s  c_used=%d W=%d sets=%d shares=%d rounds=%d maxocc=%d adds=%d  %.2f ms" % (
                c, n, "ok  " if (got == want).all() else "FAIL", t["window_bits"], t["windows"], t["bucket_sets"], t["shares"],
                r["rounds"], r["max_bucket_occupancy"], r["additions"], t["total"]), flush=True)
    ctx.set_window_bits(0)
    ctx.free_bases(slot)
