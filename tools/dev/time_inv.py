"""Development: latency of one tile inversion (Team::inv_lane0) from the selftest kernel: 32 serial inversions per team."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import gpu_groth16_prover_3x_b200 as pkg
from oracle import pyoracle as po
for curve, group in ((0, 1), (0, 2), (1, 2)):
    ctx = pkg.MsmContext(curve, 0)
    p = po.fq_modulus(curve)
    deg = po.degree(curve, group)
    rng = np.random.default_rng(1)
    a = po.ints_to_array([int.from_bytes(rng.bytes(100), "little") % p for _ in range(32 * deg)])
    res = {}
    for op in (0, 8):
        ts = []
        for i in range(15):
            t0 = time.perf_counter(); ctx.selftest_field(group, op, a, a); ts.append(time.perf_counter() - t0)
        res[op] = sorted(ts)[len(ts) // 2]
    print("curve %d group %d: mul call %.3f ms, 32 inversions call %.3f ms -> %.1f us per inversion (incl. 2 lane copies)" % (
        curve, group, res[0] * 1e3, res[8] * 1e3, (res[8] - res[0]) / 32 * 1e6))
    ctx.close()
