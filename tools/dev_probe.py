#!/usr/bin/env python3
"""Developer probe (not the bench): phase timings of one MSM configuration + microbenchmarks.
Uses the oracle only to make structured bases; numbers from here are never reported."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gpu_groth16_prover_3x_b200 as pkg  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--curve", type=int, default=0)
ap.add_argument("--group", type=int, default=1)
ap.add_argument("--log-n", type=int, default=20)
ap.add_argument("--c", type=int, nargs="*", default=[0])
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--micro", action="store_true")
ap.add_argument("--check", action="store_true")
a = ap.parse_args()

orc = po.load_oracle()
n = 1 << a.log_n
deg = po.degree(a.curve, a.group)
t0 = time.time()
p0, q = orc.base_pair(a.curve, a.group)
bases = np.zeros(n * 24 * deg, np.uint64)
orc._f("gen_bases")(a.curve, a.group, n, po._p(p0), po._p(q), po._p(bases))
rng = np.random.default_rng(7)
raw = rng.integers(0, 1 << 63, size=(n, 12), dtype=np.uint64)
raw[:, 11] &= np.uint64((1 << 47) - 1)
sc = raw.reshape(-1)
print("inputs ready in %.1fs" % (time.time() - t0), flush=True)
ctx = pkg.MsmContext(a.curve, 0)
if a.micro:
    for kind, name in ((0, "IMAD.WIDE GMAC/s"), (1, "IMAD.LO Gop/s"), (2, "Fq modmul G/s"), (3, "Fq modmul RR29 G/s")):
        print("microbench %-18s %.1f" % (name, ctx.microbench(kind, 2048)), flush=True)
slot = ctx.upload_bases(a.group, bases)
want = orc.msm_closed_form(a.curve, a.group, sc) if a.check else None
for c in a.c:
    ctx.set_window_bits(c)
    for r in range(a.reps):
        t0 = time.time()
        out = ctx.msm(slot, sc)
        wall = time.time() - t0
        t = ctx.last_timings()
        print("c=%2d W=%d wall %.1f ms | total %.2f h2d %.2f sort %.2f acc %.2f reduce %.2f | %.2f Mpts/s" % (
            t["window_bits"], t["windows"], wall * 1e3, t["total"], t["h2d_scalars"], t["recode_sort"], t["accumulate"],
            t["reduce_combine"], n / t["total"] / 1e3), flush=True)
    if want is not None:
        assert (orc.jacobian_to_affine(a.curve, a.group, out) == want).all()
        print("  parity ok")
