#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the reference's own libff (oracle/_ref/libref.so).

Run in the builder container only (needs /root/reference to have been compiled by
`make -C oracle ref`).  The fixtures pin the plain-C oracle (and, through it, the CUDA engine) to the
reference: every array named `*_out` below was produced by unmodified libff code.  While generating,
the script also cross-checks the C oracle against libff on larger random sets that are not stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
os.makedirs(GOLD, exist_ok=True)
ref = po.load_reference()
assert ref is not None, "build oracle/_ref first: make -C oracle ref"
rng = np.random.default_rng(20261018)


def rand_fp(p, n):
    return [int.from_bytes(rng.bytes(100), "little") % p for _ in range(n)]


def edge_fp(p):
    return [0, 1, 2, p - 1, p - 2, po.R % p, (po.R * po.R) % p, (1 << 752), (1 << 32) - 1, (1 << 64) - 1]


# ---- generators ----------------------------------------------------------------------------
gens = {}
buf = np.zeros(72, np.uint64)
for c in (0, 1):
    for g, which in ((1, 5), (2, 7)):
        n = ref.lib.ref_get_constant(c, which, po._p(buf))
        gens["c%d_g%d" % (c, g)] = buf[:n].copy()
np.savez_compressed(os.path.join(GOLD, "generators.npz"), **gens)
po._load_generators()
orc = po.load_oracle()

# ---- field vectors -------------------------------------------------------------------------
fld = {}
for c in (0, 1):
    p = po.fq_modulus(c)
    for field in (0, 1):
        deg = 1 if field == 0 else po.degree(c, 2)
        base = edge_fp(p)
        # all edge pairs on component 0 plus random fill
        a_vals, b_vals = [], []
        for x in base:
            for y in base[:5]:
                a_vals.append([x] + rand_fp(p, deg - 1) if deg > 1 and x else [x] + [0] * (deg - 1))
                b_vals.append([y] + rand_fp(p, deg - 1) if deg > 1 and y else [y] + [0] * (deg - 1))
        for _ in range(40):
            a_vals.append(rand_fp(p, deg))
            b_vals.append(rand_fp(p, deg))
        a = po.ints_to_array([v for e in a_vals for v in e])
        b = po.ints_to_array([v for e in b_vals for v in e])
        key = "c%d_f%d" % (c, field)
        fld[key + "_a"], fld[key + "_b"] = a, b
        for op, name in enumerate(("mul", "add", "sub", "sqr", "inv", "neg")):
            out = ref.field_op(c, field, op, a, b if op < 3 else None)
            fld["%s_%s_out" % (key, name)] = out
            assert (orc.field_op(c, field, op, a, b if op < 3 else None) == out).all(), (key, name)
        # larger unstored cross-check
        a2 = po.ints_to_array(rand_fp(p, 2000 * deg)); b2 = po.ints_to_array(rand_fp(p, 2000 * deg))
        for op in (0, 1, 2, 3, 5):
            assert (orc.field_op(c, field, op, a2, b2) == ref.field_op(c, field, op, a2, b2)).all()
    r = po.fr_modulus(c)
    s = po.ints_to_array(edge_fp(r) + rand_fp(r, 30))
    fld["c%d_fr_mont" % c] = s
    fld["c%d_fr_plain_out" % c] = ref.fr_from_mont(c, s)
    assert (orc.fr_from_mont(c, s) == fld["c%d_fr_plain_out" % c]).all()
    assert (orc.fr_to_mont(c, fld["c%d_fr_plain_out" % c]) == s).all()
np.savez_compressed(os.path.join(GOLD, "field_vectors.npz"), **fld)

# ---- scalars: python SHA512_rng restatement vs libff ------------------------------------------
for c in (0, 1):
    assert (po.gen_scalars(c, 64, 7) == ref.gen_scalars(c, 64, 7)).all()

# ---- point vectors -------------------------------------------------------------------------
pts = {}
for c in (0, 1):
    for g in (1, 2):
        deg = po.degree(c, g)
        key = "c%d_g%d" % (c, g)
        bases = ref.gen_bases(c, g, 12)
        assert (orc.gen_bases(c, g, 12) == bases).all(), key
        P = bases.reshape(12, -1)
        zero = np.zeros(24 * deg, np.uint64)
        negP0 = ref.point_op(c, g, 4, P[0])
        pairs = [(P[0], P[1]), (P[2], P[3]), (P[0], P[0]), (P[0], negP0), (zero, P[1]), (P[1], zero), (zero, zero),
                 (P[5], P[7]), (P[11], P[4])]
        A = np.concatenate([x for x, _ in pairs]); B = np.concatenate([y for _, y in pairs])
        pts[key + "_a"], pts[key + "_b"] = A, B
        for op, name in ((0, "add"), (1, "dbl"), (2, "madd"), (4, "neg")):
            outs = np.concatenate([ref.point_op(c, g, op, x, y) for x, y in pairs])
            pts["%s_%s_out" % (key, name)] = outs
            mine = np.concatenate([orc.point_op(c, g, op, x, y) for x, y in pairs])
            assert (mine == outs).all(), (key, name)
        ks = po.ints_to_array([(v * po.R) % po.fr_modulus(c) for v in [0, 1, 2, 3, po.fr_modulus(c) - 1, 0x10000, (1 << 752) + 12345] + rand_fp(po.fr_modulus(c), 3)])
        pts[key + "_k"] = ks
        outs = np.concatenate([ref.point_op(c, g, 3, P[i % 12], k=ks[12 * i:12 * i + 12]) for i in range(ks.size // 12)])
        pts[key + "_smul_out"] = outs
        mine = np.concatenate([orc.point_op(c, g, 3, P[i % 12], k=ks[12 * i:12 * i + 12]) for i in range(ks.size // 12)])
        assert (mine == outs).all(), key
        # Jacobian import (read_pt): random Z, X = x Z^2, Y = y Z^3 built with libff field ops
        fieldsel = 0 if g == 1 else 1
        p = po.fq_modulus(c)
        jac = []
        for i in range(6):
            x, y = P[i][:12 * deg], P[i][12 * deg:]
            z = po.ints_to_array(rand_fp(p, deg))
            zz = ref.field_op(c, fieldsel, 3, z); zzz = ref.field_op(c, fieldsel, 0, zz, z)
            jac.append(np.concatenate([ref.field_op(c, fieldsel, 0, x, zz), ref.field_op(c, fieldsel, 0, y, zzz), z]))
        jac.append(np.concatenate([po.ints_to_array([po.R % p] + [0] * (deg - 1))] * 2 + [np.zeros(12 * deg, np.uint64)]))  # (1,1,0) = infinity
        J = np.concatenate(jac)
        pts[key + "_jac"] = J
        outs = np.concatenate([ref.jacobian_to_affine(c, g, J[i * 36 * deg:(i + 1) * 36 * deg]) for i in range(7)])
        pts[key + "_jac_out"] = outs
        assert (outs[:6 * 24 * deg] == P[:6].reshape(-1)).all() and not outs[6 * 24 * deg:].any()
        mine = np.concatenate([orc.jacobian_to_affine(c, g, J[i * 36 * deg:(i + 1) * 36 * deg]) for i in range(7)])
        assert (mine == outs).all(), key
np.savez_compressed(os.path.join(GOLD, "point_vectors.npz"), **pts)

# ---- MSM vectors ---------------------------------------------------------------------------
msm = {}
SIZES = (0, 1, 2, 31, 32, 33, 100, 257)
for c in (0, 1):
    r = po.fr_modulus(c)
    for g in (1, 2):
        deg = po.degree(c, g)
        key = "c%d_g%d" % (c, g)
        nmax = max(SIZES)
        bases = ref.gen_bases(c, g, nmax)
        # plant infinity bases (y == 0 encoding) and a duplicate / negated pair
        bases = bases.reshape(nmax, -1).copy()
        bases[5] = 0
        bases[40] = bases[41]
        bases[43] = ref.point_op(c, g, 4, bases[42])
        bases = bases.reshape(-1)
        sc = ref.gen_scalars(c, nmax, 11).reshape(nmax, 12).copy()
        special = [0, 1, 2, r - 1, 1 << 16, (1 << 15), (1 << 16) - 1, (1 << 752), 0xFFFF8000FFFF8000]
        for i, v in enumerate(special):
            sc[10 + i] = po.int_to_limbs((v * po.R) % r)
        sc[40] = sc[41]          # same scalar on duplicate bases  -> P + P inside one bucket
        sc[43] = sc[42]          # same scalar on P and -P         -> P + (-P) inside one bucket
        sc = sc.reshape(-1)
        msm[key + "_bases"], msm[key + "_scalars"] = bases, sc
        for n in SIZES:
            b, s = bases[:n * 24 * deg], sc[:n * 12]
            want, _ = ref.msm(c, g, b, s, method=1, chunks=0, prefilter=1)     # what ./main computes
            for method, chunks, pre in ((0, 1, 0), (2, 0, 1), (1, 3, 0)):
                got, _ = ref.msm(c, g, b, s, method=method, chunks=chunks, prefilter=pre)
                assert (got == want).all(), (key, n, method)
            for method, chunks, pre in ((1, 0, 1), (0, 2, 0), (1, 1, 0)):
                got, _ = orc.msm(c, g, b, s, method=method, chunks=chunks, prefilter=pre)
                assert (got == want).all(), ("oracle", key, n, method)
            msm["%s_n%d_out" % (key, n)] = want
        # closed form on the unmodified structured bases
        b = ref.gen_bases(c, g, 300); s = ref.gen_scalars(c, 300, 5)
        want, _ = ref.msm(c, g, b, s)
        assert (orc.msm_closed_form(c, g, s) == want).all(), key
        assert (orc.msm(c, g, b, s)[0] == want).all(), key
        print("msm golden", key, "ok", flush=True)
np.savez_compressed(os.path.join(GOLD, "msm_vectors.npz"), **msm)
print("golden fixtures written to", GOLD)
for f in sorted(os.listdir(GOLD)):
    print("  %-24s %8d B" % (f, os.path.getsize(os.path.join(GOLD, f))))

# ---- H polynomial: compute_H of the reference (libfqfft through libref.so) on seeded inputs ---------------------
hv = {}
for curve in (0, 1):
    for logm in (1, 3, 6, 9):
        m = 1 << logm
        ca, cb, cc = (po.gen_scalars(curve, m, 900 + 10 * logm + k) for k in range(3))
        hv["c%d_m%d_ca" % (curve, m)], hv["c%d_m%d_cb" % (curve, m)], hv["c%d_m%d_cc" % (curve, m)] = ca, cb, cc
        hv["c%d_m%d_out" % (curve, m)] = ref.compute_h(curve, ca, cb, cc)
np.savez_compressed(os.path.join(GOLD, "h_vectors.npz"), **hv)
print("wrote h_vectors.npz")

