"""Development: compare every bucket left by the batched-affine accumulation with the oracle's sum of its entries."""
import sys, os, ctypes
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import gpu_groth16_prover_3x_b200 as pkg
from oracle import pyoracle as po
orc = po.load_oracle()
curve, group = 0, 1
c = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n = int(sys.argv[2]) if len(sys.argv) > 2 else 65
os.environ["B200MSM_SHARES"] = sys.argv[3] if len(sys.argv) > 3 else "1"
bases = orc.gen_bases(curve, group, n)
sc = po.gen_scalars(curve, n, 3)
ctx = pkg.MsmContext(curve, 0)
ctx.set_table_budget(0)
slot = ctx.upload_bases(group, bases)
ctx.set_window_bits(c)
got = orc.jacobian_to_affine(curve, group, ctx.msm(slot, sc, n))
want, _ = orc.msm(curve, group, bases, sc)
print("msm ok:", bool((got == want).all()), ctx.last_rounds())
lib = ctx.lib
dbg = (ctypes.c_uint64 * 16)()
lib.b200msm_internal_debug_layout(ctx._h, 0, dbg)
K, o_offs, o_r0, o_r1, o_scr, o_bref, o_cntv, capA, emax, o_sc, W, NB, U, o_bndr, o_bndb, o_pairs = [int(x) for x in dbg]
def rd(off, nbytes, dt=np.uint32):
    out = np.zeros(nbytes // np.dtype(dt).itemsize, dt)
    rc = lib.b200msm_internal_debug_read(ctx._h, 0, ctypes.c_uint64(off), ctypes.c_uint64(nbytes), ctypes.c_void_p(out.ctypes.data))
    assert rc == 0, rc
    return out
offs = rd(o_offs, (K + 1) * 4)
bref = rd(o_bref, K * 4)
E = int(offs[K])
print("K", K, "E", E, "W", W, "NB", NB, "capA", capA)
r = po.fr_modulus(curve)
ks = [po.limbs_to_int(x) for x in orc.fr_from_mont(curve, sc).reshape(n, 12)]
Wd = (754 + c - 1) // c
members = {}
for i, k in enumerate(ks):
    carry = 0
    for w in range(Wd):
        raw = ((k >> (w * c)) & ((1 << c) - 1)) + carry
        if raw > (1 << (c - 1)):
            d = raw - (1 << c); carry = 1
        else:
            d = raw; carry = 0
        if d:
            members.setdefault((w % W) * NB + abs(d) - 1, []).append((i, d < 0))
one = po.ints_to_array([po.R % r]); mone = po.ints_to_array([(r - 1) * po.R % r])
bad = 0
for b in range(K):
    cnt = int(offs[b + 1] - offs[b])
    mem = members.get(b, [])
    assert cnt == len(mem), (b, cnt, len(mem))
    ref = int(bref[b])
    if cnt == 0:
        if ref != 0xffffffff: print("bucket", b, "empty but ref", hex(ref)); bad += 1
        continue
    bb = np.concatenate([bases[i * 24:(i + 1) * 24] for i, _ in mem])
    ss = np.concatenate([mone if neg else one for _, neg in mem])
    w_aff, _ = orc.msm(curve, group, bb, ss)
    if ref == 0xffffffff:
        g_aff = np.zeros(24, np.uint64)
    elif ref & 0x40000000:
        g_aff = rd(o_scr + (ref & 0x3fffffff) * 192, 192, np.uint64)
    else:
        g_aff = bases[(ref & 0x3fffffff) * 24:(ref & 0x3fffffff) * 24 + 24].copy()
        if ref >> 31:
            pass  # sign: compare x only
    ok = (g_aff[:12] == w_aff[:12]).all() and ((ref >> 31) or (ref == 0xffffffff) or (g_aff[12:] == w_aff[12:]).all())
    if ref == 0xffffffff: ok = not w_aff.any()
    if not ok:
        bad += 1
        if bad <= 40:
            print("BAD bucket", b, "set", b // NB, "count", cnt, "offs", int(offs[b]), "ref", hex(ref), "members", mem[:20])
print("bad buckets:", bad, "of", K)
