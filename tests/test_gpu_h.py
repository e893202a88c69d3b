"""GPU parity of the H-polynomial path (b200msm_compute_h: the seven FFTs of compute_H on the device) against
the fixtures produced by the reference's libfqfft, against the oracle on seeded inputs, and -- at the reference's
default size -- through properties that hold for any input."""
import numpy as np
import pytest

import gpu_groth16_prover_3x_b200 as pkg
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctxs():
    d = {c: pkg.MsmContext(c, 0) for c in (0, 1)}
    yield d
    for c in d.values():
        c.close()


@pytest.mark.parametrize("curve", [0, 1])
def test_golden_vectors(ctxs, golden, curve):
    z = golden["h_vectors"]
    for m in (2, 8, 64, 512):
        k = "c%d_m%d_" % (curve, m)
        got, dev = ctxs[curve].compute_h(z[k + "ca"], z[k + "cb"], z[k + "cc"])
        assert dev and (got == z[k + "out"]).all(), m


@pytest.mark.parametrize("curve,logm", [(0, 12), (1, 12), (1, 15)])
def test_seeded_vs_oracle(ctxs, oracle, curve, logm):
    m = 1 << logm
    ca, cb, cc = (po.gen_scalars(curve, m, 40 + k) if logm <= 12 else
                  np.random.default_rng(40 + k).integers(0, 1 << 62, size=m * 12, dtype=np.uint64) for k in range(3))
    if logm > 12:   # cheap canonical residues: top limb below the modulus' top limb
        for x in (ca, cb, cc):
            x.reshape(m, 12)[:, 11] &= np.uint64((1 << 47) - 1)
    want = oracle.compute_h(curve, ca, cb, cc)
    got, _ = ctxs[curve].compute_h(ca, cb, cc)
    assert (got == want).all()
    assert not got[-12:].any()       # vector_Fr_zeros(m + 1): the extra coefficient is zero


def test_device_resident_result_feeds_the_msm(ctxs, oracle):
    """The H-query MSM takes its scalars straight from the device buffer compute_h leaves behind."""
    curve, m = 0, 1 << 10
    ctx = ctxs[curve]
    ca, cb, cc = (po.gen_scalars(curve, m, 70 + k) for k in range(3))
    h_host, dev = ctx.compute_h(ca, cb, cc)
    _, dev2 = ctx.compute_h(ca, cb, cc, to_host=False)
    assert dev2 == dev
    d = m - 1
    bases = oracle.gen_bases(curve, 1, d)
    slot = ctx.upload_bases(1, bases)
    want, _ = oracle.msm(curve, 1, bases, h_host[:d * 12])
    got = oracle.jacobian_to_affine(curve, 1, ctx.msm(slot, dev, d))
    assert (got == want).all()
    ctx.free_bases(slot)


def test_domain_limits(ctxs):
    with pytest.raises(pkg.MsmError):
        x = np.zeros(12 * 6, np.uint64)
        ctxs[0].compute_h(x, x, x)                      # 6 is not a power of two
    with pytest.raises(pkg.MsmError):
        x = np.zeros(12 * (1 << 16), np.uint64)
        ctxs[1].compute_h(x, x, x)                      # Fr(MNT6753) has 2-adicity 15 (mnt6753_init.cpp:66)


def test_default_size_properties(ctxs):
    """2^20 points (the reference's default MNT4753 instance): H is linear in cc and bilinear in (ca, cb);
    with cc = ca * cb pointwise the quotient vanishes identically."""
    import torch
    curve, m = 0, 1 << 20
    ctx = ctxs[curve]
    rng = np.random.default_rng(3)
    def rand():
        x = rng.integers(0, 1 << 62, size=(m, 12), dtype=np.uint64)
        x[:, 11] &= np.uint64((1 << 47) - 1)
        return x.reshape(-1)
    ca, cb = rand(), rand()
    zero = np.zeros(m * 12, np.uint64)
    one = np.tile(po.int_to_limbs(po.R % po.fr_modulus(curve)), m)
    # (ca * 1 - ca) / Z == 0
    h, _ = ctx.compute_h(ca, one, ca)
    assert not h.any()
    # H(ca, cb, 0) with cb = 1: equals H(ca, 1, 0); and H(ca, 1, 0) + H(0, 0, ca) == 0 (linearity in cc, sign)
    h1, _ = ctx.compute_h(ca, one, zero)
    h2, _ = ctx.compute_h(zero, zero, ca)
    r = po.fr_modulus(curve)
    a = h1.reshape(-1, 12)[:64]
    b = h2.reshape(-1, 12)[:64]
    for x, y in zip(a, b):
        assert (po.limbs_to_int(x) + po.limbs_to_int(y)) % r == 0
    t = ctx.compute_h_timings()
    assert t["compute_h_ms"] > 0
    dev = torch.from_numpy(ca.view(np.int64)).cuda()
    h3, _ = ctx.compute_h(dev, torch.from_numpy(one.view(np.int64)).cuda(), torch.from_numpy(zero.view(np.int64)).cuda())
    assert (h3 == h1).all()
    print("compute_h 2^20:", ctx.compute_h_timings())
