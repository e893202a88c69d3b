"""GPU parity of the device-side normalisation and the synthetic base generator."""
import numpy as np
import pytest

import gpu_groth16_prover_3x_b200 as pkg
from gpu_groth16_prover_3x_b200 import synthetic
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
CG = [(c, g) for c in (0, 1) for g in (1, 2)]


@pytest.fixture(scope="module")
def ctxs():
    d = {c: pkg.MsmContext(c, 0) for c in (0, 1)}
    yield d
    for c in d.values():
        c.close()


@pytest.mark.parametrize("curve,group", CG)
def test_to_affine(ctxs, oracle, golden, curve, group):
    z = golden["point_vectors"]
    key = "c%d_g%d" % (curve, group)
    J = z[key + "_jac"]
    got = ctxs[curve].to_affine(group, J)
    assert (got == z[key + "_jac_out"]).all()  # written by libff's write_g1/g2 after read_pt_*


@pytest.mark.parametrize("curve,group", CG)
def test_synthetic_bases_match_oracle(ctxs, oracle, curve, group):
    n = 1000  # not a multiple of the 64-point batch
    k0, k1 = synthetic.base_seed_scalars(curve)
    ctx = ctxs[curve]
    slot = ctx.synthetic_bases(group, n, k0, k1)
    got = ctx.download_bases(slot)
    assert (got == oracle.gen_bases(curve, group, n)).all()
    # and the engine's MSM over them equals the closed form
    sc = synthetic.random_scalars(curve, n, 3)
    want = oracle.msm_closed_form(curve, group, sc)
    assert (ctx.to_affine(group, ctx.msm(slot, sc)) == want).all()
    ctx.free_bases(slot)


def test_seed_scalars_match_oracle():
    for c in (0, 1):
        r = po.fr_modulus(c)
        k0, k1 = synthetic.base_seed_scalars(c)
        assert po.limbs_to_int(k0) == po.sha512_rng_ints(r, 1000001, 1)[0] * po.R % r
        assert po.limbs_to_int(k1) == po.sha512_rng_ints(r, 1000002, 1)[0] * po.R % r
