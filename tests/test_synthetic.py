import numpy as np

from gpu_groth16_prover_3x_b200 import synthetic
from oracle import pyoracle as po


def test_random_scalars_canonical_and_deterministic():
    for c in (0, 1):
        r = synthetic.fr_modulus(c)
        assert r == po.fr_modulus(c)
        a = synthetic.random_scalars(c, 5000, 9)
        assert a.shape == (60000,) and (a == synthetic.random_scalars(c, 5000, 9)).all()
        vals = [po.limbs_to_int(x) for x in a.reshape(-1, 12)]
        assert all(v < r for v in vals) and len(set(vals)) == 5000
        assert max(vals) > (r >> 1)  # the top bit is exercised


def test_sha512_rng_matches_oracle_restatement():
    for c in (0, 1):
        r = po.fr_modulus(c)
        assert [synthetic.sha512_rng(r, (7 << 32) + i) for i in range(8)] == po.sha512_rng_ints(r, 7 << 32, 8)
