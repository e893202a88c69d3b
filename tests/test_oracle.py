"""The plain-C oracle replayed against the golden fixtures produced by the reference's own libff
(tools/gen_golden.py, run where /root/reference was available)."""
import numpy as np
import pytest

from oracle import pyoracle as po

CG = [(c, g) for c in (0, 1) for g in (1, 2)]


@pytest.mark.parametrize("curve", [0, 1])
@pytest.mark.parametrize("field", [0, 1])
def test_field_ops(oracle, golden, curve, field):
    z = golden["field_vectors"]
    key = "c%d_f%d" % (curve, field)
    a, b = z[key + "_a"], z[key + "_b"]
    for op, name in enumerate(("mul", "add", "sub", "sqr", "inv", "neg")):
        out = oracle.field_op(curve, field, op, a, b if op < 3 else None)
        assert (out == z["%s_%s_out" % (key, name)]).all(), name


@pytest.mark.parametrize("curve", [0, 1])
def test_fr_montgomery(oracle, golden, curve):
    z = golden["field_vectors"]
    s = z["c%d_fr_mont" % curve]
    plain = oracle.fr_from_mont(curve, s)
    assert (plain == z["c%d_fr_plain_out" % curve]).all()
    assert (oracle.fr_to_mont(curve, plain) == s).all()
    r = po.fr_modulus(curve)
    for i in range(s.size // 12):
        assert po.limbs_to_int(plain[12 * i:12 * i + 12]) * po.R % r == po.limbs_to_int(s[12 * i:12 * i + 12])


@pytest.mark.parametrize("curve,group", CG)
def test_point_ops(oracle, golden, curve, group):
    z = golden["point_vectors"]
    key = "c%d_g%d" % (curve, group)
    w = 24 * po.degree(curve, group)
    A, B = z[key + "_a"].reshape(-1, w), z[key + "_b"].reshape(-1, w)
    for op, name in ((0, "add"), (1, "dbl"), (2, "madd"), (4, "neg"), (5, "add")):
        want = z["%s_%s_out" % (key, name)].reshape(-1, w)
        for i in range(A.shape[0]):
            assert (oracle.point_op(curve, group, op, A[i], B[i]) == want[i]).all(), (name, i)
    ks = z[key + "_k"].reshape(-1, 12)
    want = z[key + "_smul_out"].reshape(-1, w)
    bases = oracle.gen_bases(curve, group, 12).reshape(12, w)
    for i in range(ks.shape[0]):
        assert (oracle.point_op(curve, group, 3, bases[i % 12], k=ks[i]) == want[i]).all()
    J = z[key + "_jac"].reshape(7, -1)
    want = z[key + "_jac_out"].reshape(7, w)
    for i in range(7):
        assert (oracle.jacobian_to_affine(curve, group, J[i]) == want[i]).all()
    # the fold of Jacobian partials (multi-GPU host fold) agrees with repeated addition
    acc = np.zeros(w, np.uint64)
    for i in range(6):
        acc = oracle.point_op(curve, group, 0, acc, want[i])
    assert (oracle.fold_jacobian(curve, group, J[:6].reshape(-1)) == acc).all()


@pytest.mark.parametrize("curve,group", CG)
def test_msm_golden(oracle, golden, curve, group):
    z = golden["msm_vectors"]
    key = "c%d_g%d" % (curve, group)
    deg = po.degree(curve, group)
    bases, sc = z[key + "_bases"], z[key + "_scalars"]
    for n in (0, 1, 2, 31, 32, 33, 100, 257):
        want = z["%s_n%d_out" % (key, n)]
        for method, chunks, pre in ((1, 0, 1), (0, 1, 0)):
            got, _ = oracle.msm(curve, group, bases[:n * 24 * deg], sc[:n * 12], method=method, chunks=chunks, prefilter=pre)
            assert (got == want).all(), (n, method)


@pytest.mark.parametrize("curve,group", CG)
def test_closed_form(oracle, curve, group):
    n = 200
    bases = oracle.gen_bases(curve, group, n)
    s = po.gen_scalars(curve, n, 5)
    want, _ = oracle.msm(curve, group, bases, s)
    assert (oracle.msm_closed_form(curve, group, s) == want).all()


def test_scalar_generator_matches_fixture(golden):
    # gen_scalars (python restatement of libff::SHA512_rng) produced the fixture scalars with seed 11
    z = golden["msm_vectors"]
    for c in (0, 1):
        mine = po.gen_scalars(c, 10, 11)
        assert (mine == z["c%d_g1_scalars" % c][:120]).all()


@pytest.mark.parametrize("curve", [0, 1])
def test_compute_h_restatement_matches_reference_fixtures(oracle, golden, curve):
    """orc_compute_h (radix-2 FFTs restated from libfqfft) against compute_H run through the reference's own
    libfqfft domain (tools/gen_golden.py -> tests/golden/h_vectors.npz)."""
    z = golden["h_vectors"]
    for m in (2, 8, 64, 512):
        k = "c%d_m%d_" % (curve, m)
        assert (oracle.compute_h(curve, z[k + "ca"], z[k + "cb"], z[k + "cc"]) == z[k + "out"]).all(), m


def test_constant_headers_are_the_generators():
    """The committed constant headers -- the product's csrc/mnt753_constants.h (moduli, Montgomery constants, roots of
    unity, generators, twist-Frobenius and scalar-split constants of csrc/glv.cuh) and the oracle's -- are byte for
    byte what tools/gen_constants.py derives from the two primes.  Re-running the derivation re-runs its checks: the
    2-adicity and roots of unity, the reduced lattice basis of the G2 scalar split and, on 2000 random and 9 edge
    scalars, k0 + k1 (q mod r) = k (mod r) with |k0|, |k1| < 2^377."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_constants", os.path.join(root, "tools", "gen_constants.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    want = gen.emit(32, "MNT753_CONSTANTS_H", "/* 24 x 32-bit little-endian limbs, R = 2^768. */")
    assert open(os.path.join(root, "gpu_groth16_prover_3x_b200", "csrc", "mnt753_constants.h")).read() == want
    want = gen.emit(64, "ORACLE_CONSTANTS_H", "/* TEST INFRASTRUCTURE ONLY. 12 x 64-bit little-endian limbs, R = 2^768. */")
    assert open(os.path.join(root, "oracle", "oracle_constants.h")).read() == want
    for curve in (0, 1):
        g = gen.glv_params(curve)
        r = gen.MOD_B if curve == 0 else gen.MOD_A
        for k in (0, 1, r - 1, g["lam"], (1 << 752) % r):
            k0, k1 = g["split"](k)
            assert (k0 + k1 * g["lam"] - k) % r == 0 and max(abs(k0), abs(k1)).bit_length() <= 377
