"""GPU parity of the device field / point layer (through the C-ABI self-test hooks) against the
oracle and the golden fixtures.  Bit-exact: all arithmetic is integer."""
import numpy as np
import pytest

import gpu_groth16_prover_3x_b200 as pkg
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
CG = [(c, g) for c in (0, 1) for g in (1, 2)]


@pytest.fixture(scope="module")
def ctxs():
    d = {c: pkg.MsmContext(c, 0) for c in (0, 1)}
    yield d
    for c in d.values():
        c.close()


def to_jac(curve, group, aff):
    deg = po.degree(curve, group)
    one = po.ints_to_array([po.R % po.fq_modulus(curve)] + [0] * (deg - 1))
    if not aff[12 * deg:].any():
        return np.concatenate([one, one, np.zeros(12 * deg, np.uint64)])
    return np.concatenate([aff, one])


@pytest.mark.parametrize("curve,group", CG)
def test_field_golden(ctxs, golden, curve, group):
    z = golden["field_vectors"]
    key = "c%d_f%d" % (curve, 0 if group == 1 else 1)
    a, b = z[key + "_a"], z[key + "_b"]
    for op, name in ((0, "mul"), (1, "add"), (2, "sub"), (3, "sqr"), (5, "neg"), (7, "mul")):
        got = ctxs[curve].selftest_field(group, op, a, b)
        assert (got == z["%s_%s_out" % (key, name)]).all(), name


@pytest.mark.parametrize("curve,group", CG)
def test_field_random(ctxs, oracle, curve, group):
    rng = np.random.default_rng(100 + curve * 10 + group)
    p = po.fq_modulus(curve)
    deg = po.degree(curve, group)
    n = 5000  # not a multiple of the 32-lane team width
    a = po.ints_to_array([int.from_bytes(rng.bytes(100), "little") % p for _ in range(n * deg)])
    b = po.ints_to_array([int.from_bytes(rng.bytes(100), "little") % p for _ in range(n * deg)])
    f = 0 if group == 1 else 1
    for op in (0, 1, 2, 3, 5):
        assert (ctxs[curve].selftest_field(group, op, a, b) == oracle.field_op(curve, f, op, a, b)).all(), op


@pytest.mark.parametrize("curve,group", CG)
def test_tile_inversion(ctxs, oracle, curve, group):
    """Team::inv_lane0 on the device -- the warp-cooperative inversion of csrc/fq_inv_coop.cuh (Fq2 / Fq3 through the
    norm) -- against the oracle's field inversion: edge values (1, 2, p - 1, powers of two, small and huge) and random."""
    rng = np.random.default_rng(7 + curve * 10 + group)
    p = po.fq_modulus(curve)
    deg = po.degree(curve, group)
    edge = [1, 2, 3, p - 1, p - 2, (p - 1) // 2, (p + 1) // 2, po.R % p, (1 << 752) % p, (1 << 32) - 1, 1 << 32, (1 << 64) + 1]
    edge += [(1 << k) % p for k in range(1, 768, 37)]
    vals = []
    for e in edge:
        vals += [e] + [0] * (deg - 1)                  # elements of the base field
    for e in edge[:8]:
        vals += [0] * (deg - 1) + [e]                  # pure top coefficient
    n_rand = 700
    vals += [int.from_bytes(rng.bytes(100), "little") % p for _ in range(n_rand * deg)]
    a = po.ints_to_array(vals)
    f = 0 if group == 1 else 1
    got = ctxs[curve].selftest_field(group, 8, a)
    assert (got == oracle.field_op(curve, f, 4, a)).all()


@pytest.mark.parametrize("curve,group", CG)
def test_point_ops(ctxs, oracle, golden, curve, group):
    z = golden["point_vectors"]
    key = "c%d_g%d" % (curve, group)
    deg = po.degree(curve, group)
    w = 24 * deg
    A, B = z[key + "_a"].reshape(-1, w), z[key + "_b"].reshape(-1, w)
    n = A.shape[0]
    add_out, dbl_out = z[key + "_add_out"].reshape(-1, w), z[key + "_dbl_out"].reshape(-1, w)
    JA = np.concatenate([to_jac(curve, group, A[i]) for i in range(n)])
    JB = np.concatenate([to_jac(curve, group, B[i]) for i in range(n)])
    ctx = ctxs[curve]
    out = ctx.selftest_point(group, 1, JA, JB).reshape(n, -1)
    for i in range(n):
        assert (oracle.jacobian_to_affine(curve, group, out[i]) == add_out[i]).all(), ("add", i)
    out = ctx.selftest_point(group, 2, JA).reshape(n, -1)
    for i in range(n):
        assert (oracle.jacobian_to_affine(curve, group, out[i]) == dbl_out[i]).all(), ("dbl", i)
    # mixed add, on the pairs whose affine operand is finite (the engine filters infinite bases)
    idx = [i for i in range(n) if B[i][12 * deg:].any()]
    acc = np.concatenate([to_jac(curve, group, A[i]) for i in idx])
    q = np.concatenate([B[i] for i in idx])
    for neg in (0, 1):
        flags = np.array([(0 if A[i][12 * deg:].any() else 2) | neg for i in idx], np.uint32)
        out = ctx.selftest_point(group, 0, acc, q, flags).reshape(len(idx), -1)
        for k, i in enumerate(idx):
            bb = oracle.point_op(curve, group, 4, B[i]) if neg else B[i]
            want = oracle.point_op(curve, group, 0, A[i], bb)
            assert (oracle.jacobian_to_affine(curve, group, out[k]) == want).all(), ("madd", neg, i)


@pytest.mark.parametrize("curve,group", CG)
def test_point_chain(ctxs, oracle, curve, group):
    """40 lanes, each a different running sum with Z != 1; then acc + acc and acc - acc through add."""
    deg = po.degree(curve, group)
    w = 24 * deg
    n, steps = 40, 6
    pts = oracle.gen_bases(curve, group, n * steps).reshape(steps, n, w)
    ctx = ctxs[curve]
    acc = np.concatenate([to_jac(curve, group, np.zeros(w, np.uint64))] * n)
    flags = np.full(n, 2, np.uint32)
    want = [np.zeros(w, np.uint64) for _ in range(n)]
    for s in range(steps):
        acc = ctx.selftest_point(group, 0, acc, pts[s].reshape(-1), flags)
        flags = np.zeros(n, np.uint32)
        for i in range(n):
            want[i] = oracle.point_op(curve, group, 0, want[i], pts[s][i])
    accr = acc.reshape(n, -1)
    for i in range(n):
        assert (oracle.jacobian_to_affine(curve, group, accr[i]) == want[i]).all()
    dbl = ctx.selftest_point(group, 1, acc, acc).reshape(n, -1)
    for i in range(0, n, 7):
        assert (oracle.jacobian_to_affine(curve, group, dbl[i]) == oracle.point_op(curve, group, 1, want[i])).all()
    f = 0 if group == 1 else 1
    neg = accr.copy()
    for i in range(n):
        neg[i][12 * deg:24 * deg] = oracle.field_op(curve, f, 5, accr[i][12 * deg:24 * deg].copy())
    zero = ctx.selftest_point(group, 1, acc, neg.reshape(-1)).reshape(n, -1)
    for i in range(n):
        assert not zero[i][24 * deg:].any()
