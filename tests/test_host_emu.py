"""The exact field / curve templates the kernels instantiate (csrc/fq.cuh, fe.cuh, ec.cuh), compiled
for the host with the PTX carry-chain primitives emulated, checked against the oracle and the golden
fixtures.  This is what can be verified about the device code on a machine without a GPU."""
import ctypes

import numpy as np
import pytest

from oracle import pyoracle as po

CG = [(c, g) for c in (0, 1) for g in (1, 2)]
_p = po._p


def emu_field(emu, curve, group, op, a, b=None):
    out = np.zeros_like(a)
    n = a.size // (12 * po.degree(curve, group))
    assert emu.emu_field_op(curve, group, op, n, _p(a), _p(b), _p(out)) == 0
    return out


def emu_point(emu, curve, group, op, acc, q, flags=0):
    out = np.zeros_like(acc)
    ret = emu.emu_point_op(curve, group, op, _p(np.ascontiguousarray(acc)), _p(np.ascontiguousarray(q)) if q is not None else None, flags, _p(out))
    return out, ret


def to_jac(curve, group, aff):
    """affine wire point -> Jacobian with Z = 1 (infinity -> (1,1,0))."""
    deg = po.degree(curve, group)
    one = po.ints_to_array([po.R % po.fq_modulus(curve)] + [0] * (deg - 1))
    y = aff[12 * deg:]
    if not y.any():
        return np.concatenate([one, one, np.zeros(12 * deg, np.uint64)])
    return np.concatenate([aff, one])


@pytest.mark.parametrize("curve,group", CG)
def test_emu_field_golden(emu, golden, curve, group):
    z = golden["field_vectors"]
    key = "c%d_f%d" % (curve, 0 if group == 1 else 1)
    a, b = z[key + "_a"], z[key + "_b"]
    for op, name in ((0, "mul"), (1, "add"), (2, "sub"), (3, "sqr"), (5, "neg"), (7, "mul")):
        got = emu_field(emu, curve, group, op, a, b)
        assert (got == z["%s_%s_out" % (key, name)]).all(), name
    dbl = emu_field(emu, curve, group, 8, a)
    add = emu_field(emu, curve, group, 1, a, a)
    assert (dbl == add).all()


@pytest.mark.parametrize("curve,group", CG)
def test_emu_field_random(emu, oracle, curve, group):
    rng = np.random.default_rng(curve * 10 + group)
    p = po.fq_modulus(curve)
    deg = po.degree(curve, group)
    n = 300
    a = po.ints_to_array([int.from_bytes(rng.bytes(100), "little") % p for _ in range(n * deg)])
    b = po.ints_to_array([int.from_bytes(rng.bytes(100), "little") % p for _ in range(n * deg)])
    f = 0 if group == 1 else 1
    for op in (0, 1, 2, 3, 5):
        assert (emu_field(emu, curve, group, op, a, b) == oracle.field_op(curve, f, op, a, b)).all(), op


@pytest.mark.parametrize("curve", [0, 1])
def test_emu_from_mont(emu, golden, curve):
    z = golden["field_vectors"]
    s = z["c%d_fr_mont" % curve]
    out = np.zeros_like(s)
    emu.emu_fr_from_mont(curve, s.size // 12, _p(s), _p(out))
    assert (out == z["c%d_fr_plain_out" % curve]).all()


@pytest.mark.parametrize("curve,group", CG)
def test_emu_point_ops(emu, oracle, golden, curve, group):
    z = golden["point_vectors"]
    key = "c%d_g%d" % (curve, group)
    deg = po.degree(curve, group)
    w = 24 * deg
    A, B = z[key + "_a"].reshape(-1, w), z[key + "_b"].reshape(-1, w)
    add_out, dbl_out = z[key + "_add_out"].reshape(-1, w), z[key + "_dbl_out"].reshape(-1, w)
    for i in range(A.shape[0]):
        ja, jb = to_jac(curve, group, A[i]), to_jac(curve, group, B[i])
        # full Jacobian add
        out, _ = emu_point(emu, curve, group, 1, ja, jb)
        assert (oracle.jacobian_to_affine(curve, group, out) == add_out[i]).all(), ("add", i)
        # doubling
        out, _ = emu_point(emu, curve, group, 2, ja, None)
        assert (oracle.jacobian_to_affine(curve, group, out) == dbl_out[i]).all(), ("dbl", i)
        # mixed add (affine operand must not be infinity: the accumulate kernel never feeds one)
        if B[i][12 * deg:].any():
            a_inf = not A[i][12 * deg:].any()
            out, _ = emu_point(emu, curve, group, 0, ja, B[i], 2 if a_inf else 0)
            assert (oracle.jacobian_to_affine(curve, group, out) == add_out[i]).all(), ("madd", i)
            # negated operand: A - B
            negb = oracle.point_op(curve, group, 4, B[i])
            want = oracle.point_op(curve, group, 0, A[i], negb)
            out, _ = emu_point(emu, curve, group, 0, ja, B[i], (2 if a_inf else 0) | 1)
            assert (oracle.jacobian_to_affine(curve, group, out) == want).all(), ("msub", i)


@pytest.mark.parametrize("curve,group", CG)
def test_emu_point_chain_random_z(emu, oracle, curve, group):
    """Accumulate 10 points with madd, then add the accumulator to itself and to its negation:
    exercises Z != 1 operands in add / dbl and the P+P, P+(-P) branches."""
    deg = po.degree(curve, group)
    w = 24 * deg
    pts = oracle.gen_bases(curve, group, 10).reshape(10, w)
    acc = to_jac(curve, group, np.zeros(w, np.uint64))
    want = np.zeros(w, np.uint64)
    inf = 2
    for i in range(10):
        acc, r = emu_point(emu, curve, group, 0, acc, pts[i], inf)
        inf = 2 if r else 0
        want = oracle.point_op(curve, group, 0, want, pts[i])
    assert (oracle.jacobian_to_affine(curve, group, acc) == want).all()
    out, _ = emu_point(emu, curve, group, 1, acc, acc)      # P + P through the add path
    assert (oracle.jacobian_to_affine(curve, group, out) == oracle.point_op(curve, group, 1, want)).all()
    # P + (-P): negate Y of the Jacobian accumulator via the oracle field op
    f = 0 if group == 1 else 1
    neg = acc.copy()
    neg[12 * deg:24 * deg] = oracle.field_op(curve, f, 5, acc[12 * deg:24 * deg].copy())
    out, _ = emu_point(emu, curve, group, 1, acc, neg)
    assert not oracle.jacobian_to_affine(curve, group, out).any()
    # madd hitting P + P and P + (-P)
    out, r = emu_point(emu, curve, group, 0, to_jac(curve, group, pts[3]), pts[3], 0)
    assert (oracle.jacobian_to_affine(curve, group, out) == oracle.point_op(curve, group, 1, pts[3])).all() and r == 0
    out, r = emu_point(emu, curve, group, 0, to_jac(curve, group, pts[3]), pts[3], 1)
    assert r == 1 and not oracle.jacobian_to_affine(curve, group, out).any()


@pytest.mark.parametrize("modulus", [0, 1])
def test_emu_both_multipliers(emu, oracle, golden, modulus):
    """The 32-bit CIOS product (engine), the reduced-radix experiment and the FP64-pipe experiment (52-bit limbs
    split exactly with DFMA, fq_fp64.cuh) agree with the oracle on edge values and random operands."""
    z = golden["field_vectors"]
    key = "c%d_f0" % modulus
    rng = np.random.default_rng(5 + modulus)
    p = po.fq_modulus(modulus)
    a = np.concatenate([z[key + "_a"], po.ints_to_array([int.from_bytes(rng.bytes(100), "little") % p for _ in range(500)])])
    b = np.concatenate([z[key + "_b"], po.ints_to_array([int.from_bytes(rng.bytes(100), "little") % p for _ in range(500)])])
    want = oracle.field_op(modulus, 0, 0, a, b)
    for which in (0, 1, 2):
        out = np.zeros_like(a)
        emu.emu_fq_mul(modulus, which, a.size // 12, _p(a), _p(b), _p(out))
        assert (out == want).all(), which


@pytest.mark.parametrize("modulus", [0, 1])
def test_emu_fq_inverse(emu, oracle, golden, modulus):
    """Binary extended-Euclid inversion (fq.cuh fq_inv, used by the batched-affine accumulation) against
    the oracle's field inversion: edge values (0 -> 0, 1, 2, p-1, powers of two) and random operands."""
    z = golden["field_vectors"]
    key = "c%d_f0" % modulus
    rng = np.random.default_rng(11 + modulus)
    p = po.fq_modulus(modulus)
    edge = [0, 1, 2, 3, p - 1, p - 2, (p + 1) // 2, 1 << 752, (1 << 700) % p, po.R % p, po.R * po.R % p, 1 << 31, 1 << 32, (1 << 64) - 1]
    small = [pow(2, k, p) for k in range(0, 760, 37)] + [(p - pow(2, k, p)) % p for k in range(1, 760, 41)] + list(range(1, 40))
    a = np.concatenate([z[key + "_a"], po.ints_to_array(edge + small + [int.from_bytes(rng.bytes(100), "little") % p for _ in range(1500)])])
    want = oracle.field_op(modulus, 0, 4, a)
    nz = int(np.count_nonzero(a.reshape(-1, 12).any(axis=1) == 0))
    for which in (0, 1, 2):
        out = np.zeros_like(a)
        failed = emu.emu_fq_inv(modulus, which, a.size // 12, _p(a), _p(out))
        assert (out == want).all(), which
        assert failed == (nz if which else 0), which      # the fast path itself never needs the fallback (except for 0)


@pytest.mark.parametrize("modulus", [0, 1])
def test_warp_cooperative_inverse(warp_emu, oracle, modulus):
    """csrc/fq_inv_coop.cuh -- ONE inversion by the 32 lanes of a warp, limb i on lane i, the tile inversion of the
    batched-affine accumulation -- executed on a simulated warp (tests/host_emu/warp_sim.hpp: lock-step lanes, shuffles,
    ballots, and a check that the lanes never diverge) against the oracle's field inversion.  Input and output are in
    Montgomery form; 0 -> 0; the fast path never reports failure."""
    rng = np.random.default_rng(23 + modulus)
    p = po.fq_modulus(modulus)
    edge = [0, 1, 2, 3, p - 1, p - 2, (p + 1) // 2, 1 << 752, (1 << 700) % p, po.R % p, po.R * po.R % p, 1 << 31, 1 << 32, (1 << 64) - 1]
    edge += [pow(2, k, p) for k in (29, 30, 31, 59, 60, 61, 750)] + [(p - pow(2, k, p)) % p for k in (1, 30, 60)]
    a = po.ints_to_array(edge + [int.from_bytes(rng.bytes(100), "little") % p for _ in range(40)])
    out = np.zeros_like(a)
    assert warp_emu.emu_fq_inv_coop(modulus, a.size // 12, _p(a), _p(out)) == 0
    assert (out == oracle.field_op(modulus, 0, 4, a)).all()


@pytest.mark.parametrize("curve,group", CG)
def test_emu_tower_inverse(emu, oracle, golden, curve, group):
    """Team::inv_lane0: Fq by binary gcd, Fq2 / Fq3 through the norm (conjugate / Frobenius images)."""
    z = golden["field_vectors"]
    f = 0 if group == 1 else 1
    a = z["c%d_f%d_a" % (curve, f)]
    deg = po.degree(curve, group)
    a = a.reshape(-1, 12 * deg)
    a = a[a.any(axis=1)].reshape(-1)          # inverse of zero is never requested
    out = np.zeros_like(a)
    assert emu.emu_field_inv(curve, group, a.size // (12 * deg), _p(a), _p(out)) == 0
    assert (out == oracle.field_op(curve, f, 4, a)).all()


@pytest.mark.parametrize("curve,group", CG)
def test_emu_batched_affine_pair(emu, oracle, golden, curve, group):
    """pair_forward / tile_inverse / pair_backward (batch_affine.cuh) on single pairs: generic sums from
    the golden vectors, P + P (doubling through lambda = (3x^2 + a) / 2y), P + (-P) -> infinity, copy."""
    z = golden["point_vectors"]
    key = "c%d_g%d" % (curve, group)
    deg = po.degree(curve, group)
    w = 24 * deg
    A, B = z[key + "_a"].reshape(-1, w), z[key + "_b"].reshape(-1, w)
    add_out = z[key + "_add_out"].reshape(-1, w)

    def run(p1, p2, has2=1):
        out = np.zeros(w, np.uint64)
        inf = emu.emu_affine_add(curve, group, _p(p1), _p(p2 if p2 is not None else p1), has2, _p(out))
        return out, inf

    n = 0
    for i in range(A.shape[0]):
        if not A[i][12 * deg:].any() or not B[i][12 * deg:].any():
            continue                           # infinities are resolved by the plan kernel, never fed to a pair
        out, inf = run(A[i], B[i])
        if add_out[i].any():
            assert inf == 0 and (out == add_out[i]).all(), i
        else:
            assert inf == 1 and not out.any(), i
        n += 1
    assert n >= 4
    pts = oracle.gen_bases(curve, group, 6).reshape(6, w)
    for p in pts[:3]:
        out, inf = run(p, p)
        assert inf == 0 and (out == oracle.point_op(curve, group, 1, p)).all()
        out, inf = run(p, oracle.point_op(curve, group, 4, p))
        assert inf == 1 and not out.any()
        out, inf = run(p, None, has2=0)
        assert inf == 0 and (out == p).all()
    out, inf = run(pts[3], pts[4])
    assert (out == oracle.point_op(curve, group, 0, pts[3], pts[4])).all()


def test_round_planning_model_check(plan_emu):
    """csrc/ba_plan.cuh -- the cut of the sorted list into shares, the plan of every round of every share (pairs dealt
    over the lanes with shuffles, odd one carried over, infinite operands) and the formulas that recycle scratch slots
    every second round -- executed on a simulated warp, with a set of entry ids for every point and unions for additions
    (tests/host_emu/plan_emu.cpp).  Checked: no slot is rewritten while a reference to it is still to be read, no two
    shares write the same slot, every slot lies inside the region the host sized for it, the pieces of buckets cut between
    shares meet again in the fix-up, and every bucket ends as exactly the union of its entries -- also when sums cancel
    to infinity, with tiles of a single addition per lane, with more shares than entries and with giant buckets."""
    import ctypes
    rng = np.random.default_rng(7)
    cases = [(np.array([3, 0, 400, 350, 1, 0, 2]), 12), (np.array([0, 0, 777, 0]), 10), (np.array([1, 2, 0, 1]), 8),
             (np.ones(100, np.uint32), 7), (np.zeros(10, np.uint32), 4), (rng.poisson(2.5, 600), 16), (rng.poisson(20, 64), 1)]
    for _ in range(28):
        counts = rng.poisson(rng.choice([0.3, 2, 9]), int(rng.integers(1, 40)))
        for _ in range(int(rng.integers(0, 3))):
            counts[rng.integers(0, len(counts))] = rng.integers(50, 900)
        cases.append((counts, int(rng.integers(1, 40))))
    cut = 0
    for i, (counts, shares) in enumerate(cases):
        c = np.ascontiguousarray(counts, dtype=np.uint32)
        stats, err = (ctypes.c_uint64 * 3)(), ctypes.create_string_buffer(300)
        rc = plan_emu.emu_plan_check(len(c), c.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), shares, (2048, 1, 5)[i % 3], (0, 3, 0, 7)[i % 4],
                                     stats, err, 300)
        assert rc == 0, (list(c), shares, err.value.decode())
        cut += stats[1]
    assert cut > 20          # buckets cut between shares were part of it


def test_reduction_tree_model_check(plan_emu):
    """csrc/tree_plan.cuh -- the node layout of the bucket-reduction tree and, for every addition of every round, its two
    inputs and its node (TreePairs::locate / get / passthrough, the shipped code) -- executed with integers modulo
    2^61 - 1 for points (tests/host_emu/plan_emu.cpp).  Checked: every input of a round was written by an earlier round,
    every node is written once, Jacobian nodes fall inside the array the host sized, and the k + 1 terms of every set add
    up to sum_b (b + 1) B_b -- for 2 .. 2048 buckets per set, one to three sets, any split between affine and Jacobian
    rounds, full, sparse and empty bucket sets."""
    import ctypes
    rng = np.random.default_rng(3)
    mod = (1 << 61) - 1
    for k in range(1, 12):
        h = max(0, k - 5)
        for sets in (1, 3):
            for affine_rounds in sorted({0, h // 2, h}):
                for fill in (1.0, 0.5, 0.05, 0.0):
                    val = rng.integers(1, mod, sets << k, dtype=np.uint64)
                    val[rng.random(sets << k) >= fill] = 0
                    err = ctypes.create_string_buffer(300)
                    rc = plan_emu.emu_tree_check(sets, k, affine_rounds, val.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), err, 300)
                    assert rc == 0, (k, sets, affine_rounds, fill, err.value.decode())


def test_signed_digit_recoding(plan_emu):
    """csrc/recode.cuh (the digits k_count / k_scatter sort by), for every window width 2 .. 22: the W = ceil(754 / c)
    signed digits of a 753-bit scalar lie in [-2^(c-1) + 1, 2^(c-1)] and sum_w d_w 2^(c w) is the scalar again -- no
    borrow is lost beyond the top window -- and the same for the 12-limb halves of a split G2 scalar (|k| < 2^377,
    ceil(378 / c) digits, csrc/glv.cuh)."""
    import ctypes
    rng = np.random.default_rng(17)
    r = po.fr_modulus(0)
    full = [0, 1, 2, (1 << 753) - 1, r - 1, r - 2, 1 << 752, (1 << 752) - 1, int("5" * 226, 16) % (1 << 753), int("a" * 188, 16)]
    full += [int.from_bytes(rng.bytes(96), "little") % (1 << 753) for _ in range(40)]
    half = [0, 1, (1 << 377) - 1, 1 << 376, (1 << 376) - 1] + [int.from_bytes(rng.bytes(48), "little") % (1 << 377) for _ in range(40)]
    for c in range(2, 23):
        for vals, nl, bits in ((full, 24, 754), (half, 12, 378)):
            w = (bits + c - 1) // c
            for k in vals:
                limbs = np.frombuffer(k.to_bytes(4 * nl, "little"), dtype=np.uint32).copy()
                digits = np.zeros(w, np.int32)
                plan_emu.emu_recode(limbs.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), nl, c, w, digits.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)))
                assert digits.min() >= -(1 << (c - 1)) + 1 and digits.max() <= 1 << (c - 1), (c, k)
                assert sum(int(d) << (c * i) for i, d in enumerate(digits)) == k, (c, k)


@pytest.mark.parametrize("curve", [0, 1])
def test_g2_scalar_split(plan_emu, curve):
    """csrc/glv_split.cuh -- the arithmetic of k_glv_split, one G2 scalar per thread: k = k0 + k1 (q mod r) modulo r,
    |k0|, |k1| < 2^377, stored as 12-limb magnitudes with the sign in bit 383 -- for the edges of the decomposition
    (0, 1, r - 1, lambda and its neighbours, multiples of lambda) and random scalars; lambda from the derivation in
    tools/gen_constants.py."""
    import ctypes
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_constants", os.path.join(root, "tools", "gen_constants.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    g = gen.glv_params(curve)
    lam, r = g["lam"], po.fr_modulus(curve)
    rng = np.random.default_rng(29 + curve)
    vals = [0, 1, 2, r - 1, r - 2, lam, lam - 1, lam + 1, (2 * lam) % r, (lam * lam) % r, r // 2, (r - lam) % r, 1 << 752, (1 << 376), (1 << 377) - 1]
    vals += [int.from_bytes(rng.bytes(100), "little") % r for _ in range(500)]
    u32p = ctypes.POINTER(ctypes.c_uint32)
    for k in vals:
        limbs = np.frombuffer(k.to_bytes(96, "little"), dtype=np.uint32).copy()
        out = np.zeros(25, np.uint32)
        plan_emu.emu_glv_split(curve, limbs.ctypes.data_as(u32p), out.ctypes.data_as(u32p))
        assert out[24] == 0, k
        halves = []
        for h in range(2):
            mag = int.from_bytes(out[12 * h:12 * h + 12].tobytes(), "little")
            sign = mag >> 383
            mag &= (1 << 383) - 1
            assert mag.bit_length() <= 377, k
            halves.append(-mag if sign else mag)
        assert (halves[0] + halves[1] * lam - k) % r == 0, k
        assert halves == list(g["split"](k)), k          # and it is the split the constants were derived with


@pytest.mark.parametrize("curve", [0, 1])
def test_h_polynomial_kernels_on_a_simulated_block(fft_emu, oracle, golden, curve):
    """csrc/fft_kernels.cuh -- domain constants, power tables, the NTT kernels (one stage per launch, and the tiled
    kernel that runs several stages of a tile in shared memory between __syncthreads()), the pointwise steps and the
    bit-reversing tail -- executed on a simulated thread block (tests/host_emu/block_sim.hpp) in the launch sequence of
    csrc/fft.cu, against the oracle's restatement of libfqfft's compute_H and against the fixture libfqfft itself
    produced: sizes 1 .. 1024, tiles cut into one, two and three groups of index bits."""
    for logm, tiles in ((0, (0,)), (1, (0,)), (3, (0, 2)), (6, (0, 3, 10)), (8, (4, 10)), (10, (10,))):
        m = 1 << logm
        ca, cb, cc = (po.gen_scalars(curve, m, 40 + 3 * logm + i) for i in range(3))
        want = oracle.compute_h(curve, ca, cb, cc)
        for tile_bits in tiles:
            out = np.zeros((m + 1) * 12, np.uint64)
            assert fft_emu.emu_compute_h(curve, logm, tile_bits, _p(ca), _p(cb), _p(cc), _p(out)) == 0
            assert (out == want).all(), (logm, tile_bits)
    z = golden["h_vectors"]               # produced by the reference's libfqfft (tools/gen_golden.py)
    for m in (2, 8, 64, 512):
        k = "c%d_m%d_" % (curve, m)
        out = np.zeros((m + 1) * 12, np.uint64)
        logm = m.bit_length() - 1
        assert fft_emu.emu_compute_h(curve, logm, 0 if logm < 6 else 5, _p(z[k + "ca"]), _p(z[k + "cb"]), _p(z[k + "cc"]), _p(out)) == 0
        assert (out == z[k + "out"]).all(), m


def _signed_digits(k, c, w):
    """The recoding of csrc/recode.cuh restated: digits in [-2^(c-1) + 1, 2^(c-1)], a borrow carried upwards."""
    out, carry = [], 0
    for i in range(w):
        raw = ((k >> (c * i)) & ((1 << c) - 1)) + carry
        if raw > 1 << (c - 1):
            out.append(raw - (1 << c)); carry = 1
        else:
            out.append(raw); carry = 0
    return out


@pytest.mark.parametrize("c,tables,glv", [(3, 1, 0), (5, 4, 0), (8, 95, 0), (11, 7, 0), (6, 8, 1), (9, 84, 1)])
def test_counting_sort_on_a_simulated_block(sort_emu, c, tables, glv):
    """csrc/sort_kernels.cuh -- k_count, the exclusive scan in three kernels, k_scatter -- on a simulated thread block in
    the launch sequence of enqueue_msm: the histogram counts every non-zero digit of every finite base once, the offsets
    are its exclusive prefix sums, and the entry list holds, bucket by bucket, exactly the (table row | sign) of those
    digits -- digit w of a scalar in bucket set w mod G through table w div G; for split G2 scalars the digits of the
    two signed halves."""
    import ctypes
    rng = np.random.default_rng(100 + c)
    n = 700                                              # three blocks of k_count, the last one ragged
    bits = 378 if glv else 754
    wh = (bits + c - 1) // c
    wd = 2 * wh if glv else wh
    sets = (wd + tables - 1) // tables
    if glv:
        wh = (wh + sets - 1) // sets * sets              # a table belongs to one half (choose_cfg, host_ctx.cuh)
        wd = 2 * wh
    nb = 1 << (c - 1)
    inf = (rng.random(n) < 0.05).astype(np.uint8)
    scal = np.zeros((n, 24), np.uint32)
    want = {}
    for i in range(n):
        if glv:
            halves = [int.from_bytes(rng.bytes(48), "little") % (1 << 377) * int(rng.integers(0, 2) * 2 - 1) for _ in range(2)]
            if i < 4:
                halves = [(0, 0), ((1 << 377) - 1, -1), (-((1 << 377) - 1), 1), (1 << 376, 0)][i]
                halves = list(halves)
            digits = []
            for h, kh in enumerate(halves):
                mag = abs(kh) | ((1 << 383) if kh < 0 else 0)
                scal[i, 12 * h:12 * h + 12] = np.frombuffer(mag.to_bytes(48, "little"), dtype=np.uint32)
                digits += [(-d if kh < 0 else d) for d in _signed_digits(abs(kh), c, wh)]
        else:
            k = int.from_bytes(rng.bytes(96), "little") % (1 << 753) if i >= 3 else (0, (1 << 753) - 1, 1)[i]
            scal[i] = np.frombuffer(k.to_bytes(96, "little"), dtype=np.uint32)
            digits = _signed_digits(k, c, wd)
        if inf[i]:
            continue
        for w, d in enumerate(digits):
            if d:
                want.setdefault((w % sets) * nb + abs(d) - 1, []).append(((w // sets) * n + i) | (0x80000000 if d < 0 else 0))
    K = sets * nb
    total = sum(len(v) for v in want.values())
    count, offs, entries = np.zeros(K, np.uint32), np.zeros(K + 1, np.uint32), np.zeros(total + 8, np.uint32)
    u32p = ctypes.POINTER(ctypes.c_uint32)
    assert sort_emu.emu_sort(n, c, wd, sets, n, glv, wh, scal.ctypes.data_as(u32p), inf.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                             count.ctypes.data_as(u32p), offs.ctypes.data_as(u32p), entries.ctypes.data_as(u32p)) == 0
    assert offs[K] == total and offs[0] == 0
    assert (np.diff(offs.astype(np.int64)) == count).all()
    for b in range(K):
        assert sorted(entries[offs[b]:offs[b + 1]].tolist()) == sorted(want.get(b, [])), b


@pytest.mark.parametrize("curve,group", CG)
def test_emu_generic_tile(emu, oracle, curve, group):
    """The UNCLASSIFIED passes of a tile -- the hot path of the accumulation (batch_affine.cuh: pair_generic_forward,
    one inversion for the whole batch, pair_generic_backward; for Fq the differences are formed inside the product,
    Team::mulsub) -- on batches of 1, 2 and 9 generic additions with every combination of negated operands, against
    the oracle's point addition; and a batch that holds a doubling reports a zero product (the kernel's cue to run the
    classified passes)."""
    import ctypes
    deg = po.degree(curve, group)
    w = 24 * deg
    pts = oracle.gen_bases(curve, group, 24).reshape(-1, w)
    for batch in (1, 2, 9):
        p1 = np.ascontiguousarray(pts[:batch])
        p2 = np.ascontiguousarray(pts[12:12 + batch])
        flags = np.array([(i * 7 + batch) % 4 for i in range(batch)], np.int32)
        out = np.zeros_like(p1)
        assert emu.emu_batch_add_generic(curve, group, batch, _p(p1), _p(p2), flags.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _p(out)) == 0
        for i in range(batch):
            a = oracle.point_op(curve, group, 4, p1[i]) if flags[i] & 1 else p1[i]
            b = oracle.point_op(curve, group, 4, p2[i]) if flags[i] & 2 else p2[i]
            assert (out[i] == oracle.point_op(curve, group, 0, a, b)).all(), (batch, i)
    p1 = np.ascontiguousarray(pts[:3])
    p2 = np.ascontiguousarray(pts[[5, 1, 7]])                # the middle pair is P + P
    out = np.zeros_like(p1)
    flags = np.zeros(3, np.int32)
    assert emu.emu_batch_add_generic(curve, group, 3, _p(p1), _p(p2), flags.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), _p(out)) == 1


@pytest.mark.parametrize("curve", [0, 1])
def test_emu_twist_frobenius(emu, oracle, curve):
    """psi(x, y) = (cX Frob(x), cY Frob(y)) -- the map k_psi_many derives the upper half of the G2 window tables with
    (csrc/glv.cuh; libff G2::mul_by_q, mnt4753_g2.cpp:364-368) -- is multiplication by lambda = q mod r: the steps of the
    kernel on the host against the oracle's scalar multiplication, and psi of infinity (all zero) stays infinity."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_constants", os.path.join(root, "tools", "gen_constants.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    lam, r = gen.glv_params(curve)["lam"], po.fr_modulus(curve)
    k = po.int_to_limbs(lam * po.R % r)
    w = 24 * po.degree(curve, 2)
    for p in oracle.gen_bases(curve, 2, 6).reshape(-1, w):
        out = np.zeros(w, np.uint64)
        assert emu.emu_psi(curve, _p(p), _p(out)) == 0
        assert (out == oracle.point_op(curve, 2, 3, p, k=k)).all()
    out = np.ones(w, np.uint64)
    assert emu.emu_psi(curve, _p(np.zeros(w, np.uint64)), _p(out)) == 0 and not out.any()
