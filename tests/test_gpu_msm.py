"""GPU parity of the MSM engine, through the C ABI, against (1) the golden vectors produced by the
reference's libff, (2) the oracle on seeded inputs, (3) size-independent properties at full size."""
import numpy as np
import pytest

import gpu_groth16_prover_3x_b200 as pkg
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu
CG = [(c, g) for c in (0, 1) for g in (1, 2)]
SIZES = (0, 1, 2, 31, 32, 33, 100, 257)


@pytest.fixture(scope="module")
def ctxs():
    d = {c: pkg.MsmContext(c, 0) for c in (0, 1)}
    yield d
    for c in d.values():
        c.close()


def affine(oracle, curve, group, xyz):
    return oracle.jacobian_to_affine(curve, group, xyz)


@pytest.mark.parametrize("curve,group", CG)
def test_golden_vectors(ctxs, oracle, golden, curve, group):
    """Edge cases planted in the fixtures: infinity bases, duplicate bases with equal scalars (P+P in a
    bucket), P and -P with equal scalars, scalars 0, 1, 2, r-1, 2^15, 2^16-1, 2^16, 2^752."""
    z = golden["msm_vectors"]
    key = "c%d_g%d" % (curve, group)
    deg = po.degree(curve, group)
    bases, sc = z[key + "_bases"], z[key + "_scalars"]
    ctx = ctxs[curve]
    slot = ctx.upload_bases(group, bases)
    try:
        for c in (0, 2, 5, 8, 13, 16):
            ctx.set_window_bits(c)
            for n in SIZES:
                got = affine(oracle, curve, group, ctx.msm(slot, sc[:n * 12], n))
                assert (got == z["%s_n%d_out" % (key, n)]).all(), (c, n)
        ctx.set_window_bits(0)
        # sub-ranges of a resident base set (the L query uses scalars w[2..], cuda_prover_piecewise.cu:167)
        for off, n in ((1, 32), (5, 1), (40, 4), (200, 57)):
            want, _ = oracle.msm(curve, group, bases[off * 24 * deg:(off + n) * 24 * deg], sc[off * 12:(off + n) * 12])
            got = affine(oracle, curve, group, ctx.msm(slot, sc[off * 12:(off + n) * 12], n, offset=off))
            assert (got == want).all(), (off, n)
    finally:
        ctx.set_window_bits(0)
        ctx.free_bases(slot)
    # literal ec_reduce form: bases travel with the call
    got = affine(oracle, curve, group, ctx.ec_reduce(group, bases[:100 * 24 * deg], sc[:1200]))
    assert (got == z["%s_n100_out" % key]).all()


@pytest.mark.parametrize("curve,group,log_n", [(0, 1, 14), (0, 2, 12), (1, 1, 14), (1, 2, 11)])
def test_seeded_vs_oracle(ctxs, oracle, curve, group, log_n):
    n = (1 << log_n) + 1
    bases = oracle.gen_bases(curve, group, n)
    sc = po.gen_scalars(curve, n, 21 + log_n)
    want, _ = oracle.msm(curve, group, bases, sc)
    got = affine(oracle, curve, group, ctxs[curve].ec_reduce(group, bases, sc))
    assert (got == want).all()
    assert (oracle.msm_closed_form(curve, group, sc) == want).all()


@pytest.mark.parametrize("curve,group", CG)
def test_degenerate_inputs(ctxs, oracle, curve, group):
    """All-identical bases (multiexp_profile.cpp:24-30 style), all-equal scalars (one giant bucket per
    window), all-zero scalars, all-infinity bases."""
    deg = po.degree(curve, group)
    w = 24 * deg
    n = 3000
    r = po.fr_modulus(curve)
    ctx = ctxs[curve]
    one_pt = oracle.gen_bases(curve, group, 1)
    same = np.tile(one_pt, n)
    sc = po.gen_scalars(curve, n, 77)
    ssum = sum(po.limbs_to_int(x) * pow(po.R, -1, r) % r for x in oracle.fr_to_mont(curve, oracle.fr_from_mont(curve, sc)).reshape(n, 12)) % r
    want = oracle.point_op(curve, group, 3, one_pt, k=po.int_to_limbs(ssum * po.R % r))
    assert (affine(oracle, curve, group, ctx.ec_reduce(group, same, sc)) == want).all()
    # equal scalars on distinct bases: every entry of a window lands in the same bucket
    bases = oracle.gen_bases(curve, group, n)
    k = 0x1234567 * (1 << 700) + 0xDEADBEEFCAFE
    eq = np.tile(po.int_to_limbs(k * po.R % r), n)
    want, _ = oracle.msm(curve, group, bases, eq)
    assert (affine(oracle, curve, group, ctx.ec_reduce(group, bases, eq)) == want).all()
    # all-zero scalars and all-infinity bases give infinity, reported as Z == 0
    out = ctx.ec_reduce(group, bases, np.zeros(n * 12, np.uint64))
    assert not out[24 * deg:].any() and not affine(oracle, curve, group, out).any()
    out = ctx.ec_reduce(group, np.zeros(n * w, np.uint64), sc)
    assert not out[24 * deg:].any()


@pytest.mark.parametrize("curve,group", CG)
def test_window_tables(ctxs, oracle, golden, curve, group):
    """Window tables 2^(c*G*t) * P_i built at upload time (include/b200_msm.h): full set (one bucket set),
    partial set under a byte budget (1 < G < windows), and none; infinity bases stay infinity in every
    table; sub-ranges index all tables with the same offset."""
    z = golden["msm_vectors"]
    key = "c%d_g%d" % (curve, group)
    deg = po.degree(curve, group)
    bases, sc = z[key + "_bases"], z[key + "_scalars"]
    nb = len(bases) // (24 * deg)
    assert nb >= 257
    ctx = ctxs[curve]
    point_bytes = 2 * deg * 96
    try:
        for c, budget, want_sets in ((0, 1 << 40, "one"), (7, 5 * nb * point_bytes, "some"), (9, 0, "all"), (11, 1 << 40, "one")):
            ctx.set_window_bits(c)
            ctx.set_table_budget(budget)
            slot = ctx.upload_bases(group, bases)
            info = ctx.bases_info(slot)
            if want_sets == "one":
                assert info["bucket_sets"] == 1 and info["tables"] > 1
            elif want_sets == "some":
                # G2 scalars are split in two halves (csrc/glv.cuh), each with its own tables: an even number of them
                assert info["tables"] == (5 if group == 1 else 4) and 1 < info["bucket_sets"] < 108
            else:
                assert info["tables"] == 1
            for n in (257, 100, 1):
                got = affine(oracle, curve, group, ctx.msm(slot, sc[:n * 12], n))
                assert (got == z["%s_n%d_out" % (key, n)]).all(), (c, budget, n)
                t = ctx.last_timings()
                assert t["tables"] == info["tables"]
            off, n = 3, 200
            want, _ = oracle.msm(curve, group, bases[off * 24 * deg:(off + n) * 24 * deg], sc[off * 12:(off + n) * 12])
            got = affine(oracle, curve, group, ctx.msm(slot, sc[off * 12:(off + n) * 12], n, offset=off))
            assert (got == want).all(), (c, budget)
            ctx.free_bases(slot)
    finally:
        ctx.set_window_bits(0)
        ctx.set_table_budget(32 << 30)


@pytest.mark.parametrize("curve", [0, 1])
def test_g2_split_scalars(ctxs, oracle, curve):
    """G2 with window tables: scalars are split k = k0 + k1 (q mod r) on the device and the upper half of the tables is
    psi of the lower half (csrc/glv.cuh).  Scalars around the places where the split changes shape -- 0, 1, r - 1, the
    eigenvalue lam = q mod r and its neighbours, multiples of lam, -lam, values with k0 or k1 zero or negative -- on
    single points and mixed into a random MSM, against the oracle's plain double-and-add / BDLO12."""
    group = 2
    r, q = po.fr_modulus(curve), po.fq_modulus(curve)
    lam = q % r
    deg = po.degree(curve, group)
    n = 300                                     # >= 256: tables are built, so the split path runs
    bases = oracle.gen_bases(curve, group, n)
    special = [0, 1, 2, r - 1, r - 2, lam, lam - 1, lam + 1, (r - lam) % r, (2 * lam) % r, (lam * lam) % r, (lam * (lam - 1)) % r,
               (1 << 376) % r, (1 << 377) % r, ((1 << 377) - 1) % r, (lam << 1) % r, (r - 1) // 2, (r + 1) // 2, pow(lam, -1, r)]
    rng = np.random.default_rng(5 + curve)
    ks = special + [int.from_bytes(rng.bytes(100), "little") % r for _ in range(n - len(special))]
    sc = po.ints_to_array([k * po.R % r for k in ks])
    ctx = ctxs[curve]
    slot = ctx.upload_bases(group, bases)
    try:
        info = ctx.bases_info(slot)
        assert info["tables"] > 1 and info["tables"] % 2 == 0
        want, _ = oracle.msm(curve, group, bases, sc)
        assert (affine(oracle, curve, group, ctx.msm(slot, sc, n)) == want).all()
        for i in range(len(special)):           # one point at a time: k_i * P_i
            want = oracle.point_op(curve, group, 3, bases[i * 24 * deg:(i + 1) * 24 * deg], k=sc[i * 12:(i + 1) * 12])
            got = affine(oracle, curve, group, ctx.msm(slot, sc[i * 12:(i + 1) * 12], 1, offset=i))
            assert (got == want).all(), i
    finally:
        ctx.free_bases(slot)


@pytest.mark.parametrize("curve,group", CG)
def test_giant_buckets_are_split_and_fixed_up(ctxs, oracle, curve, group):
    """The accumulation cuts the sorted list into shares at bucket boundaries, except for a bucket larger than
    half a share, whose pieces are summed by k_ba_fixup.  Scalars with few distinct values put thousands of points
    into one bucket per window (as the 0/1 wires of a real witness do): equal scalars (one giant bucket per window,
    repeated points doubling inside it), two values, and P / -P pairs that cancel inside a giant bucket."""
    n = 3000
    deg = po.degree(curve, group)
    bases = oracle.gen_bases(curve, group, n)
    sc = po.gen_scalars(curve, n, 11).reshape(n, 12)
    ctx = ctxs[curve]
    cases = {}
    same = sc.copy(); same[:] = sc[0]
    cases["equal"] = (bases, same.reshape(-1))
    two = sc.copy(); two[::2] = sc[1]; two[1::2] = sc[2]
    cases["two values"] = (bases, two.reshape(-1))
    mixed = sc.copy(); mixed[: n // 2] = sc[3]
    cases["half equal"] = (bases, mixed.reshape(-1))
    # every base twice in a row: (P_i, P_i) doubles inside a bucket
    dup = bases.reshape(n, 24 * deg).copy(); dup[1::2] = dup[0::2]
    cases["repeated points"] = (dup.reshape(-1), same.reshape(-1))
    # P_i followed by -P_i with the same scalar: the pair cancels to infinity
    neg = bases.reshape(n, 24 * deg).copy()
    for i in range(0, n - 1, 2):
        y = neg[i, 12 * deg:].copy()
        neg[i + 1, :12 * deg] = neg[i, :12 * deg]
        neg[i + 1, 12 * deg:] = oracle.field_op(curve, 0 if group == 1 else 1, 5, y)
    cases["cancelling pairs"] = (neg.reshape(-1), same.reshape(-1))
    for c in (0, 9):
        ctx.set_window_bits(c)
        try:
            for name, (b, s) in cases.items():
                slot = ctx.upload_bases(group, b)
                got = affine(oracle, curve, group, ctx.msm(slot, s, n))
                want, _ = oracle.msm(curve, group, b, s)
                assert (got == want).all(), (name, c)
                assert ctx.last_rounds()["max_bucket_occupancy"] >= 256, name     # larger than half a share: cut and fixed up
                ctx.free_bases(slot)
        finally:
            ctx.set_window_bits(0)


def test_async_lanes_and_shards(ctxs, oracle):
    """Four MSMs in flight on four lanes (A, B1, L on G1; B2 on G2), then the point-range sharding
    identity: the fold of per-shard partials equals the unsharded MSM."""
    curve = 0
    ctx = ctxs[curve]
    n = 2000
    b1 = oracle.gen_bases(curve, 1, n)
    b2 = oracle.gen_bases(curve, 2, n)
    sc = po.gen_scalars(curve, n, 5)
    s1, s2 = ctx.upload_bases(1, b1), ctx.upload_bases(2, b2)
    ctx.msm_async(0, s1, sc)
    ctx.msm_async(1, s2, sc)
    ctx.msm_async(2, s1, sc[24:], n - 2, offset=2)
    ctx.msm_async(3, s1, sc[:12 * 100], 100)
    outs = [ctx.wait(i) for i in range(4)]
    assert (affine(oracle, curve, 1, outs[0]) == oracle.msm(curve, 1, b1, sc)[0]).all()
    assert (affine(oracle, curve, 2, outs[1]) == oracle.msm(curve, 2, b2, sc)[0]).all()
    assert (affine(oracle, curve, 1, outs[2]) == oracle.msm(curve, 1, b1[48:], sc[24:])[0]).all()
    assert (affine(oracle, curve, 1, outs[3]) == oracle.msm(curve, 1, b1[:2400], sc[:1200])[0]).all()
    for parts in (2, 3, 8):
        partials = [ctx.msm(s2, sc[off * 12:(off + ln) * 12], ln, offset=off) for off, ln in pkg.shard_ranges(n, parts)]
        folded = ctx.fold(2, np.concatenate(partials))
        assert (affine(oracle, curve, 2, folded) == affine(oracle, curve, 2, outs[1])).all()
        assert (oracle.fold_jacobian(curve, 2, np.concatenate(partials)) == affine(oracle, curve, 2, outs[1])).all()
    ctx.free_bases(s1)
    ctx.free_bases(s2)


def test_lanes_side_by_side(ctxs, oracle):
    """b200msm_set_lane_sms: four MSMs on four lanes, each confined to its own part of the GPU (the split
    b200msm_prove gives the witness MSMs of a small proof): the persistent grids and the number of shares change,
    the results do not -- neither for generic scalars nor with giant buckets that are cut between shares."""
    curve = 0
    ctx = ctxs[curve]
    n = 1 << 13
    b1 = oracle.gen_bases(curve, 1, n)
    b2 = oracle.gen_bases(curve, 2, n)
    sc = po.gen_scalars(curve, n, 9)
    skew = sc.copy().reshape(n, 12)
    skew[: n // 2] = skew[0]                       # half of the points in the same bucket of every window
    skew = skew.reshape(-1)
    s1, s2 = ctx.upload_bases(1, b1), ctx.upload_bases(2, b2)
    try:
        want = [ctx.msm(s1, sc), ctx.msm(s2, sc), ctx.msm(s1, skew), ctx.msm(s2, skew)]
        for split in ((20, 20, 60, 20), (1, 3, 7, 2), (0, 148, 1000, -1)):
            for lane, sms in enumerate(split):
                ctx.set_lane_sms(lane, sms)
            ctx.msm_async(0, s1, sc)
            ctx.msm_async(1, s2, sc)
            ctx.msm_async(2, s1, skew)
            ctx.msm_async(3, s2, skew)
            got = [ctx.wait(i) for i in range(4)]
            for i, (g, w) in enumerate(zip(got, want)):
                grp = 1 if i in (0, 2) else 2
                assert (affine(oracle, curve, grp, g) == affine(oracle, curve, grp, w)).all(), (split, i)
        assert (affine(oracle, curve, 1, want[0]) == oracle.msm(curve, 1, b1, sc)[0]).all()
        with pytest.raises(pkg.MsmError):
            ctx.set_lane_sms(5, 10)
    finally:
        for lane in range(4):
            ctx.set_lane_sms(lane, 0)
        ctx.free_bases(s1)
        ctx.free_bases(s2)


def test_errors(ctxs):
    ctx = ctxs[0]
    with pytest.raises(pkg.MsmError):
        ctx.set_window_bits(25)
    slot = ctx.upload_bases(1, np.zeros(24 * 4, np.uint64))
    with pytest.raises(pkg.MsmError):
        ctx.msm(slot, np.zeros(12 * 8, np.uint64), 8)  # more scalars than bases
    ctx.free_bases(slot)
    with pytest.raises(pkg.MsmError):
        ctx.free_bases(slot)


@pytest.mark.parametrize("curve,group,log_n", [(0, 1, 20), (1, 2, 17)])
def test_full_size_properties(ctxs, oracle, curve, group, log_n):
    """BASELINE sizes, where the CPU MSM takes minutes: closed form of the structured bases
    (sum s_i (P0 + i Q) = (sum s_i) P0 + (sum i s_i) Q), linearity in the scalars, and independence of
    the result from the window width."""
    import torch
    n = 1 << log_n
    deg = po.degree(curve, group)
    ctx = ctxs[curve]
    p0, q = oracle.base_pair(curve, group)
    bases = np.zeros(n * 24 * deg, np.uint64)
    assert oracle._f("gen_bases")(curve, group, n, po._p(p0), po._p(q), po._p(bases)) == 0
    # cheap 751-bit scalars (SHA512_rng in python would take minutes at this size)
    rng = np.random.default_rng(7)
    raw = rng.integers(0, 1 << 63, size=(n, 12), dtype=np.uint64)
    raw[:, 11] &= np.uint64((1 << 47) - 1)  # < 2^751 < r: canonical Montgomery residues of *some* scalars
    vals = raw.reshape(-1)
    slot = ctx.upload_bases(group, bases)
    want = oracle.msm_closed_form(curve, group, vals)
    got = ctx.msm(slot, vals)
    assert (affine(oracle, curve, group, got) == want).all()
    t = ctx.last_timings()
    assert t["kernel_launches"] > 0 and t["accumulate"] > 0
    # window-width independence
    ctx.set_window_bits(13)
    assert (affine(oracle, curve, group, ctx.msm(slot, vals)) == want).all()
    ctx.set_window_bits(0)
    # linearity: MSM(a) + MSM(b) == MSM(a + b) with a + b reduced mod r (Montgomery form is linear)
    half = n // 2
    a, b = vals[:half * 12], vals[half * 12:half * 24]
    ab = oracle.field_op(1 - curve, 0, 1, a.copy(), b.copy())  # Fr(curve) = Fq(other curve)
    s = ctx.fold(group, np.concatenate([ctx.msm(slot, a, half), ctx.msm(slot, b, half)]))
    assert (affine(oracle, curve, group, s) == affine(oracle, curve, group, ctx.msm(slot, ab, half))).all()
    # scalars already resident in HBM (device pointer) give the same point
    dev = torch.from_numpy(vals.view(np.int64)).cuda()
    assert (affine(oracle, curve, group, ctx.msm(slot, dev, n)) == want).all()
    ctx.free_bases(slot)


@pytest.mark.parametrize("curve,group,log_n", [(0, 1, 18), (1, 2, 14)])
def test_skewed_witness_distribution(ctxs, oracle, curve, group, log_n):
    """Real witnesses are not uniform: half of the scalars 0, a quarter 1, an eighth 2, a few r-1 and small
    values, the rest random.  All the ones land in ONE bucket (a 2^(log_n-2)-point bucket needs log_n-2 rounds of
    pairwise additions); the result is checked with the closed form of the structured base set."""
    n = 1 << log_n
    r = po.fr_modulus(curve)
    rng = np.random.default_rng(11)
    sc = po.gen_scalars(curve, 1 << 10, 99).reshape(-1, 12)
    sc = np.tile(sc, (n >> 10, 1))                      # cheap "random" part
    kind = rng.integers(0, 16, size=n)
    one, two = po.int_to_limbs(po.R % r), po.int_to_limbs(2 * po.R % r)
    rm1, small = po.int_to_limbs((r - 1) * po.R % r), po.int_to_limbs(12345 * po.R % r)
    sc[kind < 8] = 0
    sc[(kind >= 8) & (kind < 12)] = one
    sc[(kind >= 12) & (kind < 14)] = two
    sc[kind == 14] = np.where((np.arange(n)[kind == 14] % 2 == 0)[:, None], rm1[None, :], small[None, :])
    sc = np.ascontiguousarray(sc).reshape(-1)
    ctx = ctxs[curve]
    k0, k1 = (po.int_to_limbs(po.sha512_rng_ints(r, s, 1)[0] * po.R % r) for s in (1000001, 1000002))
    slot = ctx.synthetic_bases(group, n, k0, k1)
    try:
        got = affine(oracle, curve, group, ctx.msm(slot, sc, n))
        t = ctx.last_timings()
        print("skewed witness 2^%d: %.2f ms" % (log_n, t["total"]))
        assert (got == oracle.msm_closed_form(curve, group, sc)).all()
    finally:
        ctx.free_bases(slot)

