"""Synthetic MSM inputs made without the oracle: uniform scalars in [0, r) as
Montgomery-form limbs, and the two seed scalars of the structured base set P0 + i*Q (the bases
themselves are generated on the device by b200msm_bases_synthetic)."""
import hashlib

import numpy as np

from .engine import MNT4753

MOD_A = 0x1C4C62D92C41110229022EEE2CDADB7F997505B8FAFED5EB7E8F96C97D87307FDB925E8A0ED8D99D124D9A15AF79DB117E776F218059DB80F0DA5CB537E38685ACCE9767254A4638810719AC425F0E39D54522CDD119F5E9063DE245E8001
MOD_B = 0x1C4C62D92C41110229022EEE2CDADB7F997505B8FAFED5EB7E8F96C97D87307FDB925E8A0ED8D99D124D9A15AF79DB26C5C28C859A99B3EEBCA9429212636B9DFF97634993AA4D6C381BC3F0057974EA099170FA13A4FD90776E240000001
R = 1 << 768


def fr_modulus(curve):
    """Scalar field of a curve = base field of the other one (multiexp/curves.cu:421-425)."""
    return MOD_B if curve == MNT4753 else MOD_A


def int_to_limbs(x):
    return np.frombuffer(int(x).to_bytes(96, "little"), dtype=np.uint64).copy()


def sha512_rng(modulus, idx):
    """libff::SHA512_rng<Fp>(idx) as a plain integer (depends/libff/libff/common/rng.tcc:26-80)."""
    mask = (1 << modulus.bit_length()) - 1
    it = 0
    while True:
        h0 = hashlib.sha512((2 * idx).to_bytes(8, "little") + it.to_bytes(8, "little")).digest()
        h1 = hashlib.sha512((2 * idx + 1).to_bytes(8, "little") + it.to_bytes(8, "little")).digest()
        v = int.from_bytes((h0 + h1)[:96], "little") & mask
        it += 1
        if v < modulus:
            return v


def base_seed_scalars(curve, seed_p0=1000001, seed_q=1000002):
    """Montgomery-form Fr limbs of SHA512_rng(seed_p0) and SHA512_rng(seed_q)."""
    r = fr_modulus(curve)
    return tuple(int_to_limbs(sha512_rng(r, s) * R % r) for s in (seed_p0, seed_q))


def random_scalars(curve, n, seed):
    """n canonical residues, uniform in [0, r), as uint64[n*12] (read by the engine as Montgomery-form
    scalars: the represented scalars s*R^-1 mod r are uniform as well).  Rejection sampling on 753-bit
    draws from numpy's PCG64; vectorised so that 2^24 scalars take seconds."""
    r = fr_modulus(curve)
    rl = int_to_limbs(r)
    rng = np.random.default_rng(seed)
    out = np.empty((n, 12), np.uint64)
    filled = 0
    while filled < n:
        m = int((n - filled) * 1.2) + 16
        x = rng.integers(0, 1 << 64, size=(m, 12), dtype=np.uint64)
        x[:, 11] &= np.uint64((1 << 49) - 1)
        lt = np.zeros(m, bool)
        eq = np.ones(m, bool)
        for j in range(11, -1, -1):
            lt |= eq & (x[:, j] < rl[j])
            eq &= x[:, j] == rl[j]
        x = x[lt]
        k = min(len(x), n - filled)
        out[filled:filled + k] = x[:k]
        filled += k
    return out.reshape(-1)


def write_instance(curve, d, directory, device=0, seed=7, infinities=(3, 5)):
    """Write a Groth16 instance of the reference's on-disk formats (SURVEY.md appendix A;
    generate_parameters.cpp:59-108, main.cpp:35-85) with m = d + 1 variables:
        <curve>-parameters   u64 d, u64 m, A[m+1] G1, B1[m+1] G1, B2[m+1] G2, L[m-1] G1, H[d] G1
        <curve>-input        w[m+1] (w[0] = 1), ca[d+1], cb[d+1], cc[d+1], r   -- Fr, Montgomery limbs
    The five queries are structured base sets P0 + i*Q generated in HBM by the engine (distinct P0, Q per
    query, all in the prime-order subgroup) with the points at `infinities` of A, B1 and B2 replaced by the
    encoding of infinity (all-zero), as real proving keys have; the witness and the QAP evaluations are uniform
    field elements.  Any prover that follows the reference's `compute` semantics -- the reference CPU prover
    `main` included -- maps these two files to one well-defined proof; that is all the parity check needs, and it
    replaces three minutes of `generate_parameters` at the default size (d = 2^20 - 1) by seconds.
    -> (parameters path, input path)."""
    import os

    from .engine import G1, G2, MsmContext, degree
    name = "MNT4753" if curve == MNT4753 else "MNT6753"
    m = d + 1
    r = fr_modulus(curve)
    ppath, ipath = os.path.join(directory, name + "-parameters"), os.path.join(directory, name + "-input")
    with MsmContext(curve, device) as ctx:
        ctx.set_table_budget(0)                    # points only: no window tables for a set that is downloaded at once
        with open(ppath + ".tmp", "wb") as f:
            np.array([d, m], np.uint64).tofile(f)
            for q, (group, n) in enumerate(((G1, m + 1), (G1, m + 1), (G2, m + 1), (G1, m - 1), (G1, d))):
                k0, k1 = base_seed_scalars(curve, 2000001 + 16 * seed + 2 * q, 2000002 + 16 * seed + 2 * q)
                slot = ctx.synthetic_bases(group, n, k0, k1)
                pts = ctx.download_bases(slot, 0, n).reshape(n, 24 * degree(curve, group))
                ctx.free_bases(slot)
                if q < 3:
                    for i in infinities:
                        if i < n:
                            pts[i] = 0
                pts.tofile(f)
    os.replace(ppath + ".tmp", ppath)
    with open(ipath + ".tmp", "wb") as f:
        w = random_scalars(curve, m + 1, 9000 + seed).reshape(m + 1, 12)
        w[0] = int_to_limbs(R % r)                 # w[0] = 1 in Montgomery form
        w.tofile(f)
        for j in range(3):
            random_scalars(curve, d + 1, 9100 + 4 * seed + j).tofile(f)
        random_scalars(curve, 1, 9200 + seed).tofile(f)
    os.replace(ipath + ".tmp", ipath)
    return ppath, ipath
