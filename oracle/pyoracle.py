"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the CPU oracle (oracle/liboracle.so, the plain-C
restatement) and, when present, the reference's own libff (oracle/_ref/libref.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product package never does.
"""
import ctypes
import hashlib
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_u64p = ctypes.POINTER(ctypes.c_uint64)
MNT4753, MNT6753 = 0, 1
G1, G2 = 1, 2

MOD_A = 0x1C4C62D92C41110229022EEE2CDADB7F997505B8FAFED5EB7E8F96C97D87307FDB925E8A0ED8D99D124D9A15AF79DB117E776F218059DB80F0DA5CB537E38685ACCE9767254A4638810719AC425F0E39D54522CDD119F5E9063DE245E8001
MOD_B = 0x1C4C62D92C41110229022EEE2CDADB7F997505B8FAFED5EB7E8F96C97D87307FDB925E8A0ED8D99D124D9A15AF79DB26C5C28C859A99B3EEBCA9429212636B9DFF97634993AA4D6C381BC3F0057974EA099170FA13A4FD90776E240000001
R = 1 << 768


def fq_modulus(curve):
    return MOD_A if curve == MNT4753 else MOD_B


def fr_modulus(curve):
    return MOD_B if curve == MNT4753 else MOD_A


def degree(curve, group):
    return 1 if group == G1 else (2 if curve == MNT4753 else 3)


def int_to_limbs(x):
    return np.frombuffer(int(x).to_bytes(96, "little"), dtype=np.uint64).copy()


def limbs_to_int(a):
    return int.from_bytes(np.ascontiguousarray(a, dtype=np.uint64).tobytes(), "little")


def ints_to_array(xs):
    return np.frombuffer(b"".join(int(x).to_bytes(96, "little") for x in xs), dtype=np.uint64).copy()


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_u64p)


def sha512_rng_ints(modulus, idx0, n):
    """Plain-integer values of libff::SHA512_rng<Fp>(idx0 + i) (depends/libff/libff/common/rng.tcc:26-80)."""
    nbits = modulus.bit_length()
    mask = (1 << nbits) - 1
    out = []
    for i in range(n):
        idx = idx0 + i
        it = 0
        while True:
            h0 = hashlib.sha512((2 * idx).to_bytes(8, "little") + it.to_bytes(8, "little")).digest()
            h1 = hashlib.sha512((2 * idx + 1).to_bytes(8, "little") + it.to_bytes(8, "little")).digest()
            v = int.from_bytes((h0 + h1)[:96], "little") & mask
            it += 1
            if v < modulus:
                break
        out.append(v)
    return out


def gen_scalars(curve, n, seed):
    """Montgomery-form Fr limbs of SHA512_rng(seed * 2^32 + i), i < n  ->  uint64[n*12]."""
    r = fr_modulus(curve)
    return ints_to_array([(v * R) % r for v in sha512_rng_ints(r, seed << 32, n)]) if n else np.zeros(0, np.uint64)


class CpuLib:
    """Uniform wrapper over liboracle.so (prefix 'orc') and libref.so (prefix 'ref')."""

    def __init__(self, path, prefix):
        self.lib = ctypes.CDLL(path)
        self.prefix = prefix
        self.kind = "port" if prefix == "orc" else "reference"
        f = self._f
        f("msm").restype = ctypes.c_double
        f("msm").argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_size_t, _u64p, _u64p, _u64p, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        f("field_op").argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, _u64p, _u64p, _u64p]
        f("point_op").argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, _u64p, _u64p, _u64p, _u64p]
        f("fr_from_mont").argtypes = [ctypes.c_int, ctypes.c_size_t, _u64p, _u64p]
        f("fr_to_mont").argtypes = [ctypes.c_int, ctypes.c_size_t, _u64p, _u64p]
        f("jacobian_to_affine").argtypes = [ctypes.c_int, ctypes.c_int, _u64p, _u64p]
        f("set_num_threads").argtypes = [ctypes.c_int]
        f("compute_h").argtypes = [ctypes.c_int, ctypes.c_size_t, _u64p, _u64p, _u64p, _u64p]
        if prefix == "orc":
            f("gen_bases").argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_size_t, _u64p, _u64p, _u64p]
            f("msm_closed_form").argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_size_t, _u64p, _u64p, _u64p, _u64p]
            f("fold_jacobian").argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_size_t, _u64p, _u64p]
        else:
            self.lib.ref_init()
            f("gen_bases").argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.c_uint64, ctypes.c_uint64, _u64p]
            f("gen_scalars").argtypes = [ctypes.c_int, ctypes.c_size_t, ctypes.c_uint64, _u64p]

    def _f(self, name):
        return getattr(self.lib, "%s_%s" % (self.prefix, name))

    def num_threads(self):
        return self._f("num_threads")()

    def set_num_threads(self, t):
        self._f("set_num_threads")(t)

    def field_op(self, curve, field, op, a, b=None):
        deg = 1 if field == 0 else degree(curve, G2)
        n = a.size // (12 * deg)
        out = np.zeros_like(a)
        assert self._f("field_op")(curve, field, op, n, _p(a), _p(b), _p(out)) == 0
        return out

    def fr_from_mont(self, curve, a):
        out = np.zeros_like(a)
        self._f("fr_from_mont")(curve, a.size // 12, _p(a), _p(out))
        return out

    def fr_to_mont(self, curve, a):
        out = np.zeros_like(a)
        self._f("fr_to_mont")(curve, a.size // 12, _p(a), _p(out))
        return out

    def point_op(self, curve, group, op, a, b=None, k=None):
        out = np.zeros(24 * degree(curve, group), np.uint64)
        assert self._f("point_op")(curve, group, op, _p(a), _p(b), _p(k), _p(out)) == 0
        return out

    def msm(self, curve, group, bases, scalars, method=1, chunks=0, prefilter=1):
        """-> (affine result uint64[24*deg], seconds inside the MSM call)."""
        n = scalars.size // 12
        assert bases.size == n * 24 * degree(curve, group)
        out = np.zeros(24 * degree(curve, group), np.uint64)
        t = self._f("msm")(curve, group, n, _p(bases), _p(scalars), _p(out), method, chunks, prefilter)
        assert t >= 0
        return out, t

    def compute_h(self, curve, ca, cb, cc):
        """coefficients_for_H (compute_H, cuda_prover_piecewise.cu:14-49): d + 1 evaluations each -> d + 2 coefficients."""
        m = ca.size // 12
        out = np.zeros((m + 1) * 12, np.uint64)
        rc = self._f("compute_h")(curve, m - 1, _p(np.ascontiguousarray(ca)), _p(np.ascontiguousarray(cb)),
                                  _p(np.ascontiguousarray(cc)), _p(out))
        assert rc == 0, rc
        return out

    def jacobian_to_affine(self, curve, group, xyz):
        out = np.zeros(24 * degree(curve, group), np.uint64)
        assert self._f("jacobian_to_affine")(curve, group, _p(np.ascontiguousarray(xyz)), _p(out)) == 0
        return out

    # --- oracle-only helpers -------------------------------------------------------------
    def generator(self, curve, group):
        return GENERATORS[(curve, group)].copy()

    def base_pair(self, curve, group, seed_p0=1000001, seed_q=1000002):
        """P0 = SHA512_rng(seed_p0)*G, Q = SHA512_rng(seed_q)*G, affine."""
        r = fr_modulus(curve)
        k0, k1 = (ints_to_array([(sha512_rng_ints(r, s, 1)[0] * R) % r]) for s in (seed_p0, seed_q))
        g = self.generator(curve, group)
        return self.point_op(curve, group, 3, g, k=k0), self.point_op(curve, group, 3, g, k=k1)

    def gen_bases(self, curve, group, n, seed_p0=1000001, seed_q=1000002):
        out = np.zeros(n * 24 * degree(curve, group), np.uint64)
        if self.prefix == "ref":
            assert self._f("gen_bases")(curve, group, n, seed_p0, seed_q, _p(out)) == 0
        else:
            p0, q = self.base_pair(curve, group, seed_p0, seed_q)
            assert self._f("gen_bases")(curve, group, n, _p(p0), _p(q), _p(out)) == 0
        return out

    def gen_scalars(self, curve, n, seed):
        if self.prefix == "ref":
            out = np.zeros(n * 12, np.uint64)
            self._f("gen_scalars")(curve, n, seed, _p(out))
            return out
        return gen_scalars(curve, n, seed)

    def msm_closed_form(self, curve, group, scalars, seed_p0=1000001, seed_q=1000002):
        p0, q = self.base_pair(curve, group, seed_p0, seed_q)
        out = np.zeros(24 * degree(curve, group), np.uint64)
        assert self._f("msm_closed_form")(curve, group, scalars.size // 12, _p(p0), _p(q), _p(scalars), _p(out)) == 0
        return out

    def fold_jacobian(self, curve, group, xyz):
        out = np.zeros(24 * degree(curve, group), np.uint64)
        n = xyz.size // (36 * degree(curve, group))
        assert self._f("fold_jacobian")(curve, group, n, _p(np.ascontiguousarray(xyz)), _p(out)) == 0
        return out


# G1_one / G2_one of both curves in affine wire format (Montgomery limbs); values dumped from the
# reference's libff (mnt4753_init.cpp:150-170, mnt6753_init.cpp:160-185) by tools/gen_golden.py.
GENERATORS = {}


def _load_generators():
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "generators.npz")
    if os.path.exists(path):
        z = np.load(path)
        for c in (0, 1):
            for g in (1, 2):
                GENERATORS[(c, g)] = z["c%d_g%d" % (c, g)]


_load_generators()


def load_oracle():
    path = os.path.join(HERE, "liboracle.so")
    if not os.path.exists(path):
        raise FileNotFoundError("oracle/liboracle.so missing: run `make -C oracle oracle` or __graft_entry__.build()")
    return CpuLib(path, "orc")


def load_reference():
    """The reference's libff built by oracle/Makefile; None when it has not been built."""
    path = os.path.join(HERE, "_ref", "libref.so")
    return CpuLib(path, "ref") if os.path.exists(path) else None
