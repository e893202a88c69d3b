"""ctypes mirror of include/b200_msm.h.

Vocabulary follows the reference: a *base set* is one query of the proving key (A, B1, L, H in G1,
B2 in G2; cuda_prover_piecewise.cu:132-139), scalars are Montgomery-form Fr elements (12 x u64), the
result of an MSM is one Jacobian point X||Y||Z that B::read_pt_ECp / read_pt_ECpe consume
(prover_reference_functions.cpp:795-817).
"""
import ctypes
import os

import numpy as np

MNT4753, MNT6753 = 0, 1
G1, G2 = 1, 2
_ERRORS = {1: "invalid argument", 2: "CUDA error / no usable sm_100 device", 3: "out of device memory"}
_u64p = ctypes.POINTER(ctypes.c_uint64)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_lib = None


class MsmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b200msm error %d (%s): %s" % (code, _ERRORS.get(code, "?"), msg))
        self.code = code


def degree(curve, group):
    """Extension degree of the coordinate field: 1 (G1), 2 (MNT4753 G2), 3 (MNT6753 G2)."""
    return 1 if group == G1 else (2 if curve == MNT4753 else 3)


def library_path():
    """In-tree libb200msm.so; B200MSM_LIB overrides it with another build of the SAME library (A/B runs of
    kernel variants during development).  There is no other implementation to fall back to."""
    return os.environ.get("B200MSM_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libb200msm.so")


def load_library():
    """Load libb200msm.so.  There is no fallback: a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    # hardware queues for the engine's streams (see b200msm_create); only effective before the process's first CUDA call
    os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
    path = library_path()
    if not os.path.exists(path):
        raise ImportError("%s is missing: build it with `python -m gpu_groth16_prover_3x_b200.build` "
                          "(or __graft_entry__.build()); the MSM engine has no CPU fallback" % path)
    lib = ctypes.CDLL(path)
    vp, ci, sz = ctypes.c_void_p, ctypes.c_int, ctypes.c_size_t
    lib.b200msm_create.argtypes = [ci, ci, ctypes.POINTER(vp)]
    lib.b200msm_destroy.argtypes = [vp]
    lib.b200msm_destroy.restype = None
    lib.b200msm_last_error.argtypes = [vp]
    lib.b200msm_last_error.restype = ctypes.c_char_p
    lib.b200msm_bases_upload.argtypes = [vp, ci, vp, sz, ctypes.POINTER(ci)]
    lib.b200msm_bases_free.argtypes = [vp, ci]
    lib.b200msm_msm.argtypes = [vp, ci, sz, vp, sz, vp]
    lib.b200msm_msm_async.argtypes = [vp, ci, ci, sz, vp, sz, vp]
    lib.b200msm_wait.argtypes = [vp, ci]
    lib.b200msm_ec_reduce.argtypes = [vp, ci, vp, vp, sz, vp]
    lib.b200msm_bases_synthetic.argtypes = [vp, ci, sz, vp, vp, ctypes.POINTER(ci)]
    lib.b200msm_bases_download.argtypes = [vp, ci, sz, sz, vp]
    lib.b200msm_to_affine.argtypes = [vp, ci, sz, vp, vp]
    lib.b200msm_fold.argtypes = [vp, ci, vp, sz, vp]
    lib.b200msm_shard_range.argtypes = [sz, ci, ci, ctypes.POINTER(sz), ctypes.POINTER(sz)]
    lib.b200msm_set_stream.argtypes = [vp, ci, vp]
    lib.b200msm_set_window_bits.argtypes = [vp, ci]
    lib.b200msm_set_lane_sms.argtypes = [vp, ci, ci]
    lib.b200msm_set_table_budget.argtypes = [vp, sz]
    lib.b200msm_scalar_mul.argtypes = [vp, ci, vp, vp, vp]
    lib.b200msm_key_load.argtypes = [vp, vp, sz, ctypes.POINTER(vp)]
    lib.b200msm_key_load_file.argtypes = [vp, ctypes.c_char_p, ctypes.POINTER(vp)]
    lib.b200msm_key_info.argtypes = [vp, _u64p]
    lib.b200msm_key_free.argtypes = [vp, vp]
    lib.b200msm_key_free.restype = None
    lib.b200msm_proof_bytes.argtypes = [vp]
    lib.b200msm_proof_bytes.restype = sz
    lib.b200msm_input_bytes.argtypes = [vp]
    lib.b200msm_input_bytes.restype = sz
    lib.b200msm_prove.argtypes = [vp, vp, vp, sz, vp]
    lib.b200msm_prove_file.argtypes = [vp, vp, ctypes.c_char_p, vp, vp]
    lib.b200msm_key_load_shard.argtypes = [vp, vp, sz, ci, ci, ctypes.POINTER(vp)]
    lib.b200msm_key_load_sharded_file.argtypes = [ctypes.POINTER(vp), ci, ctypes.c_char_p, ctypes.POINTER(vp)]
    lib.b200msm_prove_sharded.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(vp), ci, vp, sz, vp]
    lib.b200msm_prove_sharded_file.argtypes = [ctypes.POINTER(vp), ctypes.POINTER(vp), ci, ctypes.c_char_p, vp, vp]
    lib.b200msm_pinned_alloc.argtypes = [sz]
    lib.b200msm_pinned_alloc.restype = vp
    lib.b200msm_pinned_free.argtypes = [vp]
    lib.b200msm_pinned_free.restype = None
    lib.b200msm_compute_h.argtypes = [vp, sz, vp, vp, vp, vp, ctypes.POINTER(vp)]
    lib.b200msm_compute_h_timings.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
    lib.b200msm_fft_release.argtypes = [vp]
    lib.b200msm_fft_release.restype = None
    lib.b200msm_bases_info.argtypes = [vp, ci, _u64p]
    lib.b200msm_last_timings.argtypes = [vp, ci, ctypes.POINTER(ctypes.c_float), _u64p]
    lib.b200msm_last_rounds.argtypes = [vp, ci, _u64p, ctypes.POINTER(ctypes.c_uint32), sz]
    lib.b200msm_microbench.argtypes = [vp, ci, ci, ctypes.POINTER(ctypes.c_double)]
    lib.b200msm_selftest_field.argtypes = [vp, ci, ci, sz, vp, vp, vp]
    lib.b200msm_selftest_point.argtypes = [vp, ci, ci, sz, vp, vp, vp, vp]
    _lib = lib
    return lib


def shard_ranges(n, parts):
    """Point-range sharding of an MSM of n points over `parts` GPUs: shard g owns
    [n g / parts, n (g + 1) / parts) -- the rule of b200msm_shard_range (include/b200_msm.h), which
    b200msm_key_load_shard uses as well (tests/test_abi.py asserts the two agree).  -> list of (offset, length)."""
    if parts < 1:
        raise ValueError("parts must be >= 1")
    return [(n * g // parts, n * (g + 1) // parts - n * g // parts) for g in range(parts)]


def _ptr(x):
    """Address of a numpy array (host) or torch tensor (host or device); None -> NULL."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return x.ctypes.data
    if isinstance(x, int):
        return x
    assert x.is_contiguous()
    return x.data_ptr()


def _nbytes(x):
    return x.nbytes if isinstance(x, np.ndarray) else x.numel() * x.element_size()


class MsmContext:
    """One engine context per (curve, GPU)."""

    PHASES = ("total", "h2d_scalars", "recode_sort", "accumulate", "reduce_combine", "d2h_result")

    def __init__(self, curve, device=0):
        self.lib = load_library()
        self.curve = curve
        self.device = device
        self._h = ctypes.c_void_p()
        rc = self.lib.b200msm_create(curve, device, ctypes.byref(self._h))
        if rc:
            self._h = ctypes.c_void_p()
            raise MsmError(rc, "b200msm_create(curve=%d, device=%d) failed" % (curve, device))
        self._slot_group = {}
        self._pending = {}

    def close(self):
        if self._h:
            self.lib.b200msm_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc:
            raise MsmError(rc, self.lib.b200msm_last_error(self._h).decode())

    # ---- base sets ---------------------------------------------------------------------------
    def upload_bases(self, group, affine, n=None):
        """affine: n points x||y in the wire format (uint64 limbs; host array or device tensor)."""
        per = 24 * degree(self.curve, group) * 8
        if n is None:
            n = _nbytes(affine) // per
        assert _nbytes(affine) >= n * per
        slot = ctypes.c_int(-1)
        self._check(self.lib.b200msm_bases_upload(self._h, group, _ptr(affine), n, ctypes.byref(slot)))
        self._slot_group[slot.value] = (group, n)
        return slot.value

    def synthetic_bases(self, group, n, k_p0_mont, k_q_mont):
        """Resident base set P0 + i*Q, P0 = k_p0*G, Q = k_q*G, generated on the device."""
        slot = ctypes.c_int(-1)
        k0 = np.ascontiguousarray(k_p0_mont, dtype=np.uint64)
        k1 = np.ascontiguousarray(k_q_mont, dtype=np.uint64)
        assert k0.size == 12 and k1.size == 12
        self._check(self.lib.b200msm_bases_synthetic(self._h, group, n, k0.ctypes.data, k1.ctypes.data, ctypes.byref(slot)))
        self._slot_group[slot.value] = (group, n)
        return slot.value

    def download_bases(self, slot, offset=0, n=None):
        group, nb = self._slot_group[slot]
        if n is None:
            n = nb - offset
        out = np.zeros(n * 24 * degree(self.curve, group), np.uint64)
        self._check(self.lib.b200msm_bases_download(self._h, slot, offset, n, out.ctypes.data))
        return out

    def to_affine(self, group, xyz):
        """Jacobian X||Y||Z points -> affine wire format on the device (infinity -> zeros)."""
        xyz = np.ascontiguousarray(xyz, dtype=np.uint64).reshape(-1)
        n = xyz.size // (36 * degree(self.curve, group))
        out = np.zeros(n * 24 * degree(self.curve, group), np.uint64)
        self._check(self.lib.b200msm_to_affine(self._h, group, n, xyz.ctypes.data, out.ctypes.data))
        return out

    def free_bases(self, slot):
        self._check(self.lib.b200msm_bases_free(self._h, slot))
        self._slot_group.pop(slot, None)

    # ---- MSM ---------------------------------------------------------------------------------
    def _out(self, group):
        return np.zeros(36 * degree(self.curve, group), np.uint64)

    def msm(self, slot, scalars, n=None, offset=0):
        """sum_i scalars[i] * bases[offset + i] -> Jacobian X||Y||Z (uint64[36*DEG])."""
        group, nb = self._slot_group[slot]
        if n is None:
            n = _nbytes(scalars) // 96
        out = self._out(group)
        self._check(self.lib.b200msm_msm(self._h, slot, offset, _ptr(scalars), n, out.ctypes.data))
        return out

    def msm_async(self, lane, slot, scalars, n=None, offset=0):
        group, nb = self._slot_group[slot]
        if n is None:
            n = _nbytes(scalars) // 96
        out = self._out(group)
        self._check(self.lib.b200msm_msm_async(self._h, lane, slot, offset, _ptr(scalars), n, out.ctypes.data))
        self._pending[lane] = (out, scalars)  # keep both alive until wait()

    def wait(self, lane):
        self._check(self.lib.b200msm_wait(self._h, lane))
        out, _ = self._pending.pop(lane)
        return out

    def ec_reduce(self, group, bases_affine, scalars, n=None):
        """Literal counterpart of ec_reduce_straus (multiexp/reduce.cu:131-152): bases travel with the call."""
        if n is None:
            n = _nbytes(scalars) // 96
        out = self._out(group)
        self._check(self.lib.b200msm_ec_reduce(self._h, group, _ptr(bases_affine), _ptr(scalars), n, out.ctypes.data))
        return out

    def fold(self, group, partials_xyz):
        """Sum of Jacobian partial results (one per GPU shard) -> one Jacobian point, on this GPU."""
        per = 36 * degree(self.curve, group)
        partials_xyz = np.ascontiguousarray(partials_xyz, dtype=np.uint64).reshape(-1)
        n = partials_xyz.size // per
        out = self._out(group)
        self._check(self.lib.b200msm_fold(self._h, group, partials_xyz.ctypes.data, n, out.ctypes.data))
        return out

    # ---- tuning / introspection --------------------------------------------------------------
    def set_stream(self, lane, cuda_stream):
        """Run lane `lane` on a caller-owned CUDA stream (integer handle, e.g. torch Stream.cuda_stream); 0 = internal."""
        self._check(self.lib.b200msm_set_stream(self._h, lane, ctypes.c_void_p(cuda_stream or None)))

    def set_window_bits(self, c):
        self._check(self.lib.b200msm_set_window_bits(self._h, c))

    def set_lane_sms(self, lane, sms):
        """MSMs enqueued on `lane` afterwards occupy at most `sms` SMs (0: all), so that lanes run side by side."""
        self._check(self.lib.b200msm_set_lane_sms(self._h, lane, sms))

    def last_timings(self, lane=0):
        ms = (ctypes.c_float * 6)()
        info = (ctypes.c_uint64 * 8)()
        self._check(self.lib.b200msm_last_timings(self._h, lane, ms, info))
        d = {k: float(ms[i]) for i, k in enumerate(self.PHASES)}
        d.update(window_bits=int(info[0]), windows=int(info[1]), entries=int(info[2]),
                 shares=int(info[3]), kernel_launches=int(info[4]), bucket_sets=int(info[5]), tables=int(info[6]))
        return d

    def scalar_mul(self, group, affine, k_mont):
        out = self._out(group)
        self._check(self.lib.b200msm_scalar_mul(self._h, group, _ptr(affine), _ptr(k_mont), _ptr(out)))
        return out

    # ---- whole proofs (run_prover of cuda_prover_piecewise.cu:96-230) ----
    def load_key(self, params):
        """params: path of a <curve>-parameters file, or its bytes (numpy uint8 / bytes).  -> opaque key handle."""
        key = ctypes.c_void_p()
        if isinstance(params, str):
            self._check(self.lib.b200msm_key_load_file(self._h, params.encode(), ctypes.byref(key)))
        else:
            buf = np.frombuffer(params, dtype=np.uint8) if isinstance(params, (bytes, bytearray)) else np.ascontiguousarray(params).view(np.uint8)
            self._check(self.lib.b200msm_key_load(self._h, _ptr(buf), buf.size, ctypes.byref(key)))
        return key

    def key_info(self, key):
        info = (ctypes.c_uint64 * 2)()
        self._check(self.lib.b200msm_key_info(key, info))
        return {"d": int(info[0]), "m": int(info[1])}

    def free_key(self, key):
        self.lib.b200msm_key_free(self._h, key)

    def prove(self, key, input_image):
        """input_image: bytes of the <curve>-input file (bytes or numpy uint8).  -> proof bytes (A || B || C affine)."""
        buf = np.frombuffer(input_image, dtype=np.uint8) if isinstance(input_image, (bytes, bytearray)) else np.ascontiguousarray(input_image).view(np.uint8)
        proof = np.zeros(self.lib.b200msm_proof_bytes(self._h), np.uint8)
        self._check(self.lib.b200msm_prove(self._h, key, _ptr(buf), buf.size, _ptr(proof)))
        return proof.tobytes()

    def load_key_shard(self, params, shard, nshards):
        """Shard `shard` of `nshards` (point range of every query) of a <curve>-parameters image (bytes / numpy uint8)."""
        key = ctypes.c_void_p()
        buf = np.frombuffer(params, dtype=np.uint8) if isinstance(params, (bytes, bytearray)) else np.ascontiguousarray(params).view(np.uint8)
        self._check(self.lib.b200msm_key_load_shard(self._h, _ptr(buf), buf.size, shard, nshards, ctypes.byref(key)))
        return key

    def prove_file(self, key, input_path):
        """The same from the <curve>-input file: the witness MSMs start while the rest of the file is still being read."""
        n = self.lib.b200msm_input_bytes(key)
        buf = self.lib.b200msm_pinned_alloc(n)
        if not buf:
            raise MsmError(3, "cannot allocate %d bytes of pinned memory" % n)
        try:
            proof = np.zeros(self.lib.b200msm_proof_bytes(self._h), np.uint8)
            self._check(self.lib.b200msm_prove_file(self._h, key, input_path.encode(), buf, _ptr(proof)))
        finally:
            self.lib.b200msm_pinned_free(buf)
        return proof.tobytes()

    def compute_h(self, ca, cb, cc, to_host=True):
        """coefficients_for_H of the Groth16 prover (compute_H, cuda_prover_piecewise.cu:14-49) from the d + 1
        evaluations ca, cb, cc (Montgomery Fr limbs, numpy or torch).  -> (uint64[(d + 2) * 12] or None, device pointer)."""
        m = _nbytes(ca) // 96
        out = np.zeros((m + 1) * 12, np.uint64) if to_host else None
        dev = ctypes.c_void_p()
        self._check(self.lib.b200msm_compute_h(self._h, m - 1, _ptr(ca), _ptr(cb), _ptr(cc), _ptr(out), ctypes.byref(dev)))
        return out, dev.value

    def compute_h_timings(self):
        ms = (ctypes.c_float * 2)()
        self._check(self.lib.b200msm_compute_h_timings(self._h, ms))
        return {"compute_h_ms": float(ms[0]), "tables_ms": float(ms[1])}

    def set_table_budget(self, max_bytes_per_set):
        """Byte budget of the window tables built at upload time (0: none), see include/b200_msm.h."""
        self._check(self.lib.b200msm_set_table_budget(self._h, max_bytes_per_set))

    def bases_info(self, slot):
        info = (ctypes.c_uint64 * 6)()
        self._check(self.lib.b200msm_bases_info(self._h, slot, info))
        return dict(points=int(info[0]), table_window_bits=int(info[1]), tables=int(info[2]), bucket_sets=int(info[3]),
                    bytes=int(info[4]), table_build_ms=int(info[5]) / 1e3)

    def last_rounds(self, lane=0):
        info = (ctypes.c_uint64 * 4)()
        pairs = (ctypes.c_uint32 * 32)()
        self._check(self.lib.b200msm_last_rounds(self._h, lane, info, pairs, 32))
        return dict(rounds=int(info[0]), shares=int(info[1]), max_bucket_occupancy=int(info[2]),
                    additions=int(info[3]), pairs_per_round=[int(x) for x in pairs[:int(info[0])]])

    def microbench(self, kind, iters=4096):
        g = ctypes.c_double()
        self._check(self.lib.b200msm_microbench(self._h, kind, iters, ctypes.byref(g)))
        return g.value

    # ---- self-test hooks (tests only) --------------------------------------------------------
    def selftest_field(self, group, op, a, b=None):
        a = np.ascontiguousarray(a, dtype=np.uint64)
        n = a.size // (12 * degree(self.curve, group))
        out = np.zeros_like(a)
        self._check(self.lib.b200msm_selftest_field(self._h, group, op, n, a.ctypes.data, _ptr(b), out.ctypes.data))
        return out

    def selftest_point(self, group, op, acc, q=None, flags=None):
        acc = np.ascontiguousarray(acc, dtype=np.uint64)
        n = acc.size // (36 * degree(self.curve, group))
        out = np.zeros_like(acc)
        if flags is not None:
            flags = np.ascontiguousarray(flags, dtype=np.uint32)
        self._check(self.lib.b200msm_selftest_point(self._h, group, op, n, acc.ctypes.data, _ptr(q), _ptr(flags), out.ctypes.data))
        return out


def load_key_sharded_file(ctxs, path):
    """All shards of a <curve>-parameters file, shard g into ctxs[g] (one host thread per GPU):
    b200msm_key_load_sharded_file.  -> list of key handles, to be freed with ctxs[g].free_key(keys[g])."""
    n = len(ctxs)
    lib = ctxs[0].lib
    ca = (ctypes.c_void_p * n)(*[c._h for c in ctxs])
    ka = (ctypes.c_void_p * n)()
    ctxs[0]._check(lib.b200msm_key_load_sharded_file(ca, n, path.encode(), ka))
    return [ctypes.c_void_p(ka[g]) for g in range(n)]


def prove_sharded(ctxs, keys, input_image):
    """One proof over several contexts (one per GPU; ctxs[g] / keys[g] = shard g): b200msm_prove_sharded."""
    n = len(ctxs)
    lib = ctxs[0].lib
    buf = np.frombuffer(input_image, dtype=np.uint8) if isinstance(input_image, (bytes, bytearray)) else np.ascontiguousarray(input_image).view(np.uint8)
    proof = np.zeros(lib.b200msm_proof_bytes(ctxs[0]._h), np.uint8)
    ca = (ctypes.c_void_p * n)(*[c._h for c in ctxs])
    ka = (ctypes.c_void_p * n)(*[k for k in keys])
    ctxs[0]._check(lib.b200msm_prove_sharded(ca, ka, n, _ptr(buf), buf.size, _ptr(proof)))
    return proof.tobytes()
