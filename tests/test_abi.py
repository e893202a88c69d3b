"""The C-ABI library loads and exports every symbol include/b200_msm.h declares (no compute here)."""
import ctypes
import os
import re

import pytest

import gpu_groth16_prover_3x_b200 as pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200_msm.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200msm_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(engine_lib):
    syms = declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(engine_lib, s), "libb200msm.so does not export %s" % s


def test_every_symbol_is_documented():
    """INTEGRATION.md names every entry point of the header (in full, or as `_suffix` / `create/destroy` shorthand)."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for s in declared_symbols():
        short = s[len("b200msm"):]
        assert s in doc or short in doc or short.lstrip("_") in doc, "INTEGRATION.md does not mention %s" % s


def test_no_oracle_in_product():
    """The product path never touches oracle/ (no import, no dlopen, no link)."""
    pk = os.path.join(ROOT, "gpu_groth16_prover_3x_b200")
    for dirpath, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in src and "liboracle" not in src and "libref" not in src, f
    import subprocess
    out = subprocess.run(["ldd", pkg.library_path()], capture_output=True, text=True).stdout
    assert "oracle" not in out and "libref" not in out


def test_shard_ranges():
    for n in (0, 1, 7, 8, 9, 1 << 20, (1 << 20) + 1):
        for parts in (1, 2, 4, 8):
            r = pkg.shard_ranges(n, parts)
            assert len(r) == parts and sum(l for _, l in r) == n
            assert all(r[i][0] + r[i][1] == r[i + 1][0] for i in range(parts - 1)) and r[0][0] == 0
            assert max(l for _, l in r) - min(l for _, l in r) <= 1


def test_shard_rule_is_the_abis(engine_lib):
    """engine.shard_ranges (python callers) and b200msm_shard_range (the C ABI, which b200msm_key_load_shard
    uses) are one rule: a caller that mixes sharded keys with python-side ranges stays aligned."""
    off, ln = ctypes.c_size_t(), ctypes.c_size_t()
    for n in (0, 1, 7, 9, 1000, (1 << 20) - 1, (1 << 20) + 1, (1 << 24) + 5):
        for parts in (1, 2, 3, 4, 5, 8):
            want = pkg.shard_ranges(n, parts)
            for g in range(parts):
                assert engine_lib.b200msm_shard_range(n, g, parts, ctypes.byref(off), ctypes.byref(ln)) == 0
                assert (off.value, ln.value) == want[g]
    assert engine_lib.b200msm_shard_range(8, 2, 2, ctypes.byref(off), ctypes.byref(ln)) == 1
    assert engine_lib.b200msm_shard_range(8, 0, 0, ctypes.byref(off), ctypes.byref(ln)) == 1


def test_msm_error_carries_code():
    e = pkg.MsmError(3, "cannot allocate")
    assert e.code == 3 and "out of device memory" in str(e)


def test_create_fails_loudly_without_gpu(engine_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(pkg.MsmError) as e:
        pkg.MsmContext(pkg.MNT4753, 0)
    assert e.value.code == 2
    h = ctypes.c_void_p()
    assert engine_lib.b200msm_create(7, 0, ctypes.byref(h)) == 1  # bad curve id


def test_lane_split_model(engine_lib):
    """The SM split b200msm_prove gives the five MSMs of a proof (csrc/prover.cu, pure host arithmetic): lanes get
    disjoint parts of the GPU that add up to all of it for small proofs, the large ones (bound by throughput) are
    left one after the other on the whole GPU, and the environment knob's modes 0 / 2 mean never / always."""
    f = engine_lib.b200msm_internal_lane_split_model
    f.argtypes = [ctypes.c_int, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                  ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_double)]

    def plan(curve, log_d, shard=0, nshards=1, mode=1, sm=148, delay_ms=0.0):
        sms, ns = (ctypes.c_int * 5)(), (ctypes.c_double * 2)()
        on = f(curve, (1 << log_d) - 1, 1 << log_d, shard, nshards, delay_ms * 1e6, sm, mode, sms, ns)
        return on, list(sms), ns[0], ns[1]

    for curve, log_d in ((1, 15), (0, 14), (1, 10), (0, 17)):        # default MNT6753, the two fast instances, a 2^17 shard
        on, sms, serial, side = plan(curve, log_d)
        assert on == 1 and sum(sms) == 148 and min(sms) >= 1 and side < 0.85 * serial
        assert sms[2] == max(sms)                                     # the G2 query is the largest party
        assert sms[0] == sms[1]                                       # A and B1: same size, same share
    on, sms, serial, side = plan(0, 20)                               # default MNT4753 on one GPU: throughput-bound
    assert on == 0 and sms == [0] * 5
    assert plan(0, 20, mode=2)[0] == 1 and sum(plan(0, 20, mode=2)[1]) == 148
    assert plan(1, 15, mode=0)[0] == 0
    on, sms, _, _ = plan(0, 20, shard=0, nshards=8)                   # an eighth of it, with the FFTs on this GPU
    assert on == 1 and sum(sms) == 148 and sms[4] >= 8
    # ... but not when the FFTs' input (300 MB of the input file) arrives 30 ms into the proof: file read -> FFTs -> H
    # query is then the critical path and wants the whole GPU (measured: profiles/r02_proof_8gpu_lane_trace.txt)
    assert plan(0, 20, shard=0, nshards=8, delay_ms=30.0)[0] == 0
    assert plan(1, 15, delay_ms=0.9)[0] == 1                          # 9.4 MB: no matter
    on, sms, _, _ = plan(1, 10, sm=16)                                # a small GPU: still whole SMs, still all of them
    assert on == 1 and sum(sms) == 16 and min(sms) >= 1


def test_parallel_file_reader(engine_lib, tmp_path):
    """The input / key file reader of the prover (csrc/prover.cu read_range: slices of at least 8 MB read by up to
    eight threads with pread): every range comes back byte for byte, ranges beyond the end of the file fail."""
    import numpy as np
    f = engine_lib.b200msm_internal_read_range
    f.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_void_p]
    n = (70 << 20) + 12345
    data = np.random.default_rng(5).integers(0, 256, n, dtype=np.uint8)
    path = tmp_path / "blob"
    data.tofile(path)
    for off, ln in ((0, n), (0, 96), (n - 96, 96), (4097, (64 << 20) + 5), (1 << 20, 17 << 20), (123, 0), (9, (8 << 20) - 1)):
        out = np.full(ln + 8, 0xAB, np.uint8)
        assert f(str(path).encode(), off, ln, out.ctypes.data) == 0, (off, ln)
        assert (out[:ln] == data[off:off + ln]).all() and (out[ln:] == 0xAB).all(), (off, ln)
    out = np.zeros(32 << 20, np.uint8)
    assert f(str(path).encode(), n - 100, 200, out.ctypes.data) == 1                 # short file, single slice
    assert f(str(path).encode(), n - (20 << 20), 32 << 20, out.ctypes.data) == 1     # short file, several slices
    assert f(b"/nonexistent/blob", 0, 1, out.ctypes.data) == -1


def test_argument_checks_need_no_gpu(engine_lib):
    """Entry points reject null contexts / keys before they touch CUDA (error behaviour of the boundary: a status code,
    never a crash, never a point)."""
    vp = ctypes.c_void_p
    keys = (vp * 2)()
    assert engine_lib.b200msm_key_load_sharded_file(None, 2, b"/nonexistent", keys) == 1
    ctxs = (vp * 2)()                                            # two null contexts
    assert engine_lib.b200msm_key_load_sharded_file(ctxs, 2, b"/nonexistent", keys) == 1
    assert engine_lib.b200msm_set_lane_sms(None, 0, 8) == 1
    assert engine_lib.b200msm_msm(None, 0, 0, None, 0, None) == 1
    assert engine_lib.b200msm_wait(None, 0) == 1
    assert engine_lib.b200msm_prove(None, None, None, 0, None) == 1
    assert engine_lib.b200msm_proof_bytes(None) == 0 and engine_lib.b200msm_input_bytes(None) == 0
    assert engine_lib.b200msm_last_error(None) == b"null context"
    off, ln = ctypes.c_size_t(), ctypes.c_size_t()
    assert engine_lib.b200msm_shard_range(10, 0, 1, None, ctypes.byref(ln)) == 1
