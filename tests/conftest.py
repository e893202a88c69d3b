import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run on the GPU box with -m gpu)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle as po
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        env = dict(os.environ, CC="/usr/bin/gcc")
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"], check=True, env=env)
    return po.load_oracle()


@pytest.fixture(scope="session")
def golden():
    return {name: np.load(os.path.join(GOLD, name + ".npz")) for name in
            ("field_vectors", "point_vectors", "msm_vectors", "generators", "h_vectors")}


@pytest.fixture(scope="session")
def emu():
    """The device field/curve templates compiled for the host (tests/host_emu)."""
    import ctypes
    d = os.path.join(ROOT, "tests", "host_emu")
    extra = os.environ.get("EMU_CFLAGS", "").split()   # e.g. -DMNT753_MUL_ROLL=8 to emulate a multiplier variant
    so = os.path.join(d, "libemu%s.so" % ("_" + "".join(c for c in "".join(extra) if c.isalnum()) if extra else ""))
    srcs = [os.path.join(d, "emu.cpp")] + [os.path.join(ROOT, "gpu_groth16_prover_3x_b200", "csrc", f)
                                            for f in ("prim.cuh", "fq.cuh", "fe.cuh", "ec.cuh", "curves.cuh", "batch_affine.cuh", "glv_split.cuh")]
    srcs += [os.path.join(ROOT, "tools", "experiments", f) for f in ("fq_fp64.cuh", "fq_experiments.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-DMNT753_HOST_EMU"] + extra + ["-x", "c++",
                        srcs[0], "-o", so], check=True)
    lib = ctypes.CDLL(so)
    u64p = ctypes.POINTER(ctypes.c_uint64)
    lib.emu_field_op.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, u64p, u64p, u64p]
    lib.emu_point_op.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, u64p, u64p, ctypes.c_int, u64p]
    lib.emu_fq_mul.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_size_t, u64p, u64p, u64p]
    lib.emu_affine_add.argtypes = [ctypes.c_int, ctypes.c_int, u64p, u64p, ctypes.c_int, u64p]
    lib.emu_field_inv.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_size_t, u64p, u64p]
    lib.emu_psi.argtypes = [ctypes.c_int, u64p, u64p]
    lib.emu_batch_add_generic.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_size_t, u64p, u64p, ctypes.POINTER(ctypes.c_int), u64p]
    lib.emu_fq_inv.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_size_t, u64p, u64p]
    lib.emu_fr_from_mont.argtypes = [ctypes.c_int, ctypes.c_size_t, u64p, u64p]
    return lib


@pytest.fixture(scope="session")
def warp_emu():
    """The warp-cooperative device code (csrc/fq_inv_coop.cuh) on a simulated warp: 32 lanes in lock step with
    shuffles and ballots (tests/host_emu/warp_sim.hpp)."""
    import ctypes
    d = os.path.join(ROOT, "tests", "host_emu")
    so = os.path.join(d, "libwarpemu.so")
    srcs = [os.path.join(d, "warp_emu.cpp"), os.path.join(d, "warp_sim.hpp")] + [
        os.path.join(ROOT, "gpu_groth16_prover_3x_b200", "csrc", f) for f in ("prim.cuh", "fq.cuh", "fq_inv_coop.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-DMNT753_HOST_EMU", "-x", "c++", srcs[0], "-o", so], check=True)
    lib = ctypes.CDLL(so)
    u64p = ctypes.POINTER(ctypes.c_uint64)
    lib.emu_fq_inv_coop.argtypes = [ctypes.c_int, ctypes.c_size_t, u64p, u64p]
    return lib


@pytest.fixture(scope="session")
def plan_emu():
    """The round planning of the bucket accumulation (csrc/ba_plan.cuh) on a simulated warp, with set unions for
    additions (tests/host_emu/plan_emu.cpp)."""
    import ctypes
    d = os.path.join(ROOT, "tests", "host_emu")
    so = os.path.join(d, "libplanemu.so")
    srcs = [os.path.join(d, "plan_emu.cpp"), os.path.join(d, "warp_sim.hpp")] + [os.path.join(ROOT, "gpu_groth16_prover_3x_b200", "csrc", f)
                                                                                 for f in ("ba_plan.cuh", "tree_plan.cuh", "recode.cuh", "glv_split.cuh", "fq.cuh", "prim.cuh", "mnt753_constants.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-DMNT753_HOST_EMU", "-x", "c++", srcs[0], "-o", so], check=True)
    lib = ctypes.CDLL(so)
    lib.emu_plan_check.argtypes = [ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32), ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32,
                                   ctypes.POINTER(ctypes.c_uint64), ctypes.c_char_p, ctypes.c_size_t]
    lib.emu_recode.argtypes = [ctypes.POINTER(ctypes.c_uint32), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int32)]
    lib.emu_recode.restype = None
    lib.emu_glv_split.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint32)]
    lib.emu_glv_split.restype = None
    lib.emu_tree_check.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint64), ctypes.c_char_p, ctypes.c_size_t]
    return lib


@pytest.fixture(scope="session")
def fft_emu():
    """The H-polynomial kernels (csrc/fft_kernels.cuh) on a simulated thread block (tests/host_emu/block_sim.hpp)."""
    import ctypes
    d = os.path.join(ROOT, "tests", "host_emu")
    so = os.path.join(d, "libfftemu.so")
    srcs = [os.path.join(d, "fft_emu.cpp"), os.path.join(d, "block_sim.hpp")] + [os.path.join(ROOT, "gpu_groth16_prover_3x_b200", "csrc", f)
                                                                                 for f in ("fft_kernels.cuh", "fq.cuh", "prim.cuh", "mnt753_constants.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-DMNT753_HOST_EMU", "-x", "c++", srcs[0], "-o", so], check=True)
    lib = ctypes.CDLL(so)
    u64p = ctypes.POINTER(ctypes.c_uint64)
    lib.emu_compute_h.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, u64p, u64p, u64p, u64p]
    return lib


@pytest.fixture(scope="session")
def sort_emu():
    """The recode / counting-sort kernels (csrc/sort_kernels.cuh) on a simulated thread block."""
    import ctypes
    d = os.path.join(ROOT, "tests", "host_emu")
    so = os.path.join(d, "libsortemu.so")
    srcs = [os.path.join(d, "sort_emu.cpp"), os.path.join(d, "block_sim.hpp")] + [os.path.join(ROOT, "gpu_groth16_prover_3x_b200", "csrc", f)
                                                                                  for f in ("sort_kernels.cuh", "recode.cuh", "fq.cuh", "prim.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.run(["/usr/bin/g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-DMNT753_HOST_EMU", "-x", "c++", srcs[0], "-o", so], check=True)
    lib = ctypes.CDLL(so)
    u32p = ctypes.POINTER(ctypes.c_uint32)
    lib.emu_sort.argtypes = [ctypes.c_uint32, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint32, ctypes.c_int, ctypes.c_int, u32p,
                             ctypes.POINTER(ctypes.c_uint8), u32p, u32p, u32p]
    return lib


@pytest.fixture(scope="session")
def engine_lib():
    import gpu_groth16_prover_3x_b200 as pkg
    return pkg.load_library()
