"""Compile csrc/*.cu into libb200msm.so for sm_100a (nvcc cross-compiles without a GPU).

The four groups (MNT4753/MNT6753 x G1/G2) are instantiated in their own translation units
(csrc/inst_*.cu) and compiled in parallel; msm.cu holds the C ABI.
"""
import concurrent.futures
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libb200msm.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC"]
HOST_CXX = "/usr/bin/g++"  # the image's default CXX is a relocated wrapper; use the system compiler


def units():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def sources():
    out = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith((".cu", ".cuh", ".h", ".cpp"))]
    out.append(os.path.join(os.path.dirname(HERE), "include", "b200_msm.h"))
    return out


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(s) <= t for s in sources())


def build(force=False, verbose=False, extra_flags=None, out=None):
    """Build the CUDA library in-tree; returns its path.  extra_flags / out build a VARIANT of the same
    library next to it (development A/B runs: e.g. extra_flags=["-DMNT753_MUL_ROLL=8"])."""
    global LIB, OBJ
    if extra_flags or out:
        LIB = out or LIB
        OBJ = OBJ + "_variant"
        force = True
    if not force and up_to_date():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libb200msm.so")
    os.makedirs(OBJ, exist_ok=True)
    ccbin = ["-ccbin", HOST_CXX] if os.path.exists(HOST_CXX) else []
    extra = (["-Xptxas", "-v"] if verbose else []) + list(extra_flags or [])

    def compile_one(src):
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        r = subprocess.run([nvcc] + ccbin + NVCC_FLAGS + extra + ["-c", "-o", obj, src], cwd=CSRC, capture_output=True, text=True)
        return src, obj, r

    objs = []
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for src, obj, r in ex.map(compile_one, units()):
            if verbose or r.returncode:
                print(r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError("nvcc failed on %s" % src)
            objs.append(obj)
    subprocess.run([nvcc] + ccbin + ARCH + ["-shared", "-o", LIB + ".tmp"] + objs, check=True, cwd=CSRC)
    os.replace(LIB + ".tmp", LIB)
    if LIB.endswith("libb200msm.so") and os.path.dirname(LIB) == HERE:
        # the command-line prover: plain C++ over the C ABI (no CUDA headers, no libff)
        cxx = HOST_CXX if os.path.exists(HOST_CXX) else "g++"
        subprocess.run([cxx, "-O2", "-std=c++17", "-o", os.path.join(HERE, "b200_prove"), os.path.join(CSRC, "prove_main.cpp"),
                        "-L" + HERE, "-lb200msm", "-lpthread", "-Wl,-rpath,$ORIGIN"], check=True)
    return LIB


if __name__ == "__main__":
    import sys
    defs = [a for a in sys.argv[1:] if a.startswith("-D")]
    outs = [a[len("--out="):] for a in sys.argv[1:] if a.startswith("--out=")]
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, extra_flags=defs, out=outs[0] if outs else None))
