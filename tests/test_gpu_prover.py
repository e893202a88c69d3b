"""End-to-end acceptance test of the reference's README: the proof written by the GPU prover must have the
same sha256 as the one written by the reference's CPU prover (`main <curve> compute`) on the same freshly
generated parameters and input.  Here the "GPU prover" is the reference's own driver with its five MSMs routed
through libb200msm.so (tests/integration/b200_prover.cpp, built into oracle/_ref/ by `make -C oracle prover`
where /root/reference exists; the binaries travel to the GPU box with the snapshot)."""
import hashlib
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
BINS = [os.path.join(REF, b) for b in ("generate_parameters", "main", "b200_prover")]
BUNDLE = os.path.join(REF, "b200_bundle_prover")


def sha256(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


@pytest.fixture(scope="module")
def fast_params(tmp_path_factory):
    if not all(os.path.exists(b) for b in BINS):
        pytest.skip("reference binaries not built (oracle/_ref): run __graft_entry__.build() where /root/reference exists")
    d = tmp_path_factory.mktemp("groth16")
    # `generate_parameters fast`: MNT4753 d = 2^14 - 1, MNT6753 d = 2^10 - 1 (generate_parameters.cpp:110-135)
    subprocess.run([BINS[0], "fast"], cwd=d, check=True, stdout=subprocess.DEVNULL, timeout=900)
    return d


@pytest.mark.parametrize("curve", ["MNT4753", "MNT6753"])
def test_proof_sha256_equals_reference_cpu_prover(fast_params, curve):
    """MSMs and the H polynomial on the device (default), then the same with the reference's CPU FFTs."""
    d = fast_params
    params, inp = "%s-parameters" % curve, "%s-input" % curve
    ref_out = os.path.join(d, curve + "-output-ref")
    if not os.path.exists(ref_out):
        subprocess.run([BINS[1], curve, "compute", params, inp, curve + "-output-ref"], cwd=d, check=True,
                       stdout=subprocess.DEVNULL, timeout=1800)
    for extra in ([], ["1", "cpu-h"]):
        name = curve + "-output-b200" + ("-cpuh" if extra else "")
        out = subprocess.run([BINS[2], curve, "compute", params, inp, name] + extra, cwd=d, check=True,
                             capture_output=True, text=True, timeout=900).stdout
        print(out)
        b = os.path.join(d, name)
        assert os.path.getsize(ref_out) == os.path.getsize(b) == (768 if curve == "MNT4753" else 960)
        assert sha256(ref_out) == sha256(b), extra


@pytest.mark.parametrize("curve", ["MNT4753", "MNT6753"])
def test_bundle_multiexp_adaptor(fast_params, curve):
    """the reference's prover call sequence written against the plugin bundle only
    (tests/integration/b200_bundle_prover.cpp), with B = b200_bundle<...> whose B::multiexp_G1 / multiexp_G2
    (prover_reference_functions.cpp:350-368,690-708) run on the engine: libff projective vectors marshalled to the
    affine wire format (Z == 0 -> y = 0), base sets cached per vector, results imported by B::read_pt_ECp/ECpe.
    All five multiexps of a proof go through the two bundle functions; the proof must equal `main`'s byte for byte,
    on one GPU, over two and three point-range shards (one per GPU where the box has them), and on a second proof
    (cache hits only)."""
    import re
    if not os.path.exists(BUNDLE):
        pytest.skip("oracle/_ref/b200_bundle_prover not built")
    d = fast_params
    params, inp = "%s-parameters" % curve, "%s-input" % curve
    ref_out = os.path.join(d, curve + "-output-ref")
    if not os.path.exists(ref_out):
        subprocess.run([BINS[1], curve, "compute", params, inp, curve + "-output-ref"], cwd=d, check=True,
                       stdout=subprocess.DEVNULL, timeout=1800)
    for gpus in (1, 2, 3):
        name = curve + "-output-bundle%d" % gpus
        out = subprocess.run([BUNDLE, curve, "compute", params, inp, name, str(gpus), "b200", "2"], cwd=d, check=True,
                             capture_output=True, text=True, timeout=900).stdout
        print(out)
        assert sha256(ref_out) == sha256(os.path.join(d, name)), gpus
        m = re.search(r"base-set uploads: (\d+), cache hits: (\d+)", out)
        assert m and (int(m.group(1)), int(m.group(2))) == (5, 5)   # second proof: no upload, no table build


@pytest.mark.parametrize("curve", ["MNT4753", "MNT6753"])
def test_whole_proof_through_the_c_abi(fast_params, curve):
    """b200msm_key_load_file + b200msm_prove (key resident in HBM, one call per proof, proof assembly and affine
    normalisation on the device, no libff anywhere) and the b200_prove command-line twin of the reference's
    `cuda_prover_piecewise <curve> compute`: same bytes as the reference CPU prover."""
    import gpu_groth16_prover_3x_b200 as pkg
    d = fast_params
    params, inp = os.path.join(d, "%s-parameters" % curve), os.path.join(d, "%s-input" % curve)
    ref_out = os.path.join(d, curve + "-output-ref")
    if not os.path.exists(ref_out):
        subprocess.run([BINS[1], curve, "compute", params, inp, ref_out], cwd=d, check=True, stdout=subprocess.DEVNULL, timeout=1800)
    want = open(ref_out, "rb").read()
    with pkg.MsmContext(pkg.MNT4753 if curve == "MNT4753" else pkg.MNT6753, 0) as ctx:
        key = ctx.load_key(params)
        info = ctx.key_info(key)
        assert info["m"] == info["d"] + 1
        image = open(inp, "rb").read()
        assert ctx.prove(key, image) == want
        assert ctx.prove(key, image) == want           # a resident key proves again
        with pytest.raises(pkg.MsmError):
            ctx.prove(key, image[:-96])                # truncated witness
        assert ctx.prove_file(key, inp) == want        # input read from the file while the witness MSMs run
        with pytest.raises(pkg.MsmError):
            ctx.prove_file(key, params)                # a file of the wrong size
        ctx.free_key(key)
        with pytest.raises(pkg.MsmError):
            ctx.load_key(open(params, "rb").read()[:-8])
    cli = os.path.join(ROOT, "gpu_groth16_prover_3x_b200", "b200_prove")
    out = subprocess.run([cli, curve, "compute", params, inp, os.path.join(d, curve + "-output-cli"), "2"], check=True,
                         capture_output=True, text=True, timeout=900).stdout
    print(out)
    assert sha256(os.path.join(d, curve + "-output-cli")) == hashlib.sha256(want).hexdigest()
    # the witness MSMs one after the other on the whole GPU (0) and side by side on disjoint SMs (2: always): same proof
    for mode in ("0", "2"):
        name = os.path.join(d, curve + "-output-cli-split" + mode)
        subprocess.run([cli, curve, "compute", params, inp, name, "2"], check=True, capture_output=True, text=True, timeout=900,
                       env=dict(os.environ, B200MSM_LANE_SPLIT=mode))
        assert sha256(name) == hashlib.sha256(want).hexdigest(), mode


@pytest.mark.parametrize("nshards", [2, 3])
def test_sharded_proof_through_the_c_abi(fast_params, nshards):
    """b200msm_key_load_shard + b200msm_prove_sharded[_file]: every query split by point range over `nshards` contexts
    (one per GPU where the box has that many, else several on one GPU -- the sharding logic is the same), H on shard 0,
    its coefficients handed on by peer copy, partial points folded: same proof bytes as the reference CPU prover."""
    import torch
    import gpu_groth16_prover_3x_b200 as pkg
    d = fast_params
    for curve in ("MNT4753", "MNT6753"):
        params, inp = os.path.join(d, "%s-parameters" % curve), os.path.join(d, "%s-input" % curve)
        ref_out = os.path.join(d, curve + "-output-ref")
        if not os.path.exists(ref_out):
            subprocess.run([BINS[1], curve, "compute", params, inp, ref_out], cwd=d, check=True, stdout=subprocess.DEVNULL, timeout=1800)
        want = open(ref_out, "rb").read()
        ndev = torch.cuda.device_count()
        cid = pkg.MNT4753 if curve == "MNT4753" else pkg.MNT6753
        ctxs = [pkg.MsmContext(cid, g % ndev) for g in range(nshards)]
        try:
            image = open(params, "rb").read()
            keys = [c.load_key_shard(image, g, nshards) for g, c in enumerate(ctxs)]
            assert pkg.prove_sharded(ctxs, keys, open(inp, "rb").read()) == want
            with pytest.raises(pkg.MsmError):
                pkg.prove_sharded(ctxs[::-1], keys[::-1], open(inp, "rb").read())   # shards out of order
            for c, k in zip(ctxs, keys):
                c.free_key(k)
            # the same shards straight from the parameter FILE (b200msm_key_load_sharded_file: one host thread per shard)
            keys = pkg.load_key_sharded_file(ctxs, params)
            assert pkg.prove_sharded(ctxs, keys, open(inp, "rb").read()) == want
            for c, k in zip(ctxs, keys):
                c.free_key(k)
            with pytest.raises(pkg.MsmError):
                pkg.load_key_sharded_file(ctxs, inp)                                  # not a parameter file
        finally:
            for c in ctxs:
                c.close()
    if torch.cuda.device_count() >= nshards:      # the command-line prover on real devices
        curve = "MNT4753"
        cli = os.path.join(ROOT, "gpu_groth16_prover_3x_b200", "b200_prove")
        out = os.path.join(d, curve + "-output-cli%d" % nshards)
        subprocess.run([cli, curve, "compute", os.path.join(d, curve + "-parameters"), os.path.join(d, curve + "-input"), out, "1", str(nshards)],
                       check=True, stdout=subprocess.DEVNULL, timeout=900)
        assert sha256(out) == sha256(os.path.join(d, curve + "-output-ref"))


def test_sharded_prover_two_gpus(fast_params):
    """Every query sharded by point range over two GPUs, partial points folded: same proof."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    d = fast_params
    curve = "MNT4753"
    params, inp = "%s-parameters" % curve, "%s-input" % curve
    ref_out = os.path.join(d, curve + "-output-ref")
    if not os.path.exists(ref_out):
        subprocess.run([BINS[1], curve, "compute", params, inp, curve + "-output-ref"], cwd=d, check=True, stdout=subprocess.DEVNULL, timeout=1800)
    subprocess.run([BINS[2], curve, "compute", params, inp, curve + "-output-2gpu", "2", "gpu-h", "1"], cwd=d, check=True,
                   stdout=subprocess.DEVNULL, timeout=900)
    assert sha256(ref_out) == sha256(os.path.join(d, curve + "-output-2gpu"))

