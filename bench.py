#!/usr/bin/env python3
"""Benchmark of the MSM hot path (BASELINE.json metric: MSM points/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                  [--curve 0|1] [--group 1|2] [--log-n 20]

A *step* is one full multi-scalar multiplication sum_i s_i * P_i over n = 2^log_n points per GPU
(default: MNT4753 G1, 2^20 points = BASELINE.json configs[1]), through the C ABI of
include/b200_msm.h.  Bases are the synthetic structured set P0 + i*Q generated in HBM by the engine
(b200msm_bases_synthetic); scalars are uniform in [0, r).  At N > 1 (torchrun, one rank per GPU) the
MSM has N*n points sharded by point range: every rank runs its shard with no collective on the data
path, rank 0 then folds the N partial points (weak scaling).

  value     points/s with the scalars already resident in HBM when the timed region starts
  e2e       points/s through the same call with HOST (pinned) scalars: H2D of 96 B/point and D2H of
            the result point are inside the timed region
  roofline  k_accumulate (the dominant kernel) against the integer multiply-add peak measured on the
            same GPU in this run (bound "imad": this is carry-chained big-integer work, neither HBM-
            nor tensor-bound; the HBM view of the same kernel is reported next to it)
  cpu_baseline / --impl reference
            the reference's own CPU MSM (libff multi_exp_with_mixed_addition<BDLO12>, OpenMP, all host
            threads) from oracle/_ref/libref.so when it was built, else the plain-C port of it
            (oracle/liboracle.so), on a bounded 2^16-point sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "msm_points_per_s"
UNIT = "points/s"
CPU_SAMPLE_LOG_N = 16
CURVE_NAMES = {0: "MNT4753", 1: "MNT6753"}


def workload_name(curve, group, log_n):
    return "%s G%d MSM, 2^%d points per GPU" % (CURVE_NAMES[curve], group, log_n)


# ---- clocks ---------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ---- CPU baseline (test infrastructure used only as the thing timed / checked against) -----------------
def load_cpu_lib():
    from oracle import pyoracle as po
    ref = po.load_reference()
    return (ref if ref is not None else po.load_oracle()), po


def cpu_msm_time(lib, curve, group, bases, scalars):
    out, t = lib.msm(curve, group, bases, scalars, method=1, chunks=0, prefilter=1)
    return out, t


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from gpu_groth16_prover_3x_b200 import synthetic
    lib, po = load_cpu_lib()
    n = 1 << CPU_SAMPLE_LOG_N
    orc = po.load_oracle()
    bases = orc.gen_bases(args.curve, args.group, n)
    sets = [synthetic.random_scalars(args.curve, n, 1000 + i) for i in range(2)]
    for i in range(args.warmup):
        cpu_msm_time(lib, args.curve, args.group, bases, sets[i & 1])
    t0 = time.perf_counter()
    for i in range(args.steps):
        cpu_msm_time(lib, args.curve, args.group, bases, sets[i & 1])
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = "2^%d-point MSM per step (bounded sample of the 2^%d workload), libff BDLO12 + OpenMP" % (CPU_SAMPLE_LOG_N, args.log_n)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(args.curve, args.group, args.log_n), "curve": CURVE_NAMES[args.curve],
                   "group": "G%d" % args.group, "points_per_gpu": 1 << args.log_n, "sample_points": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": lib.num_threads(), "kind": lib.kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---- our arm -------------------------------------------------------------------------------------------
def measured_hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def run_b200(args):
    import torch
    import torch.distributed as dist
    import gpu_groth16_prover_3x_b200 as pkg
    from gpu_groth16_prover_3x_b200 import sharding, synthetic

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("--gpus %d needs torchrun with --nproc-per-node %d" % (args.gpus, args.gpus))
        args.gpus = world
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    curve, group, n = args.curve, args.group, 1 << args.log_n
    deg = pkg.degree(curve, group)
    k_tower = {1: 1, 2: 3, 3: 6}[deg]
    r = synthetic.fr_modulus(curve)
    ctx = pkg.MsmContext(curve, local)
    if args.window_bits:
        ctx.set_window_bits(args.window_bits)

    # this rank's shard of the (world * n)-point MSM: bases P0 + (rank*n + i) * Q
    k0, k1 = synthetic.base_seed_scalars(curve)
    R = synthetic.R
    k0p, k1p = (int.from_bytes(k.tobytes(), "little") * pow(R, -1, r) % r for k in (k0, k1))
    k0_rank = synthetic.int_to_limbs((k0p + rank * n * k1p) % r * R % r)
    t0 = time.perf_counter()
    slot = ctx.synthetic_bases(group, n, k0_rank, k1)
    t_bases = time.perf_counter() - t0
    binfo = ctx.bases_info(slot)

    host_sets = [torch.from_numpy(synthetic.random_scalars(curve, n, 100 + 2 * rank + i).view(np.int64)).pin_memory() for i in range(2)]
    dev_sets = [h.cuda(non_blocking=False) for h in host_sets]

    stream = torch.cuda.Stream()
    ctx.set_stream(0, stream.cuda_stream)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def one_step(i, sets):
        return ctx.msm(slot, sets[i & 1], n)

    def timed(sets, steps, sampler=None):
        """-> (seconds by CUDA events on the launching stream, wall seconds, per-step phase timings, last result)."""
        for i in range(args.warmup):
            one_step(i, sets)
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        phases = []
        w0 = time.perf_counter()
        with torch.cuda.stream(stream):
            e0.record(stream)
            for i in range(steps):
                out = one_step(i, sets)
                phases.append(ctx.last_timings())
            e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - w0
        clocks = sampler.stop() if sampler else None
        barrier()
        return e0.elapsed_time(e1) / 1e3, wall, phases, out, clocks

    sampler = ClockSampler(local)
    dev_s, dev_wall, phases, out_dev, clocks = timed(dev_sets, args.steps, sampler)
    e2e_s, e2e_wall, _, out_host, _ = timed(host_sets, args.steps)

    def max_over_ranks(x):
        return sharding.max_over_ranks(x, device="cuda")

    dev_s, e2e_s = max_over_ranks(dev_s), max_over_ranks(e2e_s)
    dev_wall, e2e_wall = max_over_ranks(dev_wall), max_over_ranks(e2e_wall)

    # fold of the per-rank partial points (outside the timed region of the shards: it is 864 B per rank)
    fold_ms = None
    if world > 1:
        partials = sharding.gather_partials(out_dev, device="cuda")
        if rank == 0:
            t0 = time.perf_counter()
            ctx.fold(group, partials)
            fold_ms = (time.perf_counter() - t0) * 1e3

    if rank != 0:
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return

    total_points = n * world
    value = total_points * args.steps / dev_s
    e2e = total_points * args.steps / e2e_s
    info = phases[-1]
    acc_ms = statistics.mean(p["accumulate"] for p in phases)
    W = info["windows"]
    # executed work of the accumulation phase: one addition per sorted entry (minus one per non-empty bucket,
    # neglected); 6 field multiplications per batched-affine addition, 11 per Jacobian mixed addition
    # (SURVEY.md 8(d)); a field multiplication in the degree-k tower is k_tower Fq products of 1176 MAC
    muls_per_add = 6 if info.get("accumulator", 0) == 0 else 11
    macs_per_launch = float(n) * W * muls_per_add * k_tower * 1176
    achieved = macs_per_launch / (acc_ms * 1e-3) / 1e9
    micro = {"imad_wide_gmacs": ctx.microbench(0, 4096), "imad_lo_gops": ctx.microbench(1, 4096),
             "fq_modmul_gmuls": ctx.microbench(2, 2048)}
    peak = max(micro["imad_wide_gmacs"], micro["fq_modmul_gmuls"] * 1176)
    hbm_peak, hbm_src = measured_hbm_peak()
    aff_bytes = 2 * deg * 96
    if muls_per_add == 6:
        # batched-affine addition: forward reads x1, x2 and parks the prefix; backward reads both points and the
        # prefix and writes the sum; the pair descriptor (16 B) is read in both passes
        alg_bytes = float(n) * W * (2 * 96 * deg + 96 * deg + 2 * aff_bytes + 96 * deg + aff_bytes + 32 + 2)
    else:
        alg_bytes = float(n) * W * (aff_bytes + 4)
    traffic = None
    try:
        if (curve, group, args.log_n) == (0, 1, 20) and muls_per_add == 6:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r01_k_batch_add_round0_traffic.json")))
            traffic = {"bytes": tj["traffic_bytes"], "launch": tj["kernel"], "source": tj["source"],
                       "algorithmic_bytes_of_that_launch": float(n) * W / 2 * (alg_bytes / (float(n) * W))}
    except Exception:
        traffic = None
    # SURVEY.md 8(d) canonical count (c = 16, W = 48, Jacobian mixed add) applied to the whole step: what the
    # reference-style algorithm would have to execute to deliver the same points/s
    canonical_macs_per_point = 1176.0 * 11 * k_tower * 48
    roofline = {
        "kernel": "k_batch_add (all rounds of one MSM)" if muls_per_add == 6 else "k_accumulate", "bound": "imad",
        "field_muls_per_addition": muls_per_add, "achieved": achieved, "peak": peak, "unit": "GMAC/s", "frac": achieved / peak,
        "traffic": traffic, "launch_ms": acc_ms, "share_of_step": acc_ms / (dev_s / args.steps * 1e3),
        "peak_source": "measured in this run: max(IMAD.WIDE.U32 stream, Fq Montgomery product in registers x 1176 MAC); "
                       "a 32x32->64 multiply-add issues at 32 per clock per SM on B200 (tools/pipe_probe.cu)",
        "macs_per_launch": macs_per_launch, "micro": micro,
        "canonical_c16": {"macs_per_point": canonical_macs_per_point,
                          "step_gmacs": total_points / world * args.steps / dev_s * canonical_macs_per_point / 1e9,
                          "frac_of_peak": total_points / world * args.steps / dev_s * canonical_macs_per_point / 1e9 / peak},
        "hbm": {"algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes / (acc_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                "frac": alg_bytes / (acc_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src},
    }

    cpu_baseline = None
    parity = None
    if world == 1 and not args.no_cpu:
        lib, po = load_cpu_lib()
        ns = min(n, 1 << CPU_SAMPLE_LOG_N)
        sb = ctx.download_bases(slot, 0, ns)
        ss = host_sets[0].numpy().view(np.uint64)[:ns * 12].copy()
        want, t_cpu = cpu_msm_time(lib, curve, group, sb, ss)
        got = ctx.to_affine(group, ctx.msm(slot, ss, ns))
        parity = bool((got == want).all())
        cpu_baseline = {"value": ns / t_cpu, "unit": UNIT, "cores": lib.num_threads(), "kind": lib.kind,
                        "sample": "one 2^%d-point MSM on the first points of the same workload (libff BDLO12 + OpenMP); "
                                  "GPU result on the sample bit-equal: %s" % (CPU_SAMPLE_LOG_N, parity)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": workload_name(curve, group, args.log_n), "curve": CURVE_NAMES[curve], "group": "G%d" % group,
                   "points_per_gpu": n, "total_points": total_points, "window_bits": info["window_bits"], "windows": W,
                   "bucket_sets": info["bucket_sets"], "window_tables": info["tables"], "table_bytes_per_gpu": binfo["bytes"],
                   "table_build_s": binfo["table_build_ms"] / 1e3,
                   "sharding": "point-range, one partial point per GPU, no collective on the data path",
                   "cache": "inputs larger than L2 (bases %.0f MB + scalars %.0f MB + sorted list %.0f MB per step)" % (
                       binfo["bytes"] / 1e6, n * 96 / 1e6, n * W * 4 / 1e6),
                   "timing": "CUDA events on the launching stream around the K steps, max over ranks",
                   "bases_generation_s": t_bases},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": n * 96, "d2h_bytes_per_step": 36 * deg * 8,
                "ms_per_step": e2e_s / args.steps * 1e3},
        "gpu_launches": int(sum(p["kernel_launches"] for p in phases)),
        "wall_ms_per_step": dev_wall / args.steps * 1e3,
        "phases_ms": {k: statistics.mean(p[k] for p in phases) for k in pkg.MsmContext.PHASES},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks,
    }
    if fold_ms is not None:
        line["fold_ms"] = fold_ms
    if world == 1 and not args.no_secondary and (curve, group) == (0, 1):
        line["secondary"] = secondary_workload(args, local)
    if world == 1 and not args.no_cpu and not args.no_secondary:
        try:
            line["proof_latency"] = proof_latency()
        except Exception as e:  # the headline line must not depend on the optional leg
            line["proof_latency"] = {"error": str(e)[:200]}
    emit(line)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def secondary_workload(args, device):
    """BASELINE.json's metric names two workloads; the line's `value` is MNT4753 G1 (configs[1]).  This measures
    the other one, MNT6753 G2 over the Fq3 twist (configs[2]), the same way (device-resident scalars, CUDA events
    on the launching stream), with 3 warm-up and 3 timed MSMs, and reports it next to the headline."""
    import torch
    import gpu_groth16_prover_3x_b200 as pkg
    from gpu_groth16_prover_3x_b200 import synthetic
    curve, group, n = 1, 2, 1 << args.log_n
    ctx = pkg.MsmContext(curve, device)
    try:
        k0, k1 = synthetic.base_seed_scalars(curve)
        t0 = time.perf_counter()
        slot = ctx.synthetic_bases(group, n, k0, k1)
        t_bases = time.perf_counter() - t0
        binfo = ctx.bases_info(slot)
        dev = [torch.from_numpy(synthetic.random_scalars(curve, n, 300 + i).view(np.int64)).cuda() for i in range(2)]
        stream = torch.cuda.Stream()
        ctx.set_stream(0, stream.cuda_stream)
        for i in range(3):
            ctx.msm(slot, dev[i & 1], n)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        steps = 3
        with torch.cuda.stream(stream):
            e0.record(stream)
            for i in range(steps):
                ctx.msm(slot, dev[i & 1], n)
            e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        t = ctx.last_timings()
        return {"workload": workload_name(curve, group, args.log_n), "value": n / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                "steps": steps, "warmup": 3, "window_bits": t["window_bits"], "windows": t["windows"], "window_tables": t["tables"],
                "table_bytes": binfo["bytes"], "table_build_s": binfo["table_build_ms"] / 1e3, "bases_generation_s": t_bases,
                "phases_ms": {k: t[k] for k in pkg.MsmContext.PHASES}}
    finally:
        ctx.close()


def proof_latency():
    """BASELINE.json's second headline, end-to-end proof latency, on the reference's `generate_parameters fast`
    instance (MNT4753 d = 2^14 - 1, MNT6753 d = 2^10 - 1; the default 2^20 instance takes minutes of CPU to
    generate, see tools/full_proof.sh and DESIGN.md for that run).  Reference arm: the reference's own CPU prover
    `main <curve> compute` (oracle/_ref, test infrastructure built where /root/reference exists); ours: the
    product's command-line prover b200_prove (b200msm_key_load_file + b200msm_prove), third proof of a resident
    process.  Returns None when the reference binaries are not there."""
    import hashlib
    import re
    import tempfile
    ref = os.path.join(ROOT, "oracle", "_ref")
    bins = [os.path.join(ref, "generate_parameters"), os.path.join(ref, "main"),
            os.path.join(ROOT, "gpu_groth16_prover_3x_b200", "b200_prove")]
    if not all(os.path.exists(b) for b in bins):
        return None
    out = {"instance": "generate_parameters fast", "curves": {}}
    with tempfile.TemporaryDirectory() as d:
        subprocess.run([bins[0], "fast"], cwd=d, check=True, stdout=subprocess.DEVNULL, timeout=900)
        for curve in ("MNT4753", "MNT6753"):
            prm, inp = curve + "-parameters", curve + "-input"
            t0 = time.perf_counter()
            log = subprocess.run([bins[1], curve, "compute", prm, inp, "out-ref"], cwd=d, check=True, capture_output=True,
                                 text=True, timeout=1800).stdout
            t_ref_wall = time.perf_counter() - t0
            log2 = subprocess.run([bins[2], curve, "compute", prm, inp, "out-b200", "3"], cwd=d, check=True,
                                  capture_output=True, text=True, timeout=900).stdout
            ours = [float(x) for x in re.findall(r"Total time from input to output: ([0-9.]+) ms", log2)]
            total = re.findall(r"Total runtime \(incl. key load\): ([0-9.]+) ms", log2)
            upload = re.findall(r"key load \+ window tables: ([0-9.]+) ms", log2)
            same = hashlib.sha256(open(os.path.join(d, "out-ref"), "rb").read()).digest() == \
                hashlib.sha256(open(os.path.join(d, "out-b200"), "rb").read()).digest()
            out["curves"][curve] = {"b200_input_to_proof_s": ours[-1] / 1e3 if ours else None,
                                    "b200_first_proof_s": ours[0] / 1e3 if ours else None,
                                    "b200_key_upload_and_tables_s": float(upload[0]) / 1e3 if upload else None,
                                    "b200_process_total_3_proofs_s": float(total[0]) / 1e3 if total else None,
                                    "reference_cpu_process_wall_s": t_ref_wall,
                                    "proof_sha256_equal": bool(same)}
    return out


_REAL_STDOUT = None


def quiet_stdout():
    """Everything libraries print (NCCL's version banner, torchrun chatter) goes to stderr; the one JSON line
    of the contract is written to the real stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--curve", type=int, default=0, choices=[0, 1])
    ap.add_argument("--group", type=int, default=1, choices=[1, 2])
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--window-bits", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the MNT6753 G2 measurement reported next to the headline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
