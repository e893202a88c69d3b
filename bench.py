#!/usr/bin/env python3
"""Benchmark of the MSM hot path (BASELINE.json metric: MSM points/s; second headline: proof latency).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                  [--curve 0|1] [--group 1|2] [--log-n 20] [--proof default|fast|none] [--quick]

A *step* is one full multi-scalar multiplication sum_i s_i * P_i over n = 2^log_n points per GPU
(default: MNT4753 G1, 2^20 points = BASELINE.json configs[1]), through the C ABI of
include/b200_msm.h.  Bases are the synthetic structured set P0 + i*Q generated in HBM by the engine
(b200msm_bases_synthetic); scalars are uniform in [0, r).  At N > 1 (torchrun, one rank per GPU) the
line's `value` is WEAK scaling: an (N*n)-point MSM sharded by point range, every rank runs its n-point
shard with no collective on the data path.

  value     points/s with the scalars already resident in HBM when the timed region starts
  e2e       points/s through the same call with HOST (pinned) scalars: H2D of 96 B/point and D2H of
            the result point are inside the timed region -- this is the metric as SURVEY.md 8(d) defines it
  strong    STRONG scaling at the same N: ONE fixed 2^log_n-point MSM split N ways by b200msm_shard_range,
            gather of the N partial points and the fold INSIDE the timed region, result checked against the
            closed form (sum s_i) P0 + (sum i s_i) Q; efficiency_vs_1 = t(1 GPU) / (N * t(N GPUs))
  proof     whole proofs at N GPUs through the product's command-line prover b200_prove
            (b200msm_key_load_sharded_file + b200msm_prove_sharded_file): default-size instances of both curves
            (MNT4753 d = 2^20 - 1, MNT6753 d = 2^15 - 1), sha256 against the reference CPU prover `main`
  roofline  k_batch_add (the dominant kernel) against the integer multiply-add peak measured on the
            same GPU in this run (bound "imad": this is carry-chained big-integer work, neither HBM-
            nor tensor-bound; the HBM view of the same kernel is reported next to it)
  cpu_baseline / --impl reference
            the reference's own CPU MSM (libff multi_exp_with_mixed_addition<BDLO12>, OpenMP on ALL host
            cores whatever OMP_NUM_THREADS torchrun exported) from oracle/_ref/libref.so when it was built,
            else the plain-C port of it (oracle/liboracle.so), on the SAME 2^log_n-point workload (a smaller
            sample only where one CPU MSM would exceed ~40 s, stated in `sample`).
"""
import argparse
import hashlib
import json
import os
import re
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
CURVE_NAMES = {0: "MNT4753", 1: "MNT6753"}
CPU_STEP_BUDGET_S = 40.0      # one CPU MSM longer than this is replaced by a smaller sample (stated in `sample`)
REF_ARM_BUDGET_S = 150.0      # --impl reference: total time spent in timed CPU steps
DEFAULT_D = {"MNT4753": (1 << 20) - 1, "MNT6753": (1 << 15) - 1}   # generate_parameters.cpp:110-135 default sizes


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def config_for(curve, group, log_n, world):
    """The workload description shared VERBATIM by both arms (ours and --impl reference)."""
    n = 1 << log_n
    deg = 1 if group == 1 else (2 if curve == 0 else 3)
    return {"workload": "%s G%d MSM, 2^%d points per GPU" % (CURVE_NAMES[curve], group, log_n),
            "curve": CURVE_NAMES[curve], "group": "G%d" % group, "points_per_gpu": n, "total_points": n * world,
            "scalars": "uniform in [0, r), 96-byte Montgomery limbs", "bases": "P0 + i*Q, affine, %d B each" % (192 * deg),
            "sharding": "point-range, one partial point per GPU, no collective on the data path",
            "cache": ("inputs larger than L2: every step streams >= %d MB of scalars and >= %d MB of base points"
                      if n * (96 + 192 * deg) > 126e6 else "inputs of %d + %d MB fit L2 at this size (not a bench configuration)")
                     % (n * 96 // 1000000, n * 192 * deg // 1000000)}


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
    """libff itself (oracle/_ref/libref.so) when built, else the plain-C port; OpenMP on every host core --
    torchrun exports OMP_NUM_THREADS=1 to its workers, which must not shrink the CPU baseline."""
    from oracle import pyoracle as po
    ref = po.load_reference()
    lib = ref if ref is not None else po.load_oracle()
    lib.set_num_threads(host_cores())
    return lib, po


def cpu_sample_log_n(lib, curve, group, log_n, bases_fn, scalars_fn):
    """Largest 2^k <= 2^log_n whose CPU MSM is expected to stay under CPU_STEP_BUDGET_S, from a 2^14-point probe
    (Pippenger's cost per point FALLS with n, so the linear extrapolation over-estimates: it errs towards a smaller
    sample, never towards a longer run)."""
    k = min(log_n, 14)
    _, t = lib.msm(curve, group, bases_fn(1 << k), scalars_fn(1 << k), method=1, chunks=0, prefilter=1)
    per_point = t / (1 << k)
    best = k
    while best < log_n and per_point * (1 << (best + 1)) * 0.75 <= CPU_STEP_BUDGET_S:
        best += 1
    return best


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (libff BDLO12 + OpenMP, all host cores) on
    the same workload.  Under torchrun only rank 0 works."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from gpu_groth16_prover_3x_b200 import synthetic
    lib, po = load_cpu_lib()
    orc = po.load_oracle()
    orc.set_num_threads(host_cores())
    curve, group = args.curve, args.group
    t0 = time.perf_counter()
    nmax = 1 << args.log_n
    all_bases = [None]

    def bases_fn(n):
        if all_bases[0] is None or all_bases[0][0] < n:
            all_bases[0] = (n, orc.gen_bases(curve, group, n))
        deg = po.degree(curve, group)
        return all_bases[0][1][:n * 24 * deg]

    def scalars_fn(n, seed=1000):
        return synthetic.random_scalars(curve, n, seed)

    k = cpu_sample_log_n(lib, curve, group, args.log_n, bases_fn, scalars_fn)
    n = 1 << k
    bases = bases_fn(n).copy()
    sets = [scalars_fn(n, 1000 + i) for i in range(2)]
    t_setup = time.perf_counter() - t0
    # one untimed step sizes the run: the timed steps must fit REF_ARM_BUDGET_S
    _, t1 = lib.msm(curve, group, bases, sets[1], method=1, chunks=0, prefilter=1)
    steps = max(1, min(args.steps, int(REF_ARM_BUDGET_S / max(t1, 1e-3))))
    warm = 1
    for i in range(max(0, min(args.warmup, 1 if t1 > 2.0 else args.warmup) - 1)):
        lib.msm(curve, group, bases, sets[i & 1], method=1, chunks=0, prefilter=1)
        warm += 1
    t0 = time.perf_counter()
    for i in range(steps):
        lib.msm(curve, group, bases, sets[i & 1], method=1, chunks=0, prefilter=1)
    dt = time.perf_counter() - t0
    value = n * steps / dt
    full = k == args.log_n
    sample = ("one FULL 2^%d-point MSM per step (the workload itself)" % k if full else
              "one 2^%d-point MSM per step: a bounded sample of the 2^%d workload (a full CPU MSM would exceed %.0f s)"
              % (k, args.log_n, CPU_STEP_BUDGET_S)) + ", libff multi_exp_with_mixed_addition<BDLO12> + OpenMP, %d threads" % lib.num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": config_for(curve, group, args.log_n, world),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": lib.num_threads(), "kind": lib.kind, "sample": sample,
                         "sample_points": n, "same_size_as_gpu_arm": full},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "steps_requested": args.steps, "warmup_requested": args.warmup, "setup_s": t_setup,
        "note": "steps are bounded so that the timed CPU work stays within %.0f s" % REF_ARM_BUDGET_S,
    }
    emit(line)


# ---- our arm -------------------------------------------------------------------------------------------
def measured_hbm_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def limbs_to_ints(a):
    """uint64[n*12] -> list of n python ints."""
    b = np.ascontiguousarray(a, dtype=np.uint64).tobytes()
    return [int.from_bytes(b[i:i + 96], "little") for i in range(0, len(b), 96)]


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
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        cpu_group = dist.new_group(backend="gloo")      # host-side waits that must not spin a kernel on the GPUs

    curve, group, n = args.curve, args.group, 1 << args.log_n
    deg = pkg.degree(curve, group)
    k_tower = {1: 1, 2: 3, 3: 6}[deg]
    r = synthetic.fr_modulus(curve)
    R = synthetic.R
    ctx = pkg.MsmContext(curve, local)
    if args.window_bits:
        ctx.set_window_bits(args.window_bits)

    # this rank's shard of the (world * n)-point MSM: bases P0 + (rank*n + i) * Q
    k0, k1 = synthetic.base_seed_scalars(curve)
    k0p, k1p = (int.from_bytes(k.tobytes(), "little") * pow(R, -1, r) % r for k in (k0, k1))   # plain integers

    def shifted_p0(first_index):
        return synthetic.int_to_limbs((k0p + first_index * k1p) % r * R % r)

    t0 = time.perf_counter()
    slot = ctx.synthetic_bases(group, n, shifted_p0(rank * n), k1)
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

    def timed(step_fn, steps, warmup, sampler=None):
        """-> (seconds by CUDA events on the launching stream, wall seconds, per-step phase timings, last result, clocks)."""
        for i in range(warmup):
            step_fn(i)
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        phases = []
        out = None
        w0 = time.perf_counter()
        with torch.cuda.stream(stream):
            e0.record(stream)
            for i in range(steps):
                out = step_fn(i)
                phases.append(ctx.last_timings())
            e1.record(stream)
        torch.cuda.synchronize()
        wall = time.perf_counter() - w0
        clocks = sampler.stop() if sampler else None
        barrier()
        return e0.elapsed_time(e1) / 1e3, wall, phases, out, clocks

    def max_over_ranks(x):
        return sharding.max_over_ranks(x, device="cuda")

    sampler = ClockSampler(local)
    dev_s, dev_wall, phases, out_dev, clocks = timed(lambda i: ctx.msm(slot, dev_sets[i & 1], n), args.steps, args.warmup, sampler)
    e2e_s, e2e_wall, _, out_host, _ = timed(lambda i: ctx.msm(slot, host_sets[i & 1], n), args.steps, args.warmup)
    dev_s, e2e_s = max_over_ranks(dev_s), max_over_ranks(e2e_s)
    dev_wall, e2e_wall = max_over_ranks(dev_wall), max_over_ranks(e2e_wall)

    # ---- strong scaling: ONE fixed n-point MSM split `world` ways, gather + fold inside the timed region ----------
    strong = None
    if not args.quick:
        off, ln = pkg.shard_ranges(n, world)[rank]
        if world == 1:
            s_slot, s_dev = slot, dev_sets[0]
        else:
            s_slot = ctx.synthetic_bases(group, ln, shifted_p0(off), k1)
            s_dev = torch.from_numpy(synthetic.random_scalars(curve, ln, 700 + rank).view(np.int64)).cuda()
        result = {}

        def strong_step(i):
            part = ctx.msm(s_slot, s_dev, ln)
            if world > 1:
                allp = sharding.gather_partials(part, device="cuda")       # 288..864 B per rank over NCCL
                if rank == 0:
                    result["xyz"] = ctx.fold(group, allp)
            else:
                result["xyz"] = part
            return part

        s_steps = max(3, min(args.steps, 10))
        s_s, _, s_ph, _, _ = timed(strong_step, s_steps, 3)
        s_s = max_over_ranks(s_s)
        s_info = s_ph[-1]
        if world > 1:
            ctx.free_bases(s_slot)
        if rank == 0:
            # closed form: sum_i s_i (P0 + i Q) = ((sum s_i) k0 + (sum i s_i) k1) * G, evaluated on the device as the
            # single point of a one-element synthetic base set
            S0 = S1 = 0
            for g in range(world):
                o, l = pkg.shard_ranges(n, world)[g]
                sc = s_dev.cpu().numpy().view(np.uint64) if world == 1 else synthetic.random_scalars(curve, l, 700 + g)
                xs = limbs_to_ints(sc)
                S0 += sum(xs)
                S1 += sum((o + i) * x for i, x in enumerate(xs))
            rinv = pow(R, -1, r)
            K = (S0 * rinv % r * k0p + S1 * rinv % r * k1p) % r
            cs = ctx.synthetic_bases(group, 1, synthetic.int_to_limbs(K * R % r), k1)
            want = ctx.download_bases(cs, 0, 1)
            ctx.free_bases(cs)
            got = ctx.to_affine(group, result["xyz"])
            t1 = dev_s / args.steps            # this rank's own n-point MSM on one GPU (the weak-scaling step)
            tn = s_s / s_steps
            strong = {"workload": "one fixed %s G%d MSM of 2^%d points split %d way(s) by point range; partial points gathered "
                                  "and folded inside the timed region" % (CURVE_NAMES[curve], group, args.log_n, world),
                      "ms": tn * 1e3, "points_per_s": n / tn, "one_gpu_ms": t1 * 1e3, "efficiency_vs_1": t1 / (world * tn),
                      "speedup_vs_1": t1 / tn, "steps": s_steps, "warmup": 3, "shard_points": ln,
                      "shard_window_bits": s_info["window_bits"], "shard_phases_ms": {k: s_info[k] for k in pkg.MsmContext.PHASES},
                      "closed_form_parity": bool((got == want).all())}

    if rank != 0:
        ctx.close()
        if world > 1:
            dist.barrier(group=cpu_group)      # rank 0 runs the proof leg on all GPUs meanwhile (host-side wait, no GPU spin)
            dist.destroy_process_group()
        return

    total_points = n * world
    value = total_points * args.steps / dev_s
    e2e = total_points * args.steps / e2e_s
    info = phases[-1]
    acc_ms = statistics.mean(p["accumulate"] for p in phases)
    W = info["windows"]
    # executed work of the accumulation phase: one batched-affine addition per sorted entry (minus one per non-empty
    # bucket, neglected) at 6 field multiplications each (SURVEY.md 8(d) counts 11 for the Jacobian mixed addition);
    # a field multiplication in the degree-k tower is k_tower Fq products of 1176 MAC
    muls_per_add = 6
    macs_per_launch = float(n) * W * muls_per_add * k_tower * 1176
    achieved = macs_per_launch / (acc_ms * 1e-3) / 1e9
    micro = {"imad_wide_gmacs": ctx.microbench(0, 4096), "imad_lo_gops": ctx.microbench(1, 4096),
             "fq_modmul_gmuls": ctx.microbench(2, 2048)}
    peak = max(micro["imad_wide_gmacs"], micro["fq_modmul_gmuls"] * 1176)
    hbm_peak, hbm_src = measured_hbm_peak()
    aff_bytes = 2 * deg * 96
    # batched-affine addition: forward reads x1, x2 and parks the prefix; backward reads both points and the
    # prefix and writes the sum; the pair descriptor (16 B) is read in both passes
    alg_bytes = float(n) * W * (2 * 96 * deg + 96 * deg + 2 * aff_bytes + 96 * deg + aff_bytes + 32 + 2)
    traffic = None
    if (curve, group, args.log_n) == (0, 1, 20):
        for name in ("r02_k_batch_add_traffic.json", "r01_k_batch_add_round0_traffic.json"):
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", name)))
                traffic = {"bytes": tj["traffic_bytes"], "launch": tj["kernel"], "source": tj["source"],
                           "algorithmic_bytes_of_that_launch": tj.get("algorithmic_bytes", float(n) * W / 2 * (alg_bytes / (float(n) * W)))}
                break
            except Exception:
                continue
    # SURVEY.md 8(d) canonical count (c = 16, W = 48, Jacobian mixed add) applied to the whole step: what the
    # reference-style algorithm would have to execute to deliver the same points/s
    canonical_macs_per_point = 1176.0 * 11 * k_tower * 48
    roofline = {
        "kernel": "k_batch_add (all rounds of one MSM)", "bound": "imad",
        "field_muls_per_addition": muls_per_add, "achieved": achieved, "peak": peak, "unit": "GMAC/s", "frac": achieved / peak,
        # DRAM bytes of one launch of the dominant kernel (dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full`
        # capture, committed under profiles/); the capture it comes from is described in traffic_detail
        "traffic": traffic["bytes"] if traffic else None, "traffic_detail": traffic,
        "launch_ms": acc_ms, "share_of_step": acc_ms / (dev_s / args.steps * 1e3),
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
    if world == 1 and not args.no_cpu and not args.quick:
        lib, po = load_cpu_lib()
        k = cpu_sample_log_n(lib, curve, group, args.log_n, lambda m_: ctx.download_bases(slot, 0, m_),
                             lambda m_: host_sets[0].numpy().view(np.uint64)[:m_ * 12].copy())
        ns = 1 << k
        sb = ctx.download_bases(slot, 0, ns)
        ss = host_sets[0].numpy().view(np.uint64)[:ns * 12].copy()
        want, t_cpu = lib.msm(curve, group, sb, ss, method=1, chunks=0, prefilter=1)
        got = ctx.to_affine(group, ctx.msm(slot, ss, ns))
        parity = bool((got == want).all())
        cpu_baseline = {"value": ns / t_cpu, "unit": UNIT, "cores": lib.num_threads(), "kind": lib.kind, "seconds": t_cpu,
                        "sample_points": ns, "same_size_as_gpu_arm": ns == n,
                        "sample": ("one FULL 2^%d-point MSM, the same bases and scalars as the GPU arm" % k if ns == n else
                                   "one 2^%d-point MSM on the first points of the same workload (a full one would exceed %.0f s)"
                                   % (k, CPU_STEP_BUDGET_S)) + " (libff BDLO12 + OpenMP, %d threads); GPU result on it bit-equal: %s"
                                  % (lib.num_threads(), parity)}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": config_for(curve, group, args.log_n, world),
        "engine": {"window_bits": info["window_bits"], "windows": W, "bucket_sets": info["bucket_sets"], "window_tables": info["tables"],
                   "table_bytes_per_gpu": binfo["bytes"], "table_build_s": binfo["table_build_ms"] / 1e3, "bases_generation_s": t_bases,
                   "sorted_list_bytes_per_step": n * W * 4,
                   "timing": "CUDA events on the launching stream around the K steps, max over ranks"},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": n * 96, "d2h_bytes_per_step": 36 * deg * 8,
                "ms_per_step": e2e_s / args.steps * 1e3},
        "gpu_launches": int(sum(p["kernel_launches"] for p in phases)),
        "wall_ms_per_step": dev_wall / args.steps * 1e3,
        "phases_ms": {k: statistics.mean(p[k] for p in phases) for k in pkg.MsmContext.PHASES},
        "roofline": roofline, "cpu_baseline": cpu_baseline, "clocks": clocks, "strong": strong,
    }
    if world == 1 and not args.no_secondary and not args.quick and (curve, group) == (0, 1):
        line["secondary"] = secondary_workload(args, local, 1, 2)                  # MNT6753 G2 (Fq3), BASELINE.json configs[2]
        try:
            line["secondary_mnt4753_g2"] = secondary_workload(args, local, 0, 2)  # the B2 query of the headline curve (Fq2)
        except Exception as e:  # the headline line must not depend on an extra leg
            line["secondary_mnt4753_g2"] = {"error": str(e)[:300]}
    ctx.close()
    if args.proof != "none" and not args.quick:
        try:
            line["proof"] = proof_leg(args.proof, world, local)
        except Exception as e:  # the headline line must not depend on the optional leg
            line["proof"] = {"error": str(e)[:300]}
    emit(line)
    if world > 1:
        dist.barrier(group=cpu_group)
        dist.destroy_process_group()


def secondary_workload(args, device, curve, group):
    """BASELINE.json's metric names two workloads; the line's `value` is MNT4753 G1 (configs[1]).  This measures
    another group -- MNT6753 G2 over the Fq3 twist (configs[2]), MNT4753 G2 over Fq2 -- the same way (device-resident
    scalars, CUDA events on the launching stream), with 3 warm-up and 3 timed MSMs, and reports it next to the headline."""
    import torch
    import gpu_groth16_prover_3x_b200 as pkg
    from gpu_groth16_prover_3x_b200 import synthetic
    n = 1 << args.log_n
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
        k_tower = {1: 1, 2: 3, 3: 6}[pkg.degree(curve, group)]
        macs = float(n) * t["windows"] * 6 * k_tower * 1176
        return {"workload": "%s G%d MSM, 2^%d points per GPU" % (CURVE_NAMES[curve], group, args.log_n), "value": n / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms,
                "steps": steps, "warmup": 3, "window_bits": t["window_bits"], "windows": t["windows"], "window_tables": t["tables"],
                "table_bytes": binfo["bytes"], "table_build_s": binfo["table_build_ms"] / 1e3, "bases_generation_s": t_bases,
                "accumulate_gmacs": macs / (t["accumulate"] * 1e-3) / 1e9,
                "phases_ms": {k: t[k] for k in pkg.MsmContext.PHASES}}
    finally:
        ctx.close()


def proof_leg(size, gpus, device):
    """BASELINE.json's second headline: end-to-end proof latency at `gpus` GPUs, sha256 against the reference.

    size = "default": instances of the reference's DEFAULT sizes (MNT4753 d = 2^20 - 1, MNT6753 d = 2^15 - 1) in its own
    file formats, written by gpu_groth16_prover_3x_b200.synthetic.write_instance (structured bases generated in HBM,
    uniform witness; the reference's `generate_parameters` takes three minutes for the same sizes).  size = "fast": the
    reference's own `generate_parameters fast` (d = 2^14 - 1 / 2^10 - 1).  Reference arm: the reference's CPU prover
    `main <curve> compute` (oracle/_ref, test infrastructure) on all host cores -- run ONCE per box: its proof and wall
    time are cached next to the instance (B200_BENCH_CACHE, default /tmp/b200_bench_cache) and reused by the runs at
    other GPU counts.  Ours: the product's command-line prover `b200_prove <curve> compute ... 3 <gpus>`
    (b200msm_key_load_sharded_file + b200msm_prove_sharded_file), third proof of a resident process."""
    import gpu_groth16_prover_3x_b200 as pkg
    from gpu_groth16_prover_3x_b200 import synthetic
    ref = os.path.join(ROOT, "oracle", "_ref")
    gen, main_bin = os.path.join(ref, "generate_parameters"), os.path.join(ref, "main")
    cli = os.path.join(ROOT, "gpu_groth16_prover_3x_b200", "b200_prove")
    cache = os.path.join(os.environ.get("B200_BENCH_CACHE", "/tmp/b200_bench_cache"), size)
    os.makedirs(cache, exist_ok=True)
    env = dict(os.environ, OMP_NUM_THREADS=str(host_cores()))
    out = {"instance": "reference default sizes, synthetic instance in the reference's file formats" if size == "default"
           else "generate_parameters fast", "gpus": gpus, "host_cores": host_cores(), "curves": {}}
    if size == "fast":
        if not os.path.exists(os.path.join(cache, "MNT6753-input")):
            if not os.path.exists(gen):
                return {"error": "oracle/_ref/generate_parameters not built"}
            subprocess.run([gen, "fast"], cwd=cache, check=True, stdout=subprocess.DEVNULL, timeout=900, env=env)
    for curve in ("MNT4753", "MNT6753"):
        prm, inp = os.path.join(cache, curve + "-parameters"), os.path.join(cache, curve + "-input")
        rec = {}
        if size == "default" and not (os.path.exists(prm) and os.path.exists(inp)):
            t0 = time.perf_counter()
            synthetic.write_instance(pkg.MNT4753 if curve == "MNT4753" else pkg.MNT6753, DEFAULT_D[curve], cache, device)
            rec["instance_write_s"] = time.perf_counter() - t0
        ref_out, ref_meta = os.path.join(cache, curve + "-output-ref"), os.path.join(cache, curve + "-ref.json")
        if os.path.exists(main_bin) and not (os.path.exists(ref_out) and os.path.exists(ref_meta)):
            t0 = time.perf_counter()
            subprocess.run([main_bin, curve, "compute", prm, inp, ref_out + ".tmp"], cwd=cache, check=True, stdout=subprocess.DEVNULL,
                           stderr=subprocess.DEVNULL, timeout=3000, env=env)
            json.dump({"wall_s": time.perf_counter() - t0, "cores": host_cores()}, open(ref_meta, "w"))
            os.replace(ref_out + ".tmp", ref_out)
            rec["reference_cached"] = False
        elif os.path.exists(ref_out):
            rec["reference_cached"] = True
        ours = os.path.join(cache, "%s-output-b200-%dgpu" % (curve, gpus))
        run = subprocess.run([cli, curve, "compute", prm, inp, ours, "3", str(gpus)], cwd=cache, check=True, capture_output=True,
                             text=True, timeout=900, env=dict(os.environ, B200MSM_TRACE="1"))
        log = run.stdout
        # B200MSM_TRACE=1: the prover reports each query's timeline on shard 0's GPU (CUDA events, ms since the proof began
        # on the device) and the SMs its lane was confined to (0 = the whole GPU); kept for the last proof
        lanes = re.findall(r"\[trace\] shard 0 lane \d (\S+)\s+sms\s+(\d+) \| start\s+([0-9.]+) sorted\s+[0-9.]+ accumulated\s+([0-9.]+) reduced\s+[0-9.]+ end\s+([0-9.]+) ms", run.stderr)
        hpoly = re.findall(r"\[trace\] shard 0 compute_H \| start\s+([0-9.]+) end\s+([0-9.]+) ms", run.stderr)
        if lanes:
            rec["device_timeline_shard0_ms"] = {q: {"sms": int(sm), "start": float(a), "accumulated": float(b), "end": float(c)} for q, sm, a, b, c in lanes[-5:]}
            if hpoly:
                rec["device_timeline_shard0_ms"]["compute_H"] = {"start": float(hpoly[-1][0]), "end": float(hpoly[-1][1])}
        times = [float(x) for x in re.findall(r"Total time from input to output: ([0-9.]+) ms", log)]
        total = re.findall(r"Total runtime \(incl. key load\): ([0-9.]+) ms", log)
        upload = re.findall(r"key load \+ window tables: ([0-9.]+) ms", log)
        dm = re.findall(r"d = (\d+), m = (\d+)", log)
        rec.update({"d": int(dm[0][0]) if dm else None, "b200_input_to_proof_s": times[-1] / 1e3 if times else None,
                    "b200_first_proof_s": times[0] / 1e3 if times else None,
                    "b200_key_load_and_tables_s": float(upload[0]) / 1e3 if upload else None,
                    "b200_process_total_3_proofs_s": float(total[0]) / 1e3 if total else None})
        if os.path.exists(ref_out):
            meta = json.load(open(ref_meta))
            rec["reference_cpu_prover_s"] = meta["wall_s"]
            rec["reference_cpu_cores"] = meta["cores"]
            rec["proof_sha256_equal"] = hashlib.sha256(open(ref_out, "rb").read()).digest() == hashlib.sha256(open(ours, "rb").read()).digest()
        else:
            rec["proof_sha256_equal"] = None
            rec["note"] = "oracle/_ref/main not built: no reference proof to compare with"
        out["curves"][curve] = rec
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
    ap.add_argument("--proof", default=os.environ.get("B200_BENCH_PROOF", "default"), choices=["default", "fast", "none"],
                    help="whole-proof leg: reference default sizes (synthetic instance), generate_parameters fast, or none")
    ap.add_argument("--quick", action="store_true", help="development: headline + roofline only (no strong / cpu / secondary / proof legs)")
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
