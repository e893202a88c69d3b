"""MSM size sweep, ONE MSM of 2^k points split by point range over the ranks (BASELINE.json configs[4]).
   python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P \\
          tools/size_sweep_sharded.py curve group log_lo log_hi [step]
Every rank owns the shard b200msm_shard_range gives it (bases P0 + i*Q generated in its HBM), runs its shard, the partial
points are gathered and folded on rank 0 inside the timed region; the result is checked against the closed form
(sum s_i) P0 + (sum i s_i) Q.  One JSON line per size on rank 0 (best of 3 after a warm-up, max over ranks)."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import gpu_groth16_prover_3x_b200 as pkg
from gpu_groth16_prover_3x_b200 import sharding, synthetic

curve, group, lo, hi = (int(x) for x in sys.argv[1:5])
step = int(sys.argv[5]) if len(sys.argv) > 5 else 2
rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = pkg.MsmContext(curve, local)
r, R = synthetic.fr_modulus(curve), synthetic.R
k0, k1 = synthetic.base_seed_scalars(curve)
k0p, k1p = (int.from_bytes(k.tobytes(), "little") * pow(R, -1, r) % r for k in (k0, k1))
names = {(0, 1): "MNT4753 G1", (0, 2): "MNT4753 G2", (1, 1): "MNT6753 G1", (1, 2): "MNT6753 G2"}


def ints(a):
    b = np.ascontiguousarray(a, dtype=np.uint64).tobytes()
    return [int.from_bytes(b[i:i + 96], "little") for i in range(0, len(b), 96)]


for log_n in range(lo, hi + 1, step):
    n = 1 << log_n
    off, ln = pkg.shard_ranges(n, world)[rank]
    slot = ctx.synthetic_bases(group, ln, synthetic.int_to_limbs((k0p + off * k1p) % r * R % r), k1)
    info = ctx.bases_info(slot)
    sc_host = synthetic.random_scalars(curve, ln, 900 + rank)
    sc = torch.from_numpy(sc_host.view(np.int64)).cuda()
    best, res = None, None
    for it in range(4):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        part = ctx.msm(slot, sc, ln)
        allp = sharding.gather_partials(part, device="cuda")
        if rank == 0:
            res = ctx.fold(group, allp) if world > 1 else part
        torch.cuda.synchronize()
        dt = sharding.max_over_ranks(time.perf_counter() - t0, device="cuda")
        if it and (best is None or dt < best):
            best = dt
    t = ctx.last_timings()
    # closed form on rank 0: every rank's sum s_i and sum (off + i) s_i
    xs = ints(sc_host)
    mine = torch.tensor([float(0)], device="cuda")      # keep NCCL happy with a trivial op before the python-int gather
    s0, s1 = sum(xs), sum((off + i) * x for i, x in enumerate(xs))
    if world > 1:
        objs = [None] * world
        dist.all_gather_object(objs, (s0, s1))
        s0, s1 = sum(o[0] for o in objs), sum(o[1] for o in objs)
    if rank == 0:
        rinv = pow(R, -1, r)
        K = (s0 * rinv % r * k0p + s1 * rinv % r * k1p) % r
        cs = ctx.synthetic_bases(group, 1, synthetic.int_to_limbs(K * R % r), k1)
        want = ctx.download_bases(cs, 0, 1)
        ctx.free_bases(cs)
        ok = bool((ctx.to_affine(group, res) == want).all())
        print(json.dumps({"workload": "%s MSM, 2^%d points over %d GPU(s)" % (names[(curve, group)], log_n, world), "ms": best * 1e3,
                          "points_per_s": n / best, "shard_points": ln, "window_bits": t["window_bits"], "tables": t["tables"],
                          "bucket_sets": t["bucket_sets"], "shard_ms": t["total"], "table_bytes_per_gpu": info["bytes"],
                          "table_build_s": info["table_build_ms"] / 1e3, "closed_form_parity": ok}), flush=True)
    ctx.free_bases(slot)
    del sc
ctx.close()
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
