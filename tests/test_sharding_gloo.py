"""World-size-2 gloo run (CPU) of the multi-GPU host logic: point-range sharding, gather of the per-rank
partial Jacobian points, fold.  The per-shard MSM is done by the oracle here (no GPU in this test); on
the GPU box the same plumbing carries the engine's partials (bench.py --gpus N, test_gpu_msm.py)."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, curve, group, out_path):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from gpu_groth16_prover_3x_b200 import sharding
    from oracle import pyoracle as po
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    orc = po.load_oracle()
    deg = po.degree(curve, group)
    bases = orc.gen_bases(curve, group, n)
    sc = po.gen_scalars(curve, n, 11)
    off, ln = sharding.my_shard(n, rank, world)
    aff, _ = orc.msm(curve, group, bases[off * 24 * deg:(off + ln) * 24 * deg], sc[off * 12:(off + ln) * 12])
    # affine -> Jacobian partial (Z = 1, or Z = 0 for infinity) in Montgomery form
    one = po.int_to_limbs(po.R % po.fq_modulus(curve))
    z = np.zeros(12 * deg, np.uint64)
    if aff[12 * deg:].any():
        z[:12] = one
    part = np.concatenate([aff, z])
    allp = sharding.gather_partials(part)
    tmax = sharding.max_over_ranks(float(rank + 1))
    if rank == 0:
        folded = orc.fold_jacobian(curve, group, allp)
        whole, _ = orc.msm(curve, group, bases, sc)
        np.savez(out_path, folded=folded, whole=whole, tmax=tmax, nparts=len(allp) // (36 * deg))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("curve,group,n", [(0, 1, 101), (1, 2, 33)])
def test_shard_gather_fold_world2(tmp_path, curve, group, n):
    world = 2
    port = 29500 + (os.getpid() + 7 * curve + group) % 2000
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(world, port, n, curve, group, out), nprocs=world, join=True)
    z = np.load(out)
    assert int(z["nparts"]) == world
    assert float(z["tmax"]) == float(world)
    assert (z["folded"] == z["whole"]).all()
