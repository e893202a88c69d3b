"""Multi-GPU plumbing of the MSM path: an MSM shards by point range, every rank
returns one partial Jacobian point, and the partials are summed.  There is no collective on the data
path; torch.distributed only carries the 288..864-byte partial points to rank 0 (and the barrier /
max-over-ranks timing in bench.py).  Works with the nccl backend on GPUs and with gloo on CPU
(tests/test_sharding_gloo.py)."""
import numpy as np

from .engine import shard_ranges


def my_shard(n_total, rank, world):
    """(offset, length) of this rank's contiguous point range."""
    return shard_ranges(n_total, world)[rank]


def gather_partials(partial_xyz, device=None):
    """All-gather one partial Jacobian point (uint64 limbs, numpy) per rank -> uint64[world * len] on every
    rank, in rank order.  Without an initialised process group it returns the input unchanged."""
    import torch
    import torch.distributed as dist
    part = np.ascontiguousarray(partial_xyz, dtype=np.uint64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return part.copy()
    t = torch.from_numpy(part.view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return np.concatenate([o.cpu().numpy().view(np.uint64) for o in out])


def max_over_ranks(x, device=None):
    """Maximum of a python float over all ranks (the timing rule of bench.py)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
