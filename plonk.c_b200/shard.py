"""Multi-GPU sharding of a proof batch.

Proofs are independent, so the path shards with NO data-path collective: rank r of R owns the contiguous item
range [r*N/R, (r+1)*N/R) of the synthetic stream (inputs are a pure function of (seed, item index), so a rank
needs only (seed, start, count)).  The only exchange is the end-of-batch reduction of a few counters
(per-status histogram, accept count, proof-byte checksum: sum) and of the per-rank elapsed time (max) --
tens of bytes, over NCCL on GPUs and over gloo in the CPU tests.
"""
import numpy as np

N_COUNTERS = 18   # 0..14 status histogram, 15 other, 16 accepted, 17 checksum (sum of proof bytes)


def shard_range(n_total, rank, world):
    """Contiguous, balanced, covering partition of [0, n_total)."""
    lo = n_total * rank // world
    hi = n_total * (rank + 1) // world
    return lo, hi - lo


def tally_host(proofs, status, verdict):
    """numpy restatement of the on-device tally (pb_tally_dev) -- used for the host path and the CPU tests."""
    c = np.zeros(N_COUNTERS, np.int64)
    st = np.asarray(status)
    hist = np.bincount(np.minimum(st, 15), minlength=16)
    c[:16] = hist[:16]
    if verdict is not None:
        c[16] = int((np.asarray(verdict) == 1).sum())
    if proofs is not None:
        c[17] = int(np.asarray(proofs, dtype=np.int64).sum())
    return c


def reduce_counters(counts, elapsed_ms, group=None):
    """All-reduce the per-rank counters (sum) and the per-rank elapsed time (max).  `counts` is a torch int64
    tensor on the backend's device (CUDA for NCCL, CPU for gloo); returns (global counts as numpy, max elapsed)."""
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=counts.device)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        elapsed_ms = float(t.item())
    return counts.cpu().numpy(), float(elapsed_ms)


def _world():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) else None


def gather_scalar(x, device):
    """Every rank's value of a scalar, as a list indexed by rank (one all_gather of a float64)."""
    import torch
    dist = _world()
    if dist is None:
        return [float(x)]
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [float(o.item()) for o in out]


def reduce_max(x, device):
    import torch
    dist = _world()
    if dist is None:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(x, device):
    import torch
    dist = _world()
    if dist is None:
        return int(x)
    t = torch.tensor([int(x)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())
