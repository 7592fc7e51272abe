"""Multi-GPU plumbing of the batched engine: the ensemble shards by member (contiguous blocks, one process per GPU,
no collective on the stepping path); the only exchange is the all-reduce of ensemble statistics.
torch.distributed is used for that plumbing only (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_range(n_total, rank, world):
    """Contiguous block partition: global members [lo, hi) owned by `rank` (member i -> rank floor(i*world/n_total))."""
    lo = (n_total * rank + world - 1) // world if rank else 0
    hi = (n_total * (rank + 1) + world - 1) // world if rank + 1 < world else n_total
    return lo, hi


def allreduce_moments(m, device=None, group=None):
    """All-reduce fb_stats moments {count, sum, sum_sq, min, max} over the ranks of `group` (one small FP64 vector:
    SUM for the additive part, MIN / MAX for the extrema). Returns the global moments plus mean / std."""
    import torch
    import torch.distributed as dist
    dev = device if device is not None else ("cuda" if dist.get_backend(group) == "nccl" else "cpu")
    add = torch.tensor([float(m["count"]), m["sum"], m["sum_sq"]], dtype=torch.float64, device=dev)
    mn = torch.tensor([m["min"]], dtype=torch.float64, device=dev)
    mx = torch.tensor([m["max"]], dtype=torch.float64, device=dev)
    dist.all_reduce(add, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mn, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    cnt, s, s2 = (float(v) for v in add.tolist())
    mean = s / cnt if cnt else float("nan")
    var = max(0.0, s2 / cnt - mean * mean) if cnt else float("nan")
    return dict(count=int(cnt), sum=s, sum_sq=s2, min=float(mn.item()), max=float(mx.item()), mean=mean, std=float(np.sqrt(var)))


def host_moments(values, status=None):
    """Reference implementation of the statistics fb_stats returns (host side, for checks)."""
    v = np.asarray(values, dtype=np.float64)
    if status is not None:
        v = v[(np.asarray(status) == 0) | (np.asarray(status) == 1)]
    return dict(count=int(v.size), sum=float(v.sum()), sum_sq=float((v * v).sum()),
                min=float(v.min()) if v.size else float("inf"), max=float(v.max()) if v.size else float("-inf"))
