"""Batch sharding over GPUs. States are independent, so the global state index range
is cut into contiguous shards, one per rank; every rank generates (counter-based RNG keyed by the
global index) and evaluates its own shard; nothing is exchanged on the data path. The only
collective is the final gather of per-rank timings and checksums (torch.distributed: NCCL on the
GPUs, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous [first, first + count) of `total` states for `rank`; sizes differ by at most one."""
    base, extra = divmod(total, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def weak_shard(per_rank, rank):
    """Weak scaling: every rank owns `per_rank` states of the global index range."""
    return rank * per_rank, per_rank


def gather_summary(values, device=None):
    """All-gather a small list of floats from every rank -> tensor [world, len(values)]."""
    t = torch.tensor(values, dtype=torch.float64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return t[None, :]
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.stack(out)
