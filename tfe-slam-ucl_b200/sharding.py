"""Multi-GPU host logic: the hot path shards by independent scan streams.

Streams share nothing (own grids, pose, update counters), so stream g lives on exactly one rank and there is
NO data-path collective; the only communication is the end-of-run reduction of (units processed, elapsed time)
used for reporting. One process per GPU, torch.distributed for that plumbing (nccl on GPUs, gloo in CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n_total, world, rank):
    """Contiguous, balanced block of the global stream ids [0, n_total) owned by `rank`: (first, count)."""
    if world < 1 or not (0 <= rank < world) or n_total < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(n_total, world)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def owner_of(stream, n_total, world):
    """Rank that owns global stream id `stream` under shard_range."""
    base, rem = divmod(n_total, world)
    split = rem * (base + 1)
    if stream < split:
        return stream // (base + 1)
    return rem + (stream - split) // base if base else world - 1


def aggregate_throughput(local_units, local_seconds, device="cpu"):
    """Whole-job (units, seconds): units summed over ranks, time = MAX over ranks (never wall clock of rank 0)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(local_units), float(local_seconds)
    u = torch.tensor([float(local_units)], dtype=torch.float64, device=device)
    t = torch.tensor([float(local_seconds)], dtype=torch.float64, device=device)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(u.item()), float(t.item())


def gather_rows(local_rows, device="cpu"):
    """Host-side concatenation of per-rank result rows (e.g. final poses) in global stream order, on every rank."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local_rows
    world = dist.get_world_size()
    n = torch.tensor([local_rows.shape[0]], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    mx = int(max(int(s.item()) for s in sizes))
    pad = torch.zeros((mx,) + tuple(local_rows.shape[1:]), dtype=local_rows.dtype, device=device)
    pad[: local_rows.shape[0]] = local_rows.to(device)
    out = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[: int(s.item())] for o, s in zip(out, sizes)], dim=0)
