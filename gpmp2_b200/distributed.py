"""Multi-GPU plumbing: independent problems shard across ranks (one process per GPU, torch.distributed);
there is NO data-path collective -- only a final gather of results and costs (BASELINE.json north_star).
Backend "nccl" on the GPU box, "gloo" in the CPU tests."""
import torch
import torch.distributed as dist


def shard_range(B, rank, world):
    """Static contiguous split: rank r takes [r*B/G, (r+1)*B/G)."""
    lo = (B * rank) // world
    hi = (B * (rank + 1)) // world
    return lo, hi


def gather_to_root(t, root=0, group=None):
    """Gather equally-shaped per-rank result tensors on `root` -> (world*n, ...) tensor there, None elsewhere.
    One collective for the trajectories, one for the packed scalars: the only communication of a solve."""
    world = dist.get_world_size(group)
    if world == 1:
        return t
    rank = dist.get_rank(group)
    if rank == root:
        out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
        dist.gather(t, list(out.unbind(0)), dst=root, group=group)
        return out.reshape((-1,) + tuple(t.shape[1:]))
    dist.gather(t, None, dst=root, group=group)
    return None


def gather_ragged_to_root(t, counts, root=0, group=None):
    """Same for shards of unequal length (B not divisible by world): pad to the longest shard, gather, trim."""
    world = dist.get_world_size(group)
    if world == 1:
        return t
    nmax = max(counts)
    pad = torch.zeros((nmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    g = gather_to_root(pad, root, group)
    if g is None:
        return None
    g = g.reshape((world, nmax) + tuple(t.shape[1:]))
    return torch.cat([g[r, : counts[r]] for r in range(world)], dim=0)
