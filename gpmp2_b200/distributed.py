"""Multi-GPU plumbing: independent problems shard across ranks (one process per GPU, torch.distributed);
there is NO data-path collective -- only a final gather of results and costs (BASELINE.json north_star).
Backend "nccl" on the GPU box, "gloo" in the CPU tests.

(C / C++ callers get the same sharding without Python through gpmp2b_batch_optimize_multi, include/gpmp2b.h:
one host process, one ctx per device, the kernels storing their slice straight into device 0's buffers over NVLink.)"""
import torch
import torch.distributed as dist


def shard_range(B, rank, world):
    """Static contiguous split: rank r takes [r*B/G, (r+1)*B/G)."""
    lo = (B * rank) // world
    hi = (B * (rank + 1)) // world
    return lo, hi


def gather_to_root(t, root=0, group=None, out=None):
    """Gather equally-shaped per-rank result tensors on `root` -> (world*n, ...) tensor there, None elsewhere.
    One collective for the trajectories, one for the packed scalars: the only communication of a solve.
    `out`: preallocated (world, *t.shape) receive buffer on the root (a fresh one is allocated otherwise)."""
    world = dist.get_world_size(group)
    if world == 1:
        return t
    rank = dist.get_rank(group)
    if rank == root:
        if out is None:
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


class ShardedPlanner:
    """One rank's share of a batched solve plus the final gather, with every buffer allocated ONCE.

    Every rank owns `B_local` problems (its shard).  run_device(): inputs already on this rank's GPU -> optimize ->
    gather (trajectories, then [error, collision cost] and [iterations, status]) into preallocated buffers on the
    root.  run_host(): the same from pinned HOST inputs (host->device copies inside), and on the root the gathered
    results come back to pinned host memory -- the end-to-end path of a multi-GPU user."""

    def __init__(self, ctx, model, sdf, setting, B_local, device, group=None, root=0):
        from . import api
        self.api, self.ctx, self.model, self.sdf, self.st = api, ctx, model, sdf, setting
        self.B, self.dev, self.group, self.root = int(B_local), device, group, root
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        D, N = setting.dof, setting.total_step + 1
        self.D, self.TL = D, 2 * N * D
        f64, i32 = torch.float64, torch.int32
        self.traj = torch.empty((self.B, self.TL), dtype=f64, device=device)
        self.scal = torch.empty((2, self.B), dtype=f64, device=device)          # error | collision cost
        self.ints = torch.empty((2, self.B), dtype=i32, device=device)          # iterations | status
        self.d_in = {k: torch.empty((self.B, n), dtype=f64, device=device)
                     for k, n in (("start_conf", D), ("start_vel", D), ("end_conf", D), ("end_vel", D), ("init_traj", self.TL))}
        is_root = self.rank == root
        self.root_traj = self.root_scal = self.root_ints = None
        self.h_traj = self.h_scal = self.h_ints = None
        if self.world > 1 and is_root:
            self.root_traj = torch.empty((self.world, self.B, self.TL), dtype=f64, device=device)
            self.root_scal = torch.empty((self.world, 2, self.B), dtype=f64, device=device)
            self.root_ints = torch.empty((self.world, 2, self.B), dtype=i32, device=device)

    def run_device(self, d, stream=None):
        """d: dict of device tensors (start_conf, start_vel, end_conf, end_vel, init_traj) of this rank's shard."""
        s = stream if stream is not None else torch.cuda.current_stream(self.dev)
        self.api.batch_optimize_device(
            self.model, self.sdf, self.st, self.B, d["start_conf"].data_ptr(), d["start_vel"].data_ptr(),
            d["end_conf"].data_ptr(), d["end_vel"].data_ptr(), d["init_traj"].data_ptr(), self.traj.data_ptr(),
            self.scal[0].data_ptr(), self.scal[1].data_ptr(), self.ints[0].data_ptr(), self.ints[1].data_ptr(),
            stream=s.cuda_stream, ctx=self.ctx)
        if self.world > 1:   # the only communication: final gather of results and costs
            gather_to_root(self.traj, self.root, self.group, out=self.root_traj)
            gather_to_root(self.scal, self.root, self.group, out=self.root_scal)
            gather_to_root(self.ints, self.root, self.group, out=self.root_ints)

    def run_host(self, h):
        """h: dict of PINNED host tensors of this rank's shard.  Returns on the root (after a synchronize) the pinned
        host tensors (traj [world][B][TL], scal [world][2][B], ints [world][2][B]); None elsewhere."""
        for k, t in self.d_in.items():
            t.copy_(h[k], non_blocking=True)
        self.run_device(self.d_in)
        if self.rank != self.root:
            return None
        if self.h_traj is None:   # pinned result buffers of the root, allocated on first use
            w = self.world
            self.h_traj = torch.empty((w, self.B, self.TL), dtype=torch.float64).pin_memory()
            self.h_scal = torch.empty((w, 2, self.B), dtype=torch.float64).pin_memory()
            self.h_ints = torch.empty((w, 2, self.B), dtype=torch.int32).pin_memory()
        if self.world > 1:
            self.h_traj.copy_(self.root_traj, non_blocking=True)
            self.h_scal.copy_(self.root_scal, non_blocking=True)
            self.h_ints.copy_(self.root_ints, non_blocking=True)
        else:
            self.h_traj[0].copy_(self.traj, non_blocking=True)
            self.h_scal[0].copy_(self.scal, non_blocking=True)
            self.h_ints[0].copy_(self.ints, non_blocking=True)
        return self.h_traj, self.h_scal, self.h_ints

    def h2d_bytes(self):
        return self.B * (4 * self.D + self.TL) * 8

    def d2h_bytes_root(self):
        return self.world * self.B * (self.TL * 8 + 2 * 8 + 2 * 4)
