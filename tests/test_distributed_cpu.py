"""N > 1 path on CPU: world_size-2 gloo run of the sharding + final-gather plumbing (gpmp2_b200/distributed.py).
The per-shard compute is stood in by the oracle (tests may use it as the checker); on the GPU box the same
plumbing carries the CUDA results (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gpmp2_b200 import synth
    from gpmp2_b200.distributed import gather_ragged_to_root, gather_to_root, shard_range
    from oracle import oracle as O
    model = synth.simple_two_links_arm()
    sdf = synth.planar_dataset("OneObstacleDataset")
    st = synth.bench_setting(2, total_time=10.0, cost_sigma=0.1, epsilon=0.1, inter=4, max_iter=5)
    pr = synth.planar_problems(B, 2, seed=77)           # every rank generates the same seeded problem set
    lo, hi = shard_range(B, rank, world)
    counts = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
    sl = slice(lo, hi)
    r = O.batch_optimize(model, sdf, pr["start_conf"][sl], pr["start_vel"][sl], pr["end_conf"][sl], pr["end_vel"][sl],
                         pr["init_traj"][sl], st)
    traj = gather_ragged_to_root(torch.from_numpy(r["traj"]), counts)
    scal = gather_ragged_to_root(torch.from_numpy(np.stack([r["error"], r["coll_cost"]], axis=1)), counts)
    # equal-shard fast path
    eq = gather_to_root(torch.full((3, 2), float(rank)))
    if rank == 0:
        assert eq.shape == (3 * world, 2) and eq[3:].eq(1.0).all()
        full = O.batch_optimize(model, sdf, pr["start_conf"], pr["start_vel"], pr["end_conf"], pr["end_vel"],
                                pr["init_traj"], st)
        assert traj.shape == (B, full["traj"].shape[1])
        assert np.array_equal(traj.numpy(), full["traj"])
        assert np.array_equal(scal.numpy()[:, 0], full["error"])
        open(out_path, "w").write("ok")
    else:
        assert traj is None and scal is None
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    from gpmp2_b200.distributed import shard_range
    for B in (0, 1, 7, 64, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_shard_and_gather(tmp_path):
    out = str(tmp_path / "ok.txt")
    mp.spawn(_worker, args=(2, _free_port(), 13, out), nprocs=2, join=True)   # 13: ragged shards (6 + 7)
    assert open(out).read() == "ok"
