"""Multi-process host logic on CPU (gloo, world_size 2): the sharding plan and the final summary
gather that bench.py uses across GPUs. The per-rank work here is the oracle's CPU state generator,
which implements the same counter-based stream as the device generator."""
import os
import socket
import sys

import numpy as np
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


def _worker(rank, world, port, total, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from generalized_rbda_b200.sharding import gather_summary, shard_range
    from oracle import binding
    o = binding.OracleModel("tello")
    first, count = shard_range(total, rank, world)
    q, yd, tau = o.generate_states(count, seed=99, first_index=first, threads=1)
    ydd = o.forward_dynamics(q, yd, tau, threads=1)
    summary = gather_summary([float(first), float(count), float(ydd.sum()), float(np.abs(ydd).sum())])
    if rank == 0:
        np.save(os.path.join(out_dir, "summary.npy"), summary.numpy())
    np.save(os.path.join(out_dir, "ydd_%d.npy" % rank), ydd)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_reproduces_global_batch(tmp_path):
    total, world = 37, 2  # odd on purpose: shards of 19 and 18
    mp.spawn(_worker, args=(world, _free_port(), total, str(tmp_path)), nprocs=world, join=True)
    summary = np.load(tmp_path / "summary.npy")
    assert summary.shape == (2, 4)
    assert summary[:, 0].tolist() == [0.0, 19.0] and summary[:, 1].tolist() == [19.0, 18.0]
    sys.path.insert(0, ROOT)
    from oracle import binding
    o = binding.OracleModel("tello")
    q, yd, tau = o.generate_states(total, seed=99)
    ydd = o.forward_dynamics(q, yd, tau)
    shards = np.concatenate([np.load(tmp_path / "ydd_0.npy"), np.load(tmp_path / "ydd_1.npy")])
    assert np.array_equal(shards, ydd)
    assert abs(summary[:, 3].sum() - np.abs(ydd).sum()) <= 1e-9 * np.abs(ydd).sum()


def test_shard_ranges_partition_the_batch():
    sys.path.insert(0, ROOT)
    from generalized_rbda_b200.sharding import shard_range, weak_shard
    for total in (0, 1, 7, 1 << 20):
        for world in (1, 2, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == total
            assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))
    assert weak_shard(1 << 20, 3) == (3 << 20, 1 << 20)
