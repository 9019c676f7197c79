"""world_size-2 gloo test of the multi-GPU host logic (problem sharding + the single metrics all-gather).
The CUDA engine is replaced here by the CPU oracle's collision sweep so the test runs without a GPU; what is under test
is that (i) problem generation depends only on the global problem index, (ii) shards tile the index range, (iii) the
gathered table equals the single-process table."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOTAL, T = 10, 6


def _table(problem0, n):
    sys.path.insert(0, ROOT)
    from mpinets_b200 import scenes, franka
    from oracle import oracle as O
    tables = franka.default_tables()
    p = scenes.config_problems(4, n, problem0=problem0)
    w = np.linspace(0, 1, T, dtype=np.float32)[None, :, None]
    traj = (p["q0"][:, None] * (1 - w) + p["q_goal"][:, None] * w).astype(np.float32)
    flags, first, _ = O.sweep_flags(p, traj, tables)
    m = np.zeros((n, 8), np.float32)
    m[:, 0], m[:, 1], m[:, 2] = flags, first, T - 1
    m[:, 7] = np.arange(problem0, problem0 + n)
    return m


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from mpinets_b200.parallel import shard_range, gather_metrics
    lo, hi = shard_range(rank, world, TOTAL)
    local = torch.from_numpy(_table(lo, hi - lo))
    full = gather_metrics(local, TOTAL)
    if rank == 0:
        np.save(out, full.numpy())
    dist.destroy_process_group()


def test_shard_ranges_tile_the_index_space():
    sys.path.insert(0, ROOT)
    from mpinets_b200.parallel import shard_range
    for total in (1, 7, 8, 4096, 32768, 32771):
        for world in (1, 2, 3, 8):
            r = [shard_range(k, world, total) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1


def test_two_rank_gather_matches_single_process(tmp_path):
    out = str(tmp_path / "gathered.npy")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    exp = _table(0, TOTAL)
    assert np.array_equal(got, exp)
    assert np.array_equal(got[:, 7], np.arange(TOTAL))


class _FakeEngine:
    """what parallel.broadcast_params_ needs of an Engine: the flat parameter vector in, out, and a re-pack hook"""
    def __init__(self, flat):
        self.flat, self.synced = flat, 0

    def get_params(self):
        return self.flat.clone()

    def set_params(self, flat):
        self.flat = flat.clone()

    def weights_sync(self):
        self.synced += 1


def _ddp_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from mpinets_b200.parallel import allreduce_mean_, broadcast_params_, gather_metrics, shard_range
    # DDP wrap: every rank starts from rank 0's parameters (run_training.py:71-77)
    eng = _FakeEngine(torch.full((1000,), float(rank + 1)))
    broadcast_params_(eng)
    # gradient averaging of the flat vector
    g = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    allreduce_mean_(g)
    # uneven shards (7 problems over 2 ranks: 4 + 3) take the padded all_gather path
    lo, hi = shard_range(rank, world, 7)
    local = torch.arange(lo, hi, dtype=torch.float32)[:, None].repeat(1, 8)
    full = gather_metrics(local, 7)
    torch.save({"params": eng.flat, "synced": eng.synced, "grad": g, "gathered": full}, out + f".{rank}")
    dist.destroy_process_group()


def test_two_rank_ddp_helpers(tmp_path):
    """parameter broadcast at wrap time, mean all-reduce of the flat gradient vector, and the uneven-shard gather, on 2 gloo ranks"""
    out = str(tmp_path / "ddp")
    port = 31500 + os.getpid() % 2000
    mp.spawn(_ddp_worker, args=(2, port, out), nprocs=2, join=True)
    r = [torch.load(out + f".{k}") for k in range(2)]
    for k in range(2):
        assert torch.equal(r[k]["params"], torch.full((1000,), 1.0)) and r[k]["synced"] == 1
        assert torch.equal(r[k]["grad"], torch.arange(1000, dtype=torch.float32) * 1.5)
        assert torch.equal(r[k]["gathered"][:, 0], torch.arange(7, dtype=torch.float32)) and r[k]["gathered"].shape == (7, 8)
