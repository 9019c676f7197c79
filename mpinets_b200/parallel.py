"""Multi-GPU host logic: the rollout shards by problem index (no data-path collective); the only exchange is one
all-gather of the per-problem metrics table at the end.  Backend-agnostic (nccl on GPUs, gloo in
the CPU tests)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(rank: int, world: int, total: int) -> Tuple[int, int]:
    """Contiguous block of problem indices owned by `rank` (blocks differ by at most one problem)."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_metrics(metrics: torch.Tensor, total: int) -> torch.Tensor:
    """metrics [B_local, K] on every rank -> [total, K] on every rank, rows in global problem order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return metrics
    world = dist.get_world_size()
    sizes = [shard_range(r, world, total)[1] - shard_range(r, world, total)[0] for r in range(world)]
    if len(set(sizes)) == 1:
        out = torch.empty(total, metrics.shape[1], dtype=metrics.dtype, device=metrics.device)
        dist.all_gather_into_tensor(out, metrics.contiguous())
        return out
    pad = max(sizes)
    buf = torch.zeros(pad, metrics.shape[1], dtype=metrics.dtype, device=metrics.device)
    buf[: metrics.shape[0]] = metrics
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)


def allreduce_mean_(flat: torch.Tensor) -> torch.Tensor:
    """DDP gradient averaging (run_training.py:71-77: DDPStrategy) on the flat gradient vector of the training step:
    one all-reduce of 19.07 M floats, in place.  No-op for a single process."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return flat
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(dist.get_world_size())
    return flat


def broadcast_params_(engine, src: int = 0) -> None:
    """What DistributedDataParallel does when it wraps a module (run_training.py:71-77): every rank starts from rank `src`'s
    parameters.  Broadcasts the engine's flat parameter vector and re-packs the tensor-core weight copies.  No-op for a single
    process."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    flat = engine.get_params()
    dist.broadcast(flat, src=src)
    engine.set_params(flat)
    engine.weights_sync()
