"""``mpinets.utils`` surface (``/root/reference/mpinets/utils.py:96-127,212-244``) over the C ABI."""
from __future__ import annotations

import torch

from .runtime import get_engine


def _shape_check(t: torch.Tensor):
    assert (t.ndim == 1 and t.size(0) == 7) or (t.ndim == 2 and t.size(1) == 7) or (t.ndim == 3 and t.size(2) == 7)  # utils.py:86-90


def normalize_franka_joints(batch_trajectory: torch.Tensor, limits=(-1, 1), use_real_constraints: bool = True) -> torch.Tensor:
    if not isinstance(batch_trajectory, torch.Tensor):
        raise NotImplementedError("Only torch.Tensor (CUDA) is implemented")   # utils.py:126-127
    assert tuple(limits) == (-1, 1), "the engine normalises to [-1, 1]"
    _shape_check(batch_trajectory)
    return get_engine(batch_trajectory.device).normalize(batch_trajectory.contiguous())


def unnormalize_franka_joints(batch_trajectory: torch.Tensor, limits=(-1, 1), use_real_constraints: bool = True) -> torch.Tensor:
    if not isinstance(batch_trajectory, torch.Tensor):
        raise NotImplementedError("Only torch.Tensor (CUDA) is implemented")
    assert tuple(limits) == (-1, 1)
    _shape_check(batch_trajectory)
    assert torch.all(batch_trajectory >= limits[0]) and torch.all(batch_trajectory <= limits[1])   # utils.py:200-201
    return get_engine(batch_trajectory.device).unnormalize(batch_trajectory.contiguous())
