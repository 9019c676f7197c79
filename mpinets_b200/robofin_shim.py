"""``robofin.pointcloud.torch`` names used by the reference (``FrankaSampler``, ``FrankaCollisionSampler``; call sites
``/root/reference/mpinets/model.py:25,250,267-275,300``, ``run_inference.py:111-116,169,188``) over the C ABI."""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np
import torch

from .franka import DEFAULT_PRISMATIC_VALUE
from .runtime import get_engine


class FrankaSampler:
    def __init__(self, device, num_fixed_points: Optional[int] = None, use_cache: bool = False,
                 default_prismatic_value: float = DEFAULT_PRISMATIC_VALUE, with_base_link: bool = True):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("mpinets_b200.FrankaSampler is CUDA only")
        self.engine = get_engine(self.device)
        self.num_fixed_points = num_fixed_points
        self._step = 0

    def sample(self, q: torch.Tensor, num_points: Optional[int] = None) -> torch.Tensor:
        """q [B,7] -> [B,P,3]; a fresh keyed subset per call (robofin draws np.random.choice per call),
        a fixed one when num_fixed_points is set (loss.py:141-147)."""
        n = num_points or self.num_fixed_points or self.engine.tables.link_points.shape[0]
        step = 0 if self.num_fixed_points else self._step
        if not self.num_fixed_points:
            self._step += 1
        cloud = self.engine.sample_robot(q.contiguous().float(), n, step)
        return cloud[..., :3]

    def end_effector_pose(self, q: torch.Tensor, frame: str = "right_gripper") -> torch.Tensor:
        assert frame == "right_gripper"
        _, eef = self.engine.fk(q.contiguous().float())
        B = q.shape[0]
        out = torch.zeros(B, 4, 4, device=q.device)
        out[:, :3] = eef
        out[:, 3, 3] = 1
        return out

    def sample_end_effector(self, poses: torch.Tensor, num_points: int, frame: str = "right_gripper") -> torch.Tensor:
        assert frame == "right_gripper"
        ee = torch.from_numpy(self.engine.tables.ee_points).to(poses.device)
        perm = torch.randperm(ee.shape[0], device=poses.device)[:num_points]
        p = ee[perm]
        return torch.einsum("bij,pj->bpi", poses[:, :3, :3].float(), p) + poses[:, None, :3, 3].float()


class FrankaCollisionSampler:
    def __init__(self, device, default_prismatic_value: float = DEFAULT_PRISMATIC_VALUE, with_base_link: bool = True,
                 margin: float = 0.0):
        self.device = torch.device(device)
        self.engine = get_engine(self.device)
        self.margin = margin
        radii = self.engine.tables.sphere_radii
        self._groups = [(float(r), np.nonzero(radii == r)[0]) for r in sorted(set(radii.tolist()))]

    def compute_spheres(self, q: torch.Tensor) -> List[Tuple[float, torch.Tensor]]:
        """-> list of (radius, centres [B, n_r, 3]) grouped by radius (model.py:300-303)"""
        c = self.engine.compute_spheres(q.contiguous().float())
        return [(r + self.margin, c[:, torch.as_tensor(ix, device=c.device)]) for r, ix in self._groups]
