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
        self._ee_calls = 0
        # with_base_link=False (loss.py:141-147) drops panda_link0's points.  The engine's permuted subset runs over its whole link
        # table, so that variant is served only when the table itself carries no base-link points (RobotTables is a data input);
        # the loss container's fixed 1024-point cloud is built inside the library (mpn_bc_collision_losses), which skips them.
        if not with_base_link and int((self.engine.tables.link_ids == 0).sum()) > 0:
            raise NotImplementedError("FrankaSampler(with_base_link=False): load RobotTables without panda_link0 points "
                                      "(franka.synthetic_link_points(with_base_link=False)) or use loss.CollisionAndBCLossContainer")

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
        """poses [B,4,4] (or [B,3,4]) in the right_gripper frame -> [B,num_points,3] on the library (mpn_sample_end_effector): a keyed
        subset per call, like robofin's np.random.choice per call"""
        assert frame == "right_gripper"
        p34 = poses[:, :3, :].contiguous().float()
        out = self.engine.sample_end_effector(p34, num_points, problem0=self._ee_calls)
        self._ee_calls += p34.shape[0]
        return out


class FrankaCollisionSampler:
    def __init__(self, device, default_prismatic_value: float = DEFAULT_PRISMATIC_VALUE, with_base_link: bool = True,
                 margin: float = 0.0):
        self.device = torch.device(device)
        self.engine = get_engine(self.device)
        self.margin = margin
        radii, links = self.engine.tables.sphere_radii, self.engine.tables.sphere_links
        has_base = bool((links == 0).any())
        if with_base_link and not has_base:
            raise RuntimeError("FrankaCollisionSampler(with_base_link=True) but the engine's sphere table has no panda_link0 sphere "
                               "(franka.default_tables(with_base_link_spheres=True))")
        keep = np.ones(len(radii), bool) if with_base_link else links != 0   # with_base_link=False drops link0's sphere (model.py:269-271)
        self._groups = [(float(r), np.nonzero((radii == r) & keep)[0]) for r in sorted(set(radii[keep].tolist()))]

    def compute_spheres(self, q: torch.Tensor) -> List[Tuple[float, torch.Tensor]]:
        """-> list of (radius, centres [B, n_r, 3]) grouped by radius (model.py:300-303)"""
        c = self.engine.compute_spheres(q.contiguous().float())
        return [(r + self.margin, c[:, torch.as_tensor(ix, device=c.device)]) for r, ix in self._groups]
