"""``mpinets.geometry`` surface (``/root/reference/mpinets/geometry.py``) over the C ABI (forward only)."""
from __future__ import annotations

from typing import Dict, Sequence

import numpy as np
import torch

from .engine import Engine

_ENGINES: Dict[tuple, Engine] = {}
_MAX_ENGINES = 16   # contexts are keyed by primitive-row counts; a bounded cache keeps odd shapes from accumulating device memory


def _engine_for(device: torch.device, m1: int, m2: int) -> Engine:
    key = (device.index or 0, m1, m2)
    if key not in _ENGINES:
        if len(_ENGINES) >= _MAX_ENGINES:
            _ENGINES.pop(next(iter(_ENGINES))).close()
        _ENGINES[key] = Engine(device=key[0], max_cuboids=m1, max_cylinders=m2)
    return _ENGINES[key]


def _unit_quats(B, M, device):
    q = torch.zeros(B, M, 4, device=device)
    q[..., 0] = 1
    return q


class _Prims:
    def _scene(self):
        raise NotImplementedError

    def sdf(self, points: torch.Tensor) -> torch.Tensor:
        """points [B,N,3] -> scene SDF [B,N] (min over this family; +inf when every primitive is zero-volume)"""
        assert points.ndim == 3
        scene, eng, which = self._scene()
        return eng.sdf_points(scene, points.contiguous().float(), which)

    def sdf_sequence(self, points: torch.Tensor) -> torch.Tensor:
        """points [B,T,N,3] -> [B,T,N]"""
        assert points.ndim == 4
        B, T, N, _ = points.shape
        scene, eng, which = self._scene()
        return eng.sdf_points(scene, points.reshape(B, T * N, 3).contiguous().float(), which).reshape(B, T, N)


class TorchCuboids(_Prims):
    """geometry.py:126-347"""

    def __init__(self, centers: torch.Tensor, dims: torch.Tensor, quaternions: torch.Tensor):
        assert centers.ndim == 3 and dims.ndim == 3 and quaternions.ndim == 3
        self.centers, self.dims, self.quats = centers, dims, quaternions

    def _scene(self):
        B, M, _ = self.centers.shape
        d = self.centers.device
        z = lambda *s: torch.zeros(*s, device=d)
        scene = dict(cuboid_centers=self.centers.contiguous().float(), cuboid_dims=self.dims.contiguous().float(),
                     cuboid_quats=self.quats.contiguous().float(), cylinder_centers=z(B, 1, 3), cylinder_radii=z(B, 1, 1),
                     cylinder_heights=z(B, 1, 1), cylinder_quats=_unit_quats(B, 1, d))
        return scene, _engine_for(d, M, 1), 1

    def surface_area(self) -> torch.Tensor:
        return 2 * (self.dims[:, :, 0] * self.dims[:, :, 1] + self.dims[:, :, 0] * self.dims[:, :, 2] + self.dims[:, :, 1] * self.dims[:, :, 2])


class TorchCylinders(_Prims):
    """geometry.py:350-568"""

    def __init__(self, centers: torch.Tensor, radii: torch.Tensor, heights: torch.Tensor, quaternions: torch.Tensor):
        assert centers.ndim == 3 and radii.ndim == 3 and heights.ndim == 3 and quaternions.ndim == 3
        self.centers, self.radii, self.heights, self.quats = centers, radii, heights, quaternions

    def _scene(self):
        B, M, _ = self.centers.shape
        d = self.centers.device
        z = lambda *s: torch.zeros(*s, device=d)
        scene = dict(cuboid_centers=z(B, 1, 3), cuboid_dims=z(B, 1, 3), cuboid_quats=_unit_quats(B, 1, d),
                     cylinder_centers=self.centers.contiguous().float(), cylinder_radii=self.radii.contiguous().float(),
                     cylinder_heights=self.heights.contiguous().float(), cylinder_quats=self.quats.contiguous().float())
        return scene, _engine_for(d, 1, M), 2


def construct_mixed_point_cloud(obstacles: Sequence, num_points: int, device: int = 0, seed_problem: int = 0) -> np.ndarray:
    """geometry.py:571-608 for a list of primitives carrying ``center / dims|radius,height / quaternion`` attributes
    (geometrout-style) -> ndarray [num_points, 4] (xyz + label 1).  Empty list -> ``np.array([[]])`` like the reference."""
    if len(obstacles) == 0:
        return np.array([[]])
    cubs = [o for o in obstacles if hasattr(o, "dims")]
    cyls = [o for o in obstacles if hasattr(o, "radius")]
    m1, m2 = max(1, len(cubs)), max(1, len(cyls))
    key = ("cloud", device, num_points, m1, m2)
    if key not in _ENGINES:
        if len(_ENGINES) >= _MAX_ENGINES:   # bounded cache: drop (and close) the oldest context
            _ENGINES.pop(next(iter(_ENGINES))).close()
        _ENGINES[key] = Engine(device=device, n_robot=1, n_obstacle=num_points, n_target=0, max_cuboids=m1, max_cylinders=m2)
    eng = _ENGINES[key]
    f = lambda a: torch.tensor(np.asarray(a, dtype=np.float32), device=eng.device)
    scene = dict(cuboid_centers=torch.zeros(1, m1, 3), cuboid_dims=torch.zeros(1, m1, 3), cuboid_quats=torch.zeros(1, m1, 4),
                 cylinder_centers=torch.zeros(1, m2, 3), cylinder_radii=torch.zeros(1, m2, 1), cylinder_heights=torch.zeros(1, m2, 1),
                 cylinder_quats=torch.zeros(1, m2, 4))
    scene["cuboid_quats"][..., 0] = 1; scene["cylinder_quats"][..., 0] = 1
    for i, o in enumerate(cubs):
        scene["cuboid_centers"][0, i], scene["cuboid_dims"][0, i] = f(o.center).cpu(), f(o.dims).cpu()
        scene["cuboid_quats"][0, i] = f(getattr(o, "quaternion", [1, 0, 0, 0])).cpu()
    for i, o in enumerate(cyls):
        scene["cylinder_centers"][0, i], scene["cylinder_radii"][0, i, 0] = f(o.center).cpu(), float(o.radius)
        scene["cylinder_heights"][0, i, 0], scene["cylinder_quats"][0, i] = float(o.height), f(getattr(o, "quaternion", [1, 0, 0, 0])).cpu()
    scene = {k: v.to(eng.device).contiguous() for k, v in scene.items()}
    q0 = torch.zeros(1, 7, device=eng.device)
    target = torch.eye(4, device=eng.device)[:3].unsqueeze(0).contiguous()
    cloud = eng.build_cloud(scene, q0, target, problem0=seed_problem)
    return cloud[0, 1:1 + num_points].cpu().numpy().astype(np.float64)
