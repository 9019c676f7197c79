"""``mpinets/mpinets_types.py`` surface: the ``PlanningProblem`` record (mpinets_types.py:34-45) and its conversion to the
structure-of-arrays layout the engine consumes (the keys of data_loader.py:206-235 / model.py:213-220).

The reference's primitives and poses come from ``geometrout`` (not installable here); the minimal ``Cuboid`` / ``Cylinder``
/ ``SE3`` classes below carry the same attribute names the reference reads (``center``, ``dims``, ``radius``, ``height``,
``pose.so3.wxyz``, ``xyz``, ``so3.wxyz``, ``matrix``), and ``problems_to_soa`` only relies on those attributes, so real
geometrout objects un-pickled from a reference ``ProblemSet`` convert unchanged.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Union

import numpy as np

from .scenes import SCENE_KEYS


class SO3:
    def __init__(self, wxyz: Sequence[float]):
        q = np.asarray(wxyz, dtype=np.float64)
        self.wxyz = q / np.linalg.norm(q)

    @property
    def matrix(self) -> np.ndarray:
        w, x, y, z = self.wxyz
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


class SE3:
    """pose with the attributes the reference uses: ``xyz``, ``so3.wxyz``, ``matrix`` (run_inference.py:113-116,176-187)"""

    def __init__(self, xyz: Sequence[float], quaternion: Sequence[float] = (1, 0, 0, 0)):
        self.xyz = np.asarray(xyz, dtype=np.float64)
        self._xyz = self.xyz
        self.so3 = SO3(quaternion)

    @property
    def matrix(self) -> np.ndarray:
        m = np.eye(4)
        m[:3, :3] = self.so3.matrix
        m[:3, 3] = self.xyz
        return m

    @classmethod
    def from_matrix(cls, m: np.ndarray) -> "SE3":
        m = np.asarray(m, dtype=np.float64)
        R = m[:3, :3]
        w = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
        if w > 1e-6:
            q = [w, (R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w)]
        else:   # half-turn: take the axis from the symmetric part
            x = np.sqrt(max(0.0, (1 + R[0, 0]) / 2)); y = np.sqrt(max(0.0, (1 + R[1, 1]) / 2)); z = np.sqrt(max(0.0, (1 + R[2, 2]) / 2))
            y = np.copysign(y, R[0, 1] + R[1, 0]) if x > 0 else y
            z = np.copysign(z, R[0, 2] + R[2, 0]) if x > 0 else np.copysign(z, R[1, 2] + R[2, 1])
            q = [0.0, x, y, z]
        return cls(m[:3, 3], q)


class Cuboid:
    def __init__(self, center: Sequence[float], dims: Sequence[float], quaternion: Sequence[float] = (1, 0, 0, 0)):
        self.center = np.asarray(center, dtype=np.float64)
        self.dims = np.asarray(dims, dtype=np.float64)
        self.pose = SE3(center, quaternion)

    def is_zero_volume(self) -> bool:
        return bool(np.isclose(self.dims, 0).any())


class Cylinder:
    def __init__(self, center: Sequence[float], radius: float, height: float, quaternion: Sequence[float] = (1, 0, 0, 0)):
        self.center = np.asarray(center, dtype=np.float64)
        self.radius, self.height = float(radius), float(height)
        self.pose = SE3(center, quaternion)

    def is_zero_volume(self) -> bool:
        return bool(np.isclose(self.radius, 0) or np.isclose(self.height, 0))


Obstacles = List[Union[Cuboid, Cylinder]]


@dataclass
class PlanningProblem:
    """mpinets_types.py:34-45"""
    target: SE3                      # the target in the ``right_gripper`` frame
    target_volume: Union[Cuboid, Cylinder]
    q0: np.ndarray                   # the starting configuration
    obstacles: Optional[Obstacles] = None
    obstacle_point_cloud: Optional[np.ndarray] = None
    target_negative_volumes: Obstacles = field(default_factory=lambda: [])


ProblemSet = Dict[str, Dict[str, List[PlanningProblem]]]   # mpinets_types.py:48: env type -> problem type -> problems


def _is_cylinder(o) -> bool:
    return hasattr(o, "radius") and hasattr(o, "height")


def _quat(o) -> np.ndarray:
    return np.asarray(o.pose.so3.wxyz, dtype=np.float32)


def primitives_to_soa(groups: Sequence[Sequence], max_cuboids: int, max_cylinders: int) -> Dict[str, np.ndarray]:
    """B lists of Cuboid / Cylinder -> the padded arrays of data_loader.py:198-235: zero-volume rows are padding, padding
    quaternions are unit (data_loader.py:198-206)."""
    B = len(groups)
    out = dict(
        cuboid_centers=np.zeros((B, max_cuboids, 3), np.float32), cuboid_dims=np.zeros((B, max_cuboids, 3), np.float32),
        cuboid_quats=np.zeros((B, max_cuboids, 4), np.float32), cylinder_centers=np.zeros((B, max_cylinders, 3), np.float32),
        cylinder_radii=np.zeros((B, max_cylinders, 1), np.float32), cylinder_heights=np.zeros((B, max_cylinders, 1), np.float32),
        cylinder_quats=np.zeros((B, max_cylinders, 4), np.float32))
    out["cuboid_quats"][..., 0] = 1.0
    out["cylinder_quats"][..., 0] = 1.0
    for b, prims in enumerate(groups):
        nc = ny = 0
        for o in prims or []:
            if _is_cylinder(o):
                if ny >= max_cylinders:
                    raise ValueError(f"problem {b}: more than {max_cylinders} cylinders")
                out["cylinder_centers"][b, ny] = o.center
                out["cylinder_radii"][b, ny, 0], out["cylinder_heights"][b, ny, 0] = o.radius, o.height
                out["cylinder_quats"][b, ny] = _quat(o)
                ny += 1
            elif hasattr(o, "dims"):
                if nc >= max_cuboids:
                    raise ValueError(f"problem {b}: more than {max_cuboids} cuboids")
                out["cuboid_centers"][b, nc], out["cuboid_dims"][b, nc], out["cuboid_quats"][b, nc] = o.center, o.dims, _quat(o)
                nc += 1
            else:
                raise TypeError(f"problem {b}: unsupported primitive {type(o).__name__} (cuboids and cylinders only)")
    return out


def problems_to_soa(problems: Sequence[PlanningProblem], max_cuboids: int = 40, max_cylinders: int = 40) -> Dict[str, np.ndarray]:
    """PlanningProblem records -> engine inputs: ``q0`` [B,7], ``target`` [B,3,4] (right_gripper pose), the obstacle arrays
    (SCENE_KEYS), and the region-test volumes of metrics.py:365-384 as ``target_volume`` / ``negative_volumes`` dicts in the
    same layout (row counts = the largest count in the batch, at least 1)."""
    B = len(problems)
    out = primitives_to_soa([p.obstacles for p in problems], max_cuboids, max_cylinders)
    out["q0"] = np.stack([np.asarray(p.q0, dtype=np.float32).reshape(7) for p in problems])
    out["target"] = np.stack([np.asarray(p.target.matrix, dtype=np.float32)[:3] for p in problems])
    tv = [[p.target_volume] if p.target_volume is not None else [] for p in problems]
    nv = [list(p.target_negative_volumes or []) for p in problems]

    def counts(groups):
        return (max(1, max(sum(not _is_cylinder(o) for o in g) for g in groups)),
                max(1, max(sum(_is_cylinder(o) for o in g) for g in groups)))
    out["target_volume"] = primitives_to_soa(tv, *counts(tv))
    out["negative_volumes"] = primitives_to_soa(nv, *counts(nv))
    # problems that carry a sensed obstacle cloud (mpinets_types.py:44; filled by convert_primitive_problems_to_depth,
    # run_inference.py:194-257): padded [B, Pmax, 3] + valid counts for mpn_build_cloud_from_points
    clouds = [None if p.obstacle_point_cloud is None else np.asarray(p.obstacle_point_cloud, dtype=np.float32)[:, :3] for p in problems]
    if all(c is not None for c in clouds) and B > 0:
        pmax = max(len(c) for c in clouds)
        pts = np.zeros((B, pmax, 3), np.float32)
        for b, c in enumerate(clouds):
            pts[b, : len(c)] = c
        out["obstacle_points"] = pts
        out["obstacle_counts"] = np.array([len(c) for c in clouds], np.int32)
    assert out["q0"].shape == (B, 7)
    return out


def flatten_problem_set(problem_set: ProblemSet) -> List[tuple]:
    """[(env_type, problem_type, PlanningProblem), ...] in the iteration order of run_inference.py:460-468"""
    return [(env, kind, p) for env, kinds in problem_set.items() for kind, plist in kinds.items() for p in plist]


def soa_to_problems(soa: Dict[str, np.ndarray]) -> List[PlanningProblem]:
    """inverse of ``problems_to_soa`` for generated scenes (scenes.make_problems): one record per row, padding rows dropped"""
    out = []
    for b in range(soa["q0"].shape[0]):
        obs: Obstacles = []
        for m in range(soa["cuboid_dims"].shape[1]):
            if not np.isclose(soa["cuboid_dims"][b, m], 0).any():
                obs.append(Cuboid(soa["cuboid_centers"][b, m], soa["cuboid_dims"][b, m], soa["cuboid_quats"][b, m]))
        for m in range(soa["cylinder_radii"].shape[1]):
            r, h = float(np.ravel(soa["cylinder_radii"][b, m])[0]), float(np.ravel(soa["cylinder_heights"][b, m])[0])
            if not (np.isclose(r, 0) or np.isclose(h, 0)):
                obs.append(Cylinder(soa["cylinder_centers"][b, m], r, h, soa["cylinder_quats"][b, m]))
        tgt = np.eye(4); tgt[:3] = soa["target"][b]
        pose = SE3.from_matrix(tgt)
        out.append(PlanningProblem(target=pose, target_volume=Cuboid(pose.xyz, [0.1, 0.1, 0.1]), q0=soa["q0"][b].copy(), obstacles=obs))
    return out


__all__ = ["SE3", "SO3", "Cuboid", "Cylinder", "PlanningProblem", "ProblemSet", "problems_to_soa", "primitives_to_soa",
           "flatten_problem_set", "soa_to_problems", "SCENE_KEYS"]
