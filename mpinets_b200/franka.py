"""Franka Panda constants that the engine takes as *data inputs* (robot tables).

The reference obtains these from the un-vendored ``robofin`` v0.0.1 (``/root/reference/docker/Dockerfile:153``):
``FrankaRealRobot.JOINT_LIMITS`` (used by every (un)normalise call, ``mpinets/utils.py:50,84,150,192``),
``FrankaCollisionSampler``'s sphere table (``mpinets/model.py:269-271,300``) and ``FrankaSampler``'s cached
mesh-surface points (``mpinets/model.py:250``, ``mpinets/run_inference.py:111-116``).  Meshes and URDF are not in
the reference tree, so:

* the sphere table below is the nearest in-tree statement, ``config/franka_robot_description.yaml:57-182``
  (57 spheres; fingertip spheres are re-expressed in the finger-link frames);
* the canonical link-point table is *synthetic* (points on those spheres' surfaces, area-proportional,
  seeded) -- a real robofin table can be dropped in through ``RobotTables`` without touching any kernel.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

LINK_NAMES = (
    "panda_link0", "panda_link1", "panda_link2", "panda_link3", "panda_link4", "panda_link5",
    "panda_link6", "panda_link7", "panda_hand", "panda_leftfinger", "panda_rightfinger",
)
NUM_LINKS = len(LINK_NAMES)
DOF = 7
DEFAULT_PRISMATIC_VALUE = 0.025

# Published Franka limits (robofin FrankaRobot.JOINT_LIMITS)
JOINT_LIMITS = np.array(
    [(-2.8973, 2.8973), (-1.7628, 1.7628), (-2.8973, 2.8973), (-3.0718, -0.0698),
     (-2.8973, 2.8973), (-0.0175, 3.7525), (-2.8973, 2.8973)], dtype=np.float64)
# robofin FrankaRealRobot.JOINT_LIMITS (empirical; joint 6 lower bound differs) -- the table the reference
# always normalises with (mpinets/utils.py:118-125,235-242).
REAL_JOINT_LIMITS = np.array(
    [(-2.8973, 2.8973), (-1.7628, 1.7628), (-2.8973, 2.8973), (-3.0718, -0.0698),
     (-2.8973, 2.8973), (0.5445, 3.7525), (-2.8973, 2.8973)], dtype=np.float64)

# interactive_demo/mpinets_ros/nodes/interaction_node.py:54-75 -- FK known-answer pair (right_gripper frame)
NEUTRAL_CONFIG = np.array(
    [-0.01779206, -0.76012354, 0.01978261, -2.34205014, 0.02984053, 1.54119353, 0.75344866])
NEUTRAL_TARGET_XYZ = np.array([0.30649957, 0.00728735, 0.48663767])
NEUTRAL_TARGET_XYZW = np.array([-0.01424194, 0.99965734, 0.00846602, -0.02026548])

_FT = 0.045  # fingertip frame offset along finger z
# (link index, centre in link frame, radius) -- config/franka_robot_description.yaml:57-182
_SPHERES = [
    (0, (0.0, 0.0, 0.05), 0.08),
    (1, (0.0, -0.08, 0.0), 0.06), (1, (0.0, -0.03, 0.0), 0.06), (1, (0.0, 0.0, -0.12), 0.06), (1, (0.0, 0.0, -0.17), 0.06),
    (2, (0.0, 0.0, 0.03), 0.06), (2, (0.0, 0.0, 0.08), 0.06), (2, (0.0, -0.12, 0.0), 0.06), (2, (0.0, -0.17, 0.0), 0.06),
    (3, (0.0, 0.0, -0.06), 0.05), (3, (0.0, 0.0, -0.1), 0.06), (3, (0.08, 0.06, 0.0), 0.055), (3, (0.08, 0.02, 0.0), 0.055),
    (4, (0.0, 0.0, 0.02), 0.055), (4, (0.0, 0.0, 0.06), 0.055), (4, (-0.08, 0.095, 0.0), 0.06), (4, (-0.08, 0.06, 0.0), 0.055),
    (5, (0.0, 0.055, 0.0), 0.06), (5, (0.0, 0.075, 0.0), 0.06), (5, (0.0, 0.0, -0.22), 0.06), (5, (0.0, 0.05, -0.18), 0.05),
    (5, (0.01, 0.08, -0.14), 0.025), (5, (0.01, 0.085, -0.11), 0.025), (5, (0.01, 0.09, -0.08), 0.025), (5, (0.01, 0.095, -0.05), 0.025),
    (5, (-0.01, 0.08, -0.14), 0.025), (5, (-0.01, 0.085, -0.11), 0.025), (5, (-0.01, 0.09, -0.08), 0.025), (5, (-0.01, 0.095, -0.05), 0.025),
    (6, (0.0, 0.0, 0.0), 0.06), (6, (0.08, 0.03, 0.0), 0.06), (6, (0.08, -0.01, 0.0), 0.06),
    (7, (0.0, 0.0, 0.07), 0.05), (7, (0.02, 0.04, 0.08), 0.025), (7, (0.04, 0.02, 0.08), 0.025), (7, (0.04, 0.06, 0.085), 0.02), (7, (0.06, 0.04, 0.085), 0.02),
] + [(8, (0.0, y, z), r) for z, r in ((0.01, 0.028), (0.03, 0.026), (0.05, 0.024))
     for y in (-0.075, -0.045, -0.015, 0.015, 0.045, 0.075)] + [
    (9, (0.0, 0.0075, _FT), 0.0108),
    (10, (0.0, -0.0075, _FT), 0.0108),
]


@dataclass
class RobotTables:
    """All per-robot data the engine needs; every array is float32 / int32, C-contiguous."""

    joint_limits: np.ndarray     # [7,2]  limits used for (un)normalisation
    link_points: np.ndarray      # [P,3]  canonical surface points, link frame
    link_ids: np.ndarray         # [P]    link index of each canonical point
    ee_points: np.ndarray        # [Pe,3] gripper (hand+fingers) points in the right_gripper frame
    sphere_centers: np.ndarray   # [S,3]  link frame
    sphere_radii: np.ndarray     # [S]
    sphere_links: np.ndarray     # [S]
    prismatic: float = DEFAULT_PRISMATIC_VALUE


def collision_spheres(with_base_link: bool = True):
    """(centres[S,3], radii[S], links[S]); ``with_base_link=False`` drops panda_link0 (model.py:269-271)."""
    rows = [s for s in _SPHERES if with_base_link or s[0] != 0]
    c = np.array([s[1] for s in rows], dtype=np.float32)
    r = np.array([s[2] for s in rows], dtype=np.float32)
    l = np.array([s[0] for s in rows], dtype=np.int32)
    return c, r, l


def _sphere_surface(rng: np.random.RandomState, n: int) -> np.ndarray:
    v = rng.normal(size=(n, 3))
    return v / np.linalg.norm(v, axis=1, keepdims=True)


def synthetic_link_points(num_points: int = 4096, seed: int = 0, with_base_link: bool = True):
    """Canonical per-link surface points (synthetic stand-in for robofin's trimesh samples).

    Points lie on the union-of-spheres surface of each link, allocated proportionally to sphere area;
    points buried inside a sibling sphere of the same link are rejected so the cloud looks like a hull.
    """
    rng = np.random.RandomState(seed)
    rows = [s for s in _SPHERES if with_base_link or s[0] != 0]
    areas = np.array([4 * np.pi * s[2] ** 2 for s in rows])
    alloc = np.floor(areas / areas.sum() * num_points).astype(int)
    alloc[np.argsort(-areas)[: num_points - alloc.sum()]] += 1
    pts, ids = [], []
    for (link, c, r), n in zip(rows, alloc):
        sib = [(np.array(c2), r2) for (l2, c2, r2) in rows if l2 == link and (c2 != c or r2 != r)]
        got = np.zeros((0, 3))
        tries = 0
        while len(got) < n:
            cand = np.array(c) + r * _sphere_surface(rng, max(2 * n, 16))
            keep = np.ones(len(cand), dtype=bool)
            if tries < 8:
                for c2, r2 in sib:
                    keep &= np.linalg.norm(cand - c2, axis=1) >= r2
            got = np.concatenate([got, cand[keep]], axis=0)
            tries += 1
        pts.append(got[:n])
        ids.append(np.full(n, link, dtype=np.int32))
    return np.concatenate(pts).astype(np.float32), np.concatenate(ids)


def ee_points_from_links(link_points: np.ndarray, link_ids: np.ndarray, prismatic: float = DEFAULT_PRISMATIC_VALUE):
    """Gripper points (hand + both fingers) expressed in the right_gripper frame = hand * Tz(0.1) * Rz(pi)."""
    out = []
    for link, off in ((8, (0.0, 0.0, 0.0)), (9, (0.0, prismatic, 0.0584)), (10, (0.0, -prismatic, 0.0584))):
        p = link_points[link_ids == link].astype(np.float64) + np.array(off)
        out.append(np.stack([-p[:, 0], -p[:, 1], p[:, 2] - 0.1], axis=1))
    return np.concatenate(out).astype(np.float32)


def default_tables(num_points: int = 4096, seed: int = 0, with_base_link_spheres: bool = False,
                   use_real_constraints: bool = True) -> RobotTables:
    """Tables matching the validation sweep: FrankaSampler(with_base_link=True) points,
    FrankaCollisionSampler(with_base_link=False) spheres (model.py:267-271), real-robot limits."""
    lp, lid = synthetic_link_points(num_points, seed)
    c, r, l = collision_spheres(with_base_link_spheres)
    lim = (REAL_JOINT_LIMITS if use_real_constraints else JOINT_LIMITS).astype(np.float32)
    return RobotTables(joint_limits=np.ascontiguousarray(lim), link_points=lp, link_ids=lid,
                       ee_points=ee_points_from_links(lp, lid), sphere_centers=c, sphere_radii=r, sphere_links=l)


def fk_reference_f64(q: np.ndarray, prismatic: float = DEFAULT_PRISMATIC_VALUE):
    """Float64 FK (host-side helper for scene generation / tests): returns (frames[11,4,4], right_gripper[4,4])."""
    def T(xyz=(0, 0, 0), roll=0.0, yaw=0.0):
        cr, sr, cy, sy = np.cos(roll), np.sin(roll), np.cos(yaw), np.sin(yaw)
        Rx = np.array([[1, 0, 0], [0, cr, -sr], [0, sr, cr]])
        Rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1]])
        M = np.eye(4)
        M[:3, :3] = Rx @ Rz
        M[:3, 3] = xyz
        return M
    h = np.pi / 2
    origins = [((0, 0, 0.333), 0.0), ((0, 0, 0), -h), ((0, -0.316, 0), h), ((0.0825, 0, 0), h),
               ((-0.0825, 0.384, 0), -h), ((0, 0, 0), h), ((0.088, 0, 0), h)]
    frames = [np.eye(4)]
    for (xyz, roll), qi in zip(origins, q):
        frames.append(frames[-1] @ T(xyz, roll, qi))
    hand = frames[7] @ T((0, 0, 0.107), 0.0, -np.pi / 4)
    frames.append(hand)
    frames.append(hand @ T((0, prismatic, 0.0584)))
    frames.append(hand @ T((0, -prismatic, 0.0584)))
    grip = hand @ T((0, 0, 0.1), 0.0, np.pi)
    return np.stack(frames), grip
