"""Synthetic PlanningProblem generators (host side, numpy) -- inputs for tests and bench.py.

The reference builds its scenes with PyBullet + IKFast (``mpinets/data_pipeline/environments/*``), neither of which
is available offline, so these generators reproduce only the *primitive distributions*:

* tabletop  -- ``tabletop_environment.py:215-324`` (tables) and ``:406-441`` (3..14 objects, 30 % cylinders), ``gen_data.py:618``
* cubby     -- ``cubby_environment.py:62-74,124-264`` (5 walls + optional centre wall + middle shelf, yaw +-10 deg)
* dresser   -- ``dresser_environment.py:198-223,967-1406`` (thin-board carcass, recursive splits, open drawers)

Start / goal configurations are uniform inside the joint limits shrunk by 5 % (stand-in for the IK'd candidates);
the target pose is FK(q_goal) in the ``right_gripper`` frame (``mpinets_types.py:39``).  Output arrays use the batch
keys of ``data_loader.py:206-235`` and zero-volume padding rows with unit quaternions (``data_loader.py:198-215``).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np

from .franka import REAL_JOINT_LIMITS, fk_reference_f64

SEED_BASE = 0x4D50694E
MAX_PRIMS = 40  # CUBOID_CUTOFF = CYLINDER_CUTOFF = 40 (gen_data.py:87-88)

Cub = Tuple[np.ndarray, np.ndarray, np.ndarray]   # centre, dims, quat(wxyz)
Cyl = Tuple[np.ndarray, float, float, np.ndarray]  # centre, radius, height, quat


def _yaw_quat(a: float) -> np.ndarray:
    return np.array([np.cos(a / 2), 0.0, 0.0, np.sin(a / 2)])


def _rot_z(a: float) -> np.ndarray:
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1.0]])


def tabletop(rng: np.random.RandomState) -> Tuple[List[Cub], List[Cyl]]:
    cubs: List[Cub] = []
    cyls: List[Cyl] = []
    h = rng.uniform(0, 0.4) if rng.uniform() < 0.65 else 0.0
    z, dz = (h - 0.02) / 2, h + 0.02
    x0, x1 = rng.uniform(0.275, 0.375), rng.uniform(1.275, 1.375)
    y1 = rng.uniform(1.5, 1.65)
    side = rng.uniform() < 0.5
    y0 = -rng.uniform(0.75, 1.0) if side else -rng.uniform(0.55, 0.75)
    scal = rng.uniform(0.55, 0.65)
    wy = y1 - y0
    task = (np.array([(x0 + x1) / 2, y0 + scal * wy / 2, z]), np.array([x1 - x0, scal * wy, dz]))
    free = (np.array([(x0 + x1) / 2, (y1 + y0 + scal * wy) / 2, z]), np.array([x1 - x0, wy - scal * wy, dz]))
    tables = [task]
    cubs += [(task[0], task[1], _yaw_quat(0)), (free[0], free[1], _yaw_quat(0))]
    if side:
        sy1, sy0 = -rng.uniform(0.275, 0.325), y0
        sx1 = x0
        sx0 = sx1 - rng.uniform(0.2, 1.375)
        s2 = rng.uniform(0.55, 0.65)
        wx = sx1 - sx0
        st = (np.array([sx1 - s2 * wx / 2, (sy0 + sy1) / 2, z]), np.array([s2 * wx, sy1 - sy0, dz]))
        sf = (np.array([(sx0 + sx1 - s2 * wx) / 2, (sy0 + sy1) / 2, z]), np.array([wx - s2 * wx, sy1 - sy0, dz]))
        tables.append(st)
        cubs += [(st[0], st[1], _yaw_quat(0)), (sf[0], sf[1], _yaw_quat(0))]
    mc = np.array([rng.uniform(-0.02, 0.02), rng.uniform(-0.02, 0.02), -0.01])
    cubs.append((mc, np.array([2 * (x0 - mc[0]), rng.uniform(0.9, 0.94), 0.02]), _yaw_quat(0)))
    placed: List[Tuple[float, float, float]] = []
    for _ in range(rng.randint(3, 15)):
        for _try in range(20):
            t = tables[rng.randint(len(tables))]
            x = rng.uniform(t[0][0] - t[1][0] / 2 + 0.08, t[0][0] + t[1][0] / 2 - 0.08)
            y = rng.uniform(t[0][1] - t[1][1] / 2 + 0.08, t[0][1] + t[1][1] / 2 - 0.08)
            if np.hypot(x, y) < 0.25:
                continue
            r = rng.uniform(0.05, 0.15)
            if all(np.hypot(x - px, y - py) > r + pr + 0.05 for px, py, pr in placed):
                break
        else:
            continue
        placed.append((x, y, r))
        if rng.uniform() < 0.3:
            hh = rng.uniform(0.05, 0.35)
            cyls.append((np.array([x, y, hh / 2 + h]), r, hh, _yaw_quat(0)))
        else:
            d = np.array([rng.uniform(0.05, 0.15), rng.uniform(0.05, 0.15), rng.uniform(0.05, 0.35)])
            cubs.append((np.array([x, y, d[2] / 2 + h]), d, _yaw_quat(rng.uniform(0, np.pi / 2))))
    return cubs, cyls


def cubby(rng: np.random.RandomState, merged: bool = False) -> Tuple[List[Cub], List[Cyl]]:
    th = rng.uniform(0.01, 0.03)
    W, D, H = rng.uniform(0.7, 1.1), rng.uniform(0.25, 0.45), rng.uniform(0.5, 0.9)
    zb = rng.uniform(0.05, 0.35)
    front = rng.uniform(0.55, 0.8)
    yaw = rng.uniform(-np.pi / 18, np.pi / 18)
    centre = np.array([front + D / 2, rng.uniform(-0.15, 0.15), zb + H / 2])
    boards = [
        (np.array([0, 0, -H / 2]), np.array([D, W, th])), (np.array([0, 0, H / 2]), np.array([D, W, th])),
        (np.array([0, -W / 2, 0]), np.array([D, th, H])), (np.array([0, W / 2, 0]), np.array([D, th, H])),
        (np.array([D / 2, 0, 0]), np.array([th, W, H])),
    ]
    shelf_z = rng.uniform(-0.1, 0.1) * H
    split_y = rng.uniform(-0.15, 0.15) * W
    if not (merged and rng.uniform() < 0.5):
        boards.append((np.array([0, 0, shelf_z]), np.array([D, W, th])))
    if not (merged and rng.uniform() < 0.5):
        boards.append((np.array([0, split_y, 0]), np.array([D, th, H])))
    R = _rot_z(yaw)
    return [(centre + R @ c, d, _yaw_quat(yaw)) for c, d in boards], []


def dresser(rng: np.random.RandomState) -> Tuple[List[Cub], List[Cyl]]:
    while True:
        W, D, H = rng.uniform(0.8, 1.2), rng.uniform(0.2, 0.4), rng.uniform(0.55, 0.85)
        yaw = rng.uniform(np.pi / 2 - np.pi / 3, np.pi / 2 + np.pi / 3)
        origin = np.array([rng.uniform(0.55, 0.75), rng.uniform(-0.1, 0.1), 0.0])
        t = 0.01
        boards = [(np.array([0, 0, t / 2]), np.array([W, D, t])), (np.array([0, 0, H - t / 2]), np.array([W, D, t])),
                  (np.array([-W / 2 + t / 2, 0, H / 2]), np.array([t, D, H])), (np.array([W / 2 - t / 2, 0, H / 2]), np.array([t, D, H])),
                  (np.array([0, D / 2 - t / 2, H / 2]), np.array([W, t, H]))]
        cells: List[Tuple[float, float, float, float]] = []

        def split(x0, x1, z0, z1, p):
            horizontal = (z1 - z0) > (x1 - x0)
            span = (z1 - z0) if horizontal else (x1 - x0)
            if span > 0.6 and rng.uniform() < p:
                c = rng.uniform(0.4, 0.6)
                if horizontal:
                    zc = z0 + c * span
                    boards.append((np.array([(x0 + x1) / 2, 0, zc]), np.array([x1 - x0, D, t])))
                    split(x0, x1, z0, zc, p * 0.8); split(x0, x1, zc, z1, p * 0.8)
                else:
                    xc = x0 + c * span
                    boards.append((np.array([xc, 0, (z0 + z1) / 2]), np.array([t, D, z1 - z0])))
                    split(x0, xc, z0, z1, p * 0.8); split(xc, x1, z0, z1, p * 0.8)
            else:
                cells.append((x0, x1, z0, z1))
        split(-W / 2 + t, W / 2 - t, t, H - t, 0.7)
        open_ids = set(rng.choice(len(cells), size=min(2, len(cells)), replace=False).tolist())
        for i, (x0, x1, z0, z1) in enumerate(cells):
            pull = 0.81 * D if i in open_ids else 0.0
            cw, ch, cx, cz = (x1 - x0) - 0.01, (z1 - z0) - 0.01, (x0 + x1) / 2, (z0 + z1) / 2
            y = -pull
            s = 0.004
            boards += [
                (np.array([cx, y - D / 2 + 0.0095, cz]), np.array([cw, 0.019, ch])),            # front
                (np.array([cx, y, cz - ch / 2 + s / 2]), np.array([cw, D - 0.03, s])),            # bottom
                (np.array([cx - cw / 2 + s / 2, y, cz]), np.array([s, D - 0.03, ch * 0.8])),      # sides
                (np.array([cx + cw / 2 - s / 2, y, cz]), np.array([s, D - 0.03, ch * 0.8])),
                (np.array([cx, y + D / 2 - 0.02, cz]), np.array([cw, s, ch * 0.8])),              # back
            ]
        if len(boards) < MAX_PRIMS:
            break
    R = _rot_z(yaw - np.pi / 2)
    return [(origin + R @ c, d, _yaw_quat(yaw - np.pi / 2)) for c, d in boards], []


GENERATORS = {"tabletop": tabletop, "cubby": cubby, "merged_cubby": lambda r: cubby(r, True), "dresser": dresser}


def make_problems(B: int, scene_types=("tabletop",), seed: int = SEED_BASE, problem0: int = 0,
                  max_cuboids: int = MAX_PRIMS, max_cylinders: int = MAX_PRIMS) -> Dict[str, np.ndarray]:
    """B seeded problems; global problem g = problem0 + i uses scene_types[g % len] and a RandomState keyed by (seed, g),
    so a shard generated with problem0 = rank*B equals the corresponding slice of one big batch."""
    out = dict(
        q0=np.zeros((B, 7), np.float32), q_goal=np.zeros((B, 7), np.float32), target=np.zeros((B, 3, 4), np.float32),
        cuboid_centers=np.zeros((B, max_cuboids, 3), np.float32), cuboid_dims=np.zeros((B, max_cuboids, 3), np.float32),
        cuboid_quats=np.zeros((B, max_cuboids, 4), np.float32), cylinder_centers=np.zeros((B, max_cylinders, 3), np.float32),
        cylinder_radii=np.zeros((B, max_cylinders, 1), np.float32), cylinder_heights=np.zeros((B, max_cylinders, 1), np.float32),
        cylinder_quats=np.zeros((B, max_cylinders, 4), np.float32), scene_type=np.zeros(B, np.int32),
    )
    out["cuboid_quats"][..., 0] = 1.0
    out["cylinder_quats"][..., 0] = 1.0
    lim = REAL_JOINT_LIMITS
    mid, half = lim.mean(axis=1), (lim[:, 1] - lim[:, 0]) / 2 * 0.95
    names = list(GENERATORS)
    for i in range(B):
        rng = np.random.RandomState((seed * 1000003 + (problem0 + i) * 7919) % (2 ** 32))
        st = scene_types[(problem0 + i) % len(scene_types)]   # depends on the GLOBAL problem index only
        cubs, cyls = GENERATORS[st](rng)
        cubs, cyls = cubs[:max_cuboids], cyls[:max_cylinders]
        out["scene_type"][i] = names.index(st)
        for m, (c, d, q) in enumerate(cubs):
            out["cuboid_centers"][i, m], out["cuboid_dims"][i, m], out["cuboid_quats"][i, m] = c, d, q
        for m, (c, r, h, q) in enumerate(cyls):
            out["cylinder_centers"][i, m], out["cylinder_radii"][i, m, 0] = c, r
            out["cylinder_heights"][i, m, 0], out["cylinder_quats"][i, m] = h, q
        q0 = mid + half * rng.uniform(-1, 1, size=7)
        qg = mid + half * rng.uniform(-1, 1, size=7)
        out["q0"][i], out["q_goal"][i] = q0, qg
        out["target"][i] = fk_reference_f64(qg)[1][:3]
    return out


SCENE_KEYS = ("cuboid_centers", "cuboid_dims", "cuboid_quats", "cylinder_centers", "cylinder_radii", "cylinder_heights",
              "cylinder_quats")


def config_problems(config: int, B: int, seed: int = SEED_BASE, problem0: int = 0) -> Dict[str, np.ndarray]:
    """Scene mixes of BASELINE.json configs: 1/2 tabletop; 3 cubby(+merged)+dresser; 4 mixed thirds."""
    mix = {1: ("tabletop",), 2: ("tabletop",), 3: ("cubby", "dresser", "merged_cubby", "dresser"),
           4: ("tabletop", "cubby", "dresser", "tabletop", "merged_cubby", "dresser")}[config]
    return make_problems(B, mix, seed, problem0)
