"""Batched stand-in for mpinets/metrics.py:Evaluator over the C ABI (mpn_evaluate).

The reference evaluates one trajectory at a time on the CPU (Bullet collision checks, numpy FK, metrics.py:436-523);
here B trajectories are evaluated by one kernel launch and appended to the current metric group under the same keys
`add_metric` uses (metrics.py:470-523), so `Evaluator.metrics(group)` aggregates exactly like the reference's
(metrics.py:566-664).  SPARC smoothness (metrics.py:387-409) runs on the device too (`mpn_sparc`; the speed profiles are
formed with torch ops from the trajectory and one batched FK).  Not reproduced: Bullet collision depths; `collision` is the
validation sphere sweep of model.py:293-314 and `self_collision` a sphere-sphere stand-in (see include/mpinets_b200.h).
"""
from __future__ import annotations

from typing import Any, Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .runtime import get_engine


def percent_true(arr: Sequence) -> float:
    """metrics.py:50-57"""
    return 100 * np.count_nonzero(arr) / len(arr)


class Evaluator:
    def __init__(self, engine=None):
        self.engine = engine
        self.groups: Dict[str, Dict[str, Any]] = {}
        self.current_group: Optional[Dict[str, Any]] = None
        self.current_group_key: Optional[str] = None

    def create_new_group(self, key: str):
        """metrics.py:149-158"""
        self.groups[key] = {}
        self.current_group_key = key
        self.current_group = self.groups[key]

    def evaluate_trajectories(self, trajectories: torch.Tensor, dt: float, target: torch.Tensor, obstacles: Dict[str, torch.Tensor],
                              target_volume: Optional[Dict[str, torch.Tensor]] = None,
                              target_negative_volumes: Optional[Dict[str, torch.Tensor]] = None,
                              time: Optional[Sequence[float]] = None, num_poses: Optional[torch.Tensor] = None) -> torch.Tensor:
        """evaluate_trajectory (metrics.py:436-523) for a batch: trajectories [B,T+1,7] CUDA fp32 (mpn_rollout's buffer),
        target [B,3,4] right_gripper poses, obstacles = the scene dict the rollout used.  Returns the raw [B,16] table and
        appends every column to the current group."""
        if self.current_group is None:
            self.create_new_group("default")
        eng = self.engine or get_engine(trajectories.device)
        table = eng.evaluate(obstacles, trajectories, target, num_poses=num_poses, target_volume=target_volume,
                             negative_volumes=target_negative_volumes)
        host = table.cpu().numpy()
        B = host.shape[0]
        config_sparc, eff_sparc = self.calculate_smoothness(trajectories, dt, num_poses, eng)
        g = self.current_group
        booleans = ("collision", "joint_limit_violation", "self_collision", "physical_violations", "success")
        for i, name in enumerate(_lib.EVAL_COLUMNS):
            col = host[:, i]
            vals = [bool(v) for v in col] if name in booleans else [float(v) for v in col]
            if name == "num_steps":
                vals = [int(v) for v in col]
            g[name] = g.get(name, []) + vals
        g["time"] = g.get("time", []) + ([float(t) for t in time] if time is not None else [float(dt) * n for n in host[:, 10]])
        g["collision_depths"] = g.get("collision_depths", []) + [[float(d)] if d > 0 else [] for d in host[:, 13]]
        g["config_smoothness"] = g.get("config_smoothness", []) + [float(v) for v in config_sparc.cpu().numpy()]
        g["eff_smoothness"] = g.get("eff_smoothness", []) + [float(v) for v in eff_sparc.cpu().numpy()]
        return table

    @staticmethod
    def calculate_smoothness(trajectories: torch.Tensor, dt: float, num_poses: Optional[torch.Tensor] = None, engine=None):
        """metrics.py:387-409 for a batch: SPARC of the configuration-space speed |dq|/dt and of the end-effector speed
        |d xyz(right_gripper)|/dt; returns (config_sparc [B], eff_sparc [B])"""
        eng = engine or get_engine(trajectories.device)
        B, T1, _ = trajectories.shape
        if T1 < 2:
            z = torch.zeros(B, device=trajectories.device)
            return z, z
        cfg_speed = (torch.linalg.norm(torch.diff(trajectories, dim=1), dim=2) / dt).contiguous()
        _, eef = eng.fk(trajectories.reshape(B * T1, 7).contiguous())
        pos = eef[:, :, 3].reshape(B, T1, 3)
        eff_speed = (torch.linalg.norm(torch.diff(pos, dim=1), dim=2) / dt).contiguous()
        n = None if num_poses is None else (num_poses - 1).clamp(min=0).to(torch.int32).contiguous()
        return eng.sparc(cfg_speed, 1.0 / dt, n), eng.sparc(eff_speed, 1.0 / dt, n)

    @staticmethod
    def metrics(group: Dict[str, Any]) -> Dict[str, Any]:
        """metrics.py:566-664 (same keys)"""
        success = np.asarray(group["success"], dtype=bool)
        pos, ori = np.asarray(group["position_error"]), np.asarray(group["orientation_error"])
        times, steps = np.asarray(group["time"]), np.asarray(group["num_steps"])
        ppl, opl = np.asarray(group["eff_position_path_length"]), np.asarray(group["eff_orientation_path_length"])
        depths = np.asarray([d for ds in group["collision_depths"] for d in ds])

        def mean_std(a):
            return (float(np.mean(a)), float(np.std(a))) if len(a) else (float("nan"), float("nan"))

        return {
            "success": percent_true(success),
            "total": len(success),
            "skips": 0,
            "time": mean_std(times[success]),
            "step time": mean_std(times[success] / steps[success]),
            "env collision": percent_true(group["collision"]),
            "self collision": percent_true(group["self_collision"]),
            "joint violation": percent_true(group["joint_limit_violation"]),
            "physical violations": percent_true(group["physical_violations"]),
            "average collision depth": 100 * float(np.mean(depths)) if len(depths) else float("nan"),
            "median collision depth": 100 * float(np.median(depths)) if len(depths) else float("nan"),
            "1 cm": percent_true(pos < 1),
            "5 cm": percent_true(pos < 5),
            "15 deg": percent_true(ori < 15),
            "30 deg": percent_true(ori < 30),
            "165 deg": percent_true(ori > 165),
            "is smooth": percent_true(np.logical_and(np.asarray(group["config_smoothness"]) < -1.6,
                                                     np.asarray(group["eff_smoothness"]) < -1.6)),
            "average config sparc": float(np.mean(group["config_smoothness"])),
            "average eff sparc": float(np.mean(group["eff_smoothness"])),
            "eff position path length": mean_std(ppl[success]),
            "eff orientation path length": mean_std(opl[success]),
        }

    @staticmethod
    def print_metrics(group: Dict[str, Any]):
        """metrics.py:666-706"""
        m = Evaluator.metrics(group)
        print(f"Total problems: {m['total']}")
        print(f"% Success: {m['success']:4.2f}")
        print(f"% Within 1cm: {m['1 cm']:4.2f}")
        print(f"% Within 5cm: {m['5 cm']:4.2f}")
        print(f"% Within 15deg: {m['15 deg']:4.2f}")
        print(f"% Within 30deg: {m['30 deg']:4.2f}")
        print(f"% With Environment Collision: {m['env collision']:4.2f}")
        print(f"% With Self Collision: {m['self collision']:4.2f}")
        print(f"% With Joint Limit Violations: {m['joint violation']:4.2f}")
        print(f"% With Physical Violations: {m['physical violations']:4.2f}")
        print(f"Average Config SPARC: {m['average config sparc']:4.2f}")
        print(f"Average End Eff SPARC: {m['average eff sparc']:4.2f}")
        print(f"% Smooth: {m['is smooth']:4.2f}")
        print(f"Average End Eff Position Path Length: {m['eff position path length'][0]:4.2f} ± {m['eff position path length'][1]:4.2f}")
        print(f"Average End Eff Orientation Path Length: {m['eff orientation path length'][0]:4.2f} ± {m['eff orientation path length'][1]:4.2f}")

    def print_group_metrics(self, key: Optional[str] = None):
        """metrics.py:737-746"""
        self.print_metrics(self.current_group if key is None else self.groups[key])

    def print_overall_metrics(self):
        """metrics.py:748-760: all groups concatenated"""
        merged: Dict[str, Any] = {}
        for g in self.groups.values():
            for k, v in g.items():
                merged[k] = merged.get(k, []) + list(v)
        self.print_metrics(merged)
