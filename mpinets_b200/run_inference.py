"""Batched counterpart of ``mpinets/run_inference.py`` (``calculate_metrics``, run_inference.py:426-516): every problem of a
``ProblemSet`` is stepped in lock-step on the GPU (``rollout_until_success`` semantics: at most 150 steps, per-problem stop at
1 cm / 15 degrees, run_inference.py:137-191) and evaluated by the device-side ``Evaluator``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .metrics import Evaluator
from .model import MAX_ROLLOUT_LENGTH, PRECISIONS, MotionPolicyNetwork
from .mpinets_types import PlanningProblem, ProblemSet, flatten_problem_set, problems_to_soa
from .scenes import SCENE_KEYS


def run_problems(mdl: MotionPolicyNetwork, problems: Sequence[PlanningProblem], device: Optional[torch.device] = None,
                 max_steps: int = MAX_ROLLOUT_LENGTH, dt: float = 0.08, evaluator: Optional[Evaluator] = None,
                 problem0: int = 0) -> Dict[str, torch.Tensor]:
    """clouds (make_point_cloud_from_primitives, run_inference.py:93-134) -> rollout_until_success -> evaluate_trajectory
    for the whole list at once.  Returns the trajectories, the number of poses of each, and the 16-column evaluation table."""
    device = device or torch.device("cuda", torch.cuda.current_device())
    soa = problems_to_soa(problems)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)  # noqa: E731
    scene = {k: dev(soa[k]) for k in SCENE_KEYS}
    q0, target = dev(soa["q0"]), dev(soa["target"])
    eng = mdl.sync_engine(device)
    if "obstacle_points" in soa:     # make_point_cloud_from_problem (run_inference.py:58-90): sensed obstacle clouds
        cloud = eng.build_cloud_from_points(q0, target, dev(soa["obstacle_points"]), dev(soa["obstacle_counts"]), problem0=problem0)
    else:                            # make_point_cloud_from_primitives (run_inference.py:93-134)
        cloud = eng.build_cloud(scene, q0, target, problem0=problem0)
    traj, metrics = eng.rollout(scene, cloud, q0, target, max_steps, early_exit=True, precision=PRECISIONS[mdl.precision])
    num_poses = (metrics[:, 2].to(torch.int32) + 1).contiguous()          # MPN_M_STEPS + the start configuration
    ev = evaluator or Evaluator(eng)
    table = ev.evaluate_trajectories(traj, dt, target, scene, target_volume={k: dev(v) for k, v in soa["target_volume"].items()},
                                     target_negative_volumes={k: dev(v) for k, v in soa["negative_volumes"].items()},
                                     num_poses=num_poses)
    return dict(trajectories=traj, num_poses=num_poses, eval=table, rollout_metrics=metrics)


def calculate_metrics(mdl: MotionPolicyNetwork, problem_set, device: Optional[torch.device] = None,
                      max_steps: int = MAX_ROLLOUT_LENGTH, environment_type: str = "all", problem_type: str = "all") -> Evaluator:
    """run_inference.calculate_metrics (run_inference.py:426-516): one metric group per (environment, problem type).
    ``problem_set``: a ProblemSet, or the path of a reference problem pickle (read as run_inference.py:460-468 does, through
    problem_io.load_problem_set -- no geometrout needed)."""
    if isinstance(problem_set, (str, bytes)) or hasattr(problem_set, "__fspath__"):
        from .problem_io import load_problem_set
        problem_set = load_problem_set(problem_set, environment_type, problem_type)
    ev = Evaluator()
    flat = flatten_problem_set(problem_set)
    groups: Dict[str, List[int]] = {}
    for i, (env, kind, _) in enumerate(flat):
        groups.setdefault(f"{env}, {kind}", []).append(i)
    offset = 0
    for key, idx in groups.items():
        ev.create_new_group(key)
        run_problems(mdl, [flat[i][2] for i in idx], device, max_steps, evaluator=ev, problem0=offset)
        offset += len(idx)
    return ev


# camera->world poses of the evaluation views (run_inference.py:215-243 builds SE3(xyz, quaternion).inverse = world->camera)
_EVAL_CAMERAS = {
    "dresser": ((0.08307640315968651, 1.986952324350807, 0.9996085854670145),
                (-0.10162310189063647, -0.06726290364234049, 0.5478233048853433, 0.8276702686337273)),
    "cubby": ((0.08307640315968651, 1.986952324350807, 0.9996085854670145),
              (-0.10162310189063647, -0.06726290364234049, 0.5478233048853433, 0.8276702686337273)),
    "tabletop": ((1.5031788593125708, -1.817341016921562, 1.278088299149147),
                 (0.8687241016192855, 0.4180885960330695, 0.11516106409944685, 0.23928704613569252)),
}


def eval_camera(environment_type: str) -> np.ndarray:
    """camera->world [3,4] of the view the reference evaluates `environment_type` with (run_inference.py:215-247)"""
    from .mpinets_types import SE3
    for key, (xyz, quat) in _EVAL_CAMERAS.items():
        if key in environment_type:
            return SE3(xyz, quat).matrix[:3].astype(np.float32)
    raise NotImplementedError(f"Camera angle is not implemented for environment type: {environment_type}")   # run_inference.py:244-247


def convert_primitive_problems_to_depth(problems: ProblemSet, device: Optional[torch.device] = None, width: int = 640,
                                        height: int = 480, fov_y_deg: float = 60.0, near: float = 0.01, far: float = 10.0,
                                        engine=None):
    """run_inference.convert_primitive_problems_to_depth (run_inference.py:194-257), in place: every problem gets an
    ``obstacle_point_cloud`` seen from its environment's evaluation camera.  The reference renders with Bullet and removes
    the robot; here the primitives are ray-cast analytically on the GPU (``mpn_render_depth_cloud``), one launch per
    environment type."""
    from .runtime import get_engine
    device = device or torch.device("cuda", torch.cuda.current_device())
    eng = engine or get_engine(device)
    for environment_type, scene_sets in problems.items():
        cam = torch.from_numpy(eval_camera(environment_type)).to(device).contiguous()
        plist = [p for problem_set in scene_sets.values() for p in problem_set]
        if not plist:
            continue
        soa = problems_to_soa(plist)
        scene = {k: torch.from_numpy(np.ascontiguousarray(soa[k])).to(device) for k in SCENE_KEYS}
        pts, cnt = eng.render_depth_cloud(scene, cam, width, height, fov_y_deg, near, far)
        pts, cnt = pts.cpu().numpy(), cnt.cpu().numpy()
        for i, p in enumerate(plist):
            p.obstacle_point_cloud = pts[i, :cnt[i]].copy()
