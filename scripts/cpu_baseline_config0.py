#!/usr/bin/env python
"""CPU baseline of BASELINE.json configs[0] exactly as SURVEY.md section 8d words it, run in the BUILD container (the only place
/root/reference exists):  python scripts/cpu_baseline_config0.py > profiles/r2_cpu_baseline_config0.json

One tabletop PlanningProblem: build the 4096-point obstacle cloud, then the link-sphere SDF sweep of 50 poses (joint-space
interpolation between two in-limit configurations), no network.  Timed with
  * the oracle port (oracle/mpn_oracle.c) at 1 thread and on all host cores  -- the same code bench.py's cpu_baseline.config0 times
    on the GPU box, and
  * the REAL reference classes TorchCuboids / TorchCylinders.sdf_sequence (mpinets/geometry.py:290-347,509-568) imported from
    /root/reference with the geometrout stub of tests/golden/make_golden.py, fed the oracle's sphere centres, reduced like
    model.py:301-312 -- at torch.set_num_threads(1) and at all cores; >= 20 repetitions, median.
The reference's cloud builder (geometrout sample_surface) and sphere FK (robofin) are un-vendored, so only the SDF half has a
"reference" timing."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

import bench  # noqa: E402
from mpinets_b200 import franka, scenes  # noqa: E402
from oracle import oracle as O  # noqa: E402


def reference_sweep(geo, p, centers, radii, threads, reps):
    torch.set_num_threads(threads)
    t = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in p.items() if k.startswith(("cuboid", "cylinder"))}
    groups = [(float(r), np.nonzero(radii == r)[0]) for r in np.unique(radii)]   # compute_spheres returns the spheres grouped by radius

    def run():
        cub = geo.TorchCuboids(t["cuboid_centers"], t["cuboid_dims"], t["cuboid_quats"])
        cyl = geo.TorchCylinders(t["cylinder_centers"], t["cylinder_radii"].reshape(1, -1, 1), t["cylinder_heights"].reshape(1, -1, 1),
                                 t["cylinder_quats"])
        has = torch.zeros(1, dtype=torch.bool)
        for r, idx in groups:                                             # model.py:301-312
            seq = torch.from_numpy(centers[:, :, idx])                      # [1, 50, n_r, 3]
            sdf = torch.minimum(cub.sdf_sequence(seq), cyl.sdf_sequence(seq))
            has = torch.logical_or(torch.any(sdf.reshape(1, -1) <= r, dim=-1), has)
        return bool(has[0])

    flag = run()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        run()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), flag


def main():
    reps = 25
    out = {"oracle_port": bench.cpu_config0(reps), "host": {"cores": os.cpu_count(), "torch": torch.__version__}}
    if os.path.isdir("/root/reference"):
        from make_golden import load_reference_geometry
        geo = load_reference_geometry()
        tables = franka.default_tables()
        P = scenes.config_problems(2, 1)
        p = {k: v[:1] for k, v in P.items() if isinstance(v, np.ndarray)}
        w = np.linspace(0.0, 1.0, 50, dtype=np.float32)[None, :, None]
        traj = (P["q0"][:1, None, :] * (1 - w) + P["q_goal"][:1, None, :] * w).astype(np.float32)
        centers = O.spheres(traj[0], tables)[None]                         # [1, 50, S, 3]
        oflag = bool(O.sweep_flags(p, traj, tables)[0][0])
        m1, f1 = reference_sweep(geo, p, centers, tables.sphere_radii, 1, reps)
        mall, f2 = reference_sweep(geo, p, centers, tables.sphere_radii, os.cpu_count(), reps)
        out["reference_classes"] = {"kind": "reference", "what": "TorchCuboids/TorchCylinders.sdf_sequence + has_collision reduction (model.py:301-312) "
                                                                    "on the oracle's sphere centres, 1 problem x 50 poses",
                                    "sweep_50_poses_ms_1_thread": 1e3 * m1, "pose_checks_per_s_1_thread": 50.0 / m1,
                                    "sweep_50_poses_ms_all_cores": 1e3 * mall, "pose_checks_per_s_all_cores": 50.0 / mall,
                                    "collision_flag": f1, "flag_equals_oracle": f1 == oflag and f2 == oflag, "reps": reps}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
