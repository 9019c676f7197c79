#!/usr/bin/env python
"""Per-step latency of the rollout at SMALL batches (the reference's real callers are B = 1: run_inference.py:268-303 assumes 80 ms
per step, planning_node.py:78-151), eager launches vs one CUDA graph of the whole T-step rollout:
    python scripts/latency_small_batch.py [T] > profiles/r2_latency_small_batch.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mpinets_b200 import scenes, _lib
from mpinets_b200.engine import Engine
from oracle import oracle as O   # weights init only

T = int(sys.argv[1]) if len(sys.argv) > 1 else 50
eng = Engine()
eng.load_state_dict(O.reference_state_dict(0))
eng.reserve(256)
out = {"T": T, "reference_assumed_ms_per_step": 80.0, "rows": []}
for B in (1, 16, 256):
    p = scenes.config_problems(2, B)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    sc = {k: dev(p[k]) for k in scenes.SCENE_KEYS}
    q0, tg = dev(p["q0"]), dev(p["target"])
    for name in ("bf16x3", "bf16", "fp32"):
        prec = _lib.PRECISIONS[name]
        cloud0 = eng.build_cloud(sc, q0, tg)
        cloud = cloud0.clone()
        traj = torch.empty(B, T + 1, 7, device="cuda"); metrics = torch.empty(B, 8, device="cuda")
        run = lambda: eng.rollout(sc, cloud, q0, tg, T, check_every_step=True, precision=prec, traj=traj, metrics=metrics)
        run(); torch.cuda.synchronize()
        ms = []
        for _ in range(5):
            cloud.copy_(cloud0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            l0 = eng.launch_count
            t0 = time.perf_counter(); e0.record(); run(); e1.record(); host_ms = 1e3 * (time.perf_counter() - t0)
            torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
        launches = eng.launch_count - l0
        eager = float(np.median(ms))
        ref_traj = traj.clone()
        row = {"B": B, "mode": name, "eager_ms_per_step": eager / T, "host_enqueue_ms_per_step": host_ms / T, "launches_per_step": launches / T}
        try:
            g = torch.cuda.CUDAGraph()
            cloud.copy_(cloud0)
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                run()
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            cloud.copy_(cloud0)
            with torch.cuda.graph(g):
                run()
            gm = []
            for _ in range(5):
                cloud.copy_(cloud0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); gm.append(e0.elapsed_time(e1))
            row["graph_ms_per_step"] = float(np.median(gm)) / T
            row["graph_equals_eager"] = bool(torch.equal(traj, ref_traj))
        except Exception as e:
            row["graph_error"] = repr(e)[:300]
        row["speedup_vs_reference_assumption"] = 80.0 / min(row.get("graph_ms_per_step", 1e9), row["eager_ms_per_step"])
        out["rows"].append(row)
        print(row, file=sys.stderr)
print(json.dumps(out, indent=1))
