#!/usr/bin/env python
"""bench.py -- env steps/sec of the batched MPiNets rollout hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--precision bf16|fp32]

One "step" = one lock-step policy step over the whole per-GPU batch of PlanningProblems: PointNet++ encoder (FPS,
ball query, grouping, shared MLPs) + delta-q head + clamp/unnormalise + FK + robot-surface resample into the cloud +
link-sphere SDF collision check.  N = 1 workload = BASELINE.json configs[1] (4096 tabletop problems, 6272-point
clouds); N > 1 shards 4096 problems per GPU (weak scaling, no data-path collective, one NCCL all-gather of the
metrics table at the end).

Printed JSON (rank 0): value = problems*K / device time of K steps with everything resident in HBM; e2e = the same
job through the public API with HOST problem buffers (pinned H2D of the problem SoA, cloud build, K steps, D2H of
trajectories + metrics inside the timed region); roofline = dominant kernel vs MEASURED_PEAKS.json; cpu_baseline =
the CPU oracle port timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env steps/sec (4096-pt cloud, 32k probs) at 1/2/4/8 B200; collision-flag match"
PROBLEMS_PER_GPU = 4096
# algorithmic work per env step per problem (SURVEY.md section 8d / DESIGN.md)
FLOP_PER_STEP = {"sa1": 2 * 553_648_128, "sa2": 2 * 945_815_552, "sa3": 2 * 117_571_584, "fc": 2 * 16_777_216 + 0,
                 "heads": 2 * 1_291_904}
FLOP_TOTAL = 2 * 1_635_159_136
BYTES_PER_STEP = {"sample_robot": 2048 * 16 + 11 * 48, "sweep": 28 + 3200 + 1, "build_cloud": 6272 * 16,
                  "fps1": 6272 * 16 + 512 * 16, "fps2": 512 * 12 + 128 * 16}


def load_traffic():
    """per-problem DRAM bytes (read + write) of each stage's dominant kernel from the committed ncu --set full capture"""
    p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    return json.load(open(p)) if os.path.exists(p) else {}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); power.append(float(r[3]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


class CpuPort:
    """The CPU port of the same step (oracle/): torch-CPU MLPs on the host threads + C geometry.  Problems, weights and the
    t = 0 clouds are prepared once (untimed, like the GPU arm's resident inputs); run() times rollout + collision sweep."""

    def __init__(self, n_problems: int, seed: int):
        import torch
        from mpinets_b200 import scenes, franka
        from oracle import oracle as O
        if torch.get_num_threads() == 1 and (os.cpu_count() or 1) > 2:   # torchrun pins OMP_NUM_THREADS=1: undo it for the CPU arm
            torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))      # (physical cores; SMT siblings only slow the GEMMs down)
        self.O, self.n, self.seed, self.threads = O, n_problems, seed, torch.get_num_threads()
        self.tables = franka.default_tables()
        self.sd = O.reference_state_dict(0)
        self.p = scenes.config_problems(2, n_problems)
        self.cloud = O.build_cloud(self.p["q0"], self.p["target"], self.p, self.tables, seed)
        self.qn = O.normalize(self.p["q0"], self.tables.joint_limits)

    def run(self, steps: int) -> float:
        t0 = time.perf_counter()
        traj = self.O.rollout(self.sd, self.cloud.copy(), self.qn, self.tables, steps, self.seed, emulate_bf16=False)
        self.O.sweep_flags(self.p, traj, self.tables)
        return time.perf_counter() - t0


def cpu_baseline(seed: int, budget_s: float = 10.0, n_problems: int = 16, steps: int = 4, max_calls: int = 8):
    """bounded sample of the bench workload on the host cores: calls of (n_problems x steps) until ~budget_s of CPU work"""
    port = CpuPort(n_problems, seed)
    port.run(1)   # warm-up (thread pools, page faults)
    total, calls = 0.0, 0
    while calls < max_calls and (calls == 0 or total < budget_s):
        total += port.run(steps)
        calls += 1
    done = calls * n_problems * steps
    return {"value": done / total, "unit": "env steps/s", "cores": port.threads, "kind": "port",
            "sample": f"{calls} x ({n_problems} problems x {steps} steps) of the same workload = {done} problem-steps in {total:.1f} s "
                      f"(cloud build excluded), torch-CPU fp32 MLPs + C oracle geometry"}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference cannot be imported
    (pointnet2_ops is CUDA-only; robofin/geometrout are not installable offline), so this times the oracle port on all host
    threads.  One bench step = one lock-step env step of a bounded sample of the workload (16 problems)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = 16
    port = CpuPort(n, 1)
    for _ in range(max(args.warmup, 1)):
        port.run(1)
    total = 0.0
    for _ in range(args.steps):
        total += port.run(1)
    value = n * args.steps / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "env steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": "configs[1]: tabletop problems, 6272-pt clouds, lock-step policy rollout + SDF sweep",
                       "sample": f"{n} problems x 1 step per bench step"},
            "cpu_baseline": {"value": value, "unit": "env steps/s", "cores": port.threads, "kind": "port",
                             "sample": f"{n} problems x {args.steps} steps (cloud build excluded), torch-CPU MLPs + C oracle geometry"},
            "e2e": {"value": value, "unit": "env steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    # keep stdout to the single JSON line: NCCL prints its version banner to stdout at NCCL_DEBUG=VERSION / WARN / INFO
    os.environ.pop("NCCL_DEBUG", None)
    if os.environ.get("MPN_NCCL_DEBUG"):
        os.environ["NCCL_DEBUG"] = os.environ["MPN_NCCL_DEBUG"]
        os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/mpn_nccl_%h_%p.log")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--precision", default=os.environ.get("MPN_BENCH_PRECISION", "auto"))
    ap.add_argument("--problems-per-gpu", type=int, default=PROBLEMS_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from mpinets_b200 import scenes, _lib
    from mpinets_b200.engine import Engine
    from mpinets_b200.parallel import gather_metrics, shard_range
    from oracle import oracle as O   # only for reference_state_dict (weights init) and the cpu_baseline leg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B, K, W = args.problems_per_gpu, args.steps, max(args.warmup, 3)
    eng = Engine(device=local)
    eng.load_state_dict(O.reference_state_dict(0))
    eng.reserve(B)
    precision = args.precision
    if precision == "auto":
        precision = "bf16"
        try:
            eng.encoder_forward(torch.zeros(1, 6272, 4, device="cuda"), _lib.PREC_BF16)
            torch.cuda.synchronize()
        except _lib.MpnError:
            precision = "fp32"
    prec = _lib.PREC_BF16 if precision == "bf16" else _lib.PREC_FP32

    # ---- problems: host (pinned) SoA, shard = rank's contiguous block of problem indices
    lo, hi = shard_range(rank, world, world * B)
    assert (lo, hi) == (rank * B, (rank + 1) * B)
    p = scenes.config_problems(2, B, problem0=lo)
    host = {k: torch.from_numpy(np.ascontiguousarray(p[k])).pin_memory() for k in scenes.SCENE_KEYS + ("q0", "target")}
    h2d_bytes = sum(t.numel() * t.element_size() for t in host.values())

    def upload():
        return {k: v.cuda(non_blocking=True) for k, v in host.items()}

    d = upload()
    sc = {k: d[k] for k in scenes.SCENE_KEYS}
    cloud = eng.build_cloud(sc, d["q0"], d["target"], problem0=rank * B)
    traj = torch.empty(B, K + 1, 7, device="cuda")
    metrics = torch.empty(B, _lib.METRICS_COLS, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up (W untimed steps), then exactly K timed steps
    eng.rollout(sc, cloud, d["q0"], d["target"], W, check_every_step=True, precision=prec)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    eng.profile(True)
    l0 = eng.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    eng.rollout(sc, cloud, d["q0"], d["target"], K, check_every_step=True, precision=prec, traj=traj, metrics=metrics)
    if world > 1:
        gathered = gather_metrics(metrics, world * B)    # the single collective: final metrics table (NCCL all-gather)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count - l0
    stages = eng.profile_read()
    eng.profile(False)
    clk = clocks.stop()
    t = torch.tensor([ms], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * K / (ms / 1000.0)

    # ---- e2e: host buffers -> H2D -> cloud build -> K steps -> D2H traj + metrics, all inside the timed region
    traj_h = torch.empty(B, K + 1, 7).pin_memory()
    metrics_h = torch.empty(B, _lib.METRICS_COLS).pin_memory()
    # the one-off cloud build (FK + 6272 rows per problem), timed on its own with CUDA events
    bms = []
    for _ in range(3):
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        cloud_tmp = eng.build_cloud(sc, d["q0"], d["target"], problem0=rank * B)
        b1.record()
        torch.cuda.synchronize()
        bms.append(b0.elapsed_time(b1))
    del cloud_tmp
    build_ms = min(bms)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    d2 = upload()
    sc2 = {k: d2[k] for k in scenes.SCENE_KEYS}
    cloud2 = eng.build_cloud(sc2, d2["q0"], d2["target"], problem0=rank * B)
    eng.rollout(sc2, cloud2, d2["q0"], d2["target"], K, check_every_step=True, precision=prec, traj=traj, metrics=metrics)
    traj_h.copy_(traj, non_blocking=True)
    metrics_h.copy_(metrics, non_blocking=True)
    e1.record()
    barrier()
    t2 = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_ms = float(t2.item())
    e2e_value = world * B * K / (e2e_ms / 1000.0)
    d2h_bytes = traj_h.numel() * 4 + metrics_h.numel() * 4

    if rank == 0:
        peaks = load_peaks()
        per_stage = {k: (v["ms"] / max(v["launches"], 1)) for k, v in stages.items() if v["launches"]}
        total_stage_ms = sum(v["ms"] for v in stages.values())
        dom = max((k for k in per_stage if k in FLOP_PER_STEP or k in BYTES_PER_STEP), key=lambda k: stages[k]["ms"])
        traffic = load_traffic()
        dom_traffic = traffic[dom]["dram_bytes_per_problem"] * B if dom in traffic else None
        if dom in FLOP_PER_STEP:
            ach = FLOP_PER_STEP[dom] * B / (per_stage[dom] / 1000.0) / 1e12
            peak = peaks["bf16_sustained"]
            roof = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": dom_traffic, "peak_source": peaks["source"] + ", sustained bf16",
                    "algorithmic_flop_per_launch": FLOP_PER_STEP[dom] * B, "avg_launch_ms": per_stage[dom],
                    "share_of_step": stages[dom]["ms"] / total_stage_ms}
        else:
            ach = BYTES_PER_STEP[dom] * B / (per_stage[dom] / 1000.0) / 1e9
            roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / peaks["hbm_gbs"], "traffic": dom_traffic, "peak_source": peaks["source"],
                    "algorithmic_bytes_per_launch": BYTES_PER_STEP[dom] * B, "avg_launch_ms": per_stage[dom],
                    "share_of_step": stages[dom]["ms"] / total_stage_ms}
        tensor_kernels = {}
        for k in ("sa1", "sa2", "sa3", "fc"):
            if k in per_stage:
                tf = FLOP_PER_STEP[k] * B / (per_stage[k] / 1000.0) / 1e12
                tensor_kernels[k] = {"TFLOPs": tf, "frac_of_bf16_sustained": tf / peaks["bf16_sustained"], "ms": per_stage[k]}
        hbm_kernels = {}
        for k in ("sample_robot", "sweep", "fps1"):
            if k in per_stage:
                g = BYTES_PER_STEP[k] * B / (per_stage[k] / 1000.0) / 1e9
                hbm_kernels[k] = {"GBps": g, "frac_of_hbm_peak": g / peaks["hbm_gbs"], "avg_launch_ms": per_stage[k]}
        g = (BYTES_PER_STEP["build_cloud"] + 3276) * B / (build_ms / 1000.0) / 1e9
        hbm_kernels["build_cloud"] = {"GBps": g, "frac_of_hbm_peak": g / peaks["hbm_gbs"], "avg_launch_ms": build_ms,
                                      "note": "one-off per problem (FK kernel + cloud kernel + output allocation), outside the timed regions"}
        line = {
            "metric": METRIC, "value": value, "unit": "env steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if precision == "bf16" else "fp32", "data": "synthetic",
            "config": {"workload": "configs[1]: 4096 tabletop PlanningProblems per GPU, 6272-pt clouds (2048 robot + 4096 obstacle + 128 target), "
                                   "lock-step policy rollout (FK + cloud resample + PointNet++ + delta-q + per-step SDF sweep)",
                       "problems_per_gpu": B, "global_problems": world * B, "parallelism": f"problem-sharded x{world}",
                       "precision": precision + (" (tcgen05, fp32 accumulate)" if precision == "bf16" else " (SIMT FMA parity mode)"),
                       "l2": "inputs larger than L2 (clouds 411 MB/GPU per step)", "weights": "random init, seed 0"},
            "gpu_launches": launches, "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "env steps/s", "h2d_bytes_per_step": h2d_bytes / K, "d2h_bytes_per_step": d2h_bytes / K,
                    "ms_total": e2e_ms},
            "roofline": roof,
            "stage_ms_per_step": {k: v["ms"] / K for k, v in stages.items() if v["launches"]},
            "hbm_kernels": hbm_kernels, "tensor_kernels": tensor_kernels,
            "tensor_flops_per_step": FLOP_TOTAL * B,
            "achieved_tflops_whole_step": FLOP_TOTAL * B * K * world / (ms / 1000.0) / 1e12,
            "collision_rate": float(metrics[:, 0].mean().item()),
        }
        if not args.no_cpu_baseline and world == 1:   # rank 0 at N = 1 only: the other ranks must not wait on 10 s of CPU work
            line["cpu_baseline"] = cpu_baseline(eng.cfg.seed)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
