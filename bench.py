#!/usr/bin/env python
"""bench.py -- env steps/sec of the batched MPiNets rollout hot path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload 2|3|4|5] [--precision bf16x3|bf16|fp32] [--impl reference]

One "step" = one lock-step policy step over the whole per-GPU batch of PlanningProblems: PointNet++ encoder (FPS,
ball query, grouping, shared MLPs) + delta-q head + clamp/unnormalise + FK + robot-surface resample into the cloud +
link-sphere SDF collision check.

Workloads (numbering of SURVEY.md section 8d = scenes.config_problems):
  2  BASELINE configs[1]: 4096 tabletop problems per GPU                               (default at N = 1)
  3  BASELINE configs[2]: 4096 cubby(+merged) + dresser problems, collision check every step, T = 70
  4  BASELINE configs[3]: 4096 per GPU of the mixed thirds (tabletop / cubby / dresser), T = 70, one NCCL gather
                                                                                        (default at N > 1: 32768 problems on 8)
  5  BASELINE configs[4]: training step, 8192 samples per GPU, DDP all-reduce           (samples/s; --steps training steps)

Printed JSON (rank 0).  `value` / `dtype` describe the PARITY-GRADE mode (bf16x3: split-bf16 operands on tcgen05, delta-q within
1e-5 of the fp32 reference); `fast_mode` is the same job in the bf16 throughput mode with its delta-q error.  value = problems*K /
device time of K steps with everything resident in HBM; e2e = the same job through the public API with HOST problem buffers
(pinned H2D of the problem SoA, cloud build, K steps, D2H of trajectories + metrics inside the timed region); parity = match
fields of the metric (collision flags, FPS indices, delta-q) against the CPU oracle, computed outside the timed regions;
roofline = dominant kernel vs MEASURED_PEAKS.json; cpu_baseline = the CPU oracle port on this box's host cores.
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
# bf16 MMA flops the bf16x3 mode actually issues per algorithmic flop (three products per MAC; SA2's first layer is evaluated per
# point instead of per pair, so its 67-wide layer costs 1/32 of the pairwise form)
X3_MMA_FACTOR = {"sa1": 3.0, "sa2": 3.0 * (128 * 128 + 128 * 256) / (67 * 128 + 128 * 128 + 128 * 256), "sa3": 3.0, "fc": 3.0}
# The fused SA kernels push only the DISTINCT neighbour rows of a group through the shared MLP, packed into 128-row MMA tiles
# (mpn_sa_tile_counts): groups per problem, and the bf16 MMA flops ONE tile issues in each mode (count of tcgen05.mma x 2*M*N*K):
SA_GROUPS = {"sa1": 512, "sa2": 128}
MMA_FLOP_PER_TILE = {
    "bf16": {"sa1": 12 * 2 * 128 * 64 * 16,        # 3 x (bias step + K steps): 2 + 5 + 5
             "sa2": 29 * 2 * 128 * 128 * 16},      # 5 + 8 + 2 x 8
    "bf16x3": {"sa1": 27 * 2 * 128 * 64 * 16,      # 1 + 2 x (bias step + 3 x 4)
               "sa2": 72 * 2 * 128 * 128 * 16},    # 3 x 8 + 2 x 3 x 8; layer 1 = the per-point GEMM below
}
SA2_PRE_MMA_FLOP = 3 * 2 * 512 * 80 * 128          # bf16x3: per-point layer-1 GEMM of SA2, per problem
BYTES_PER_STEP = {"sample_robot": 2048 * 16 + 11 * 48, "sweep": 28 + 3200 + 1, "build_cloud": 6272 * 16,
                  "fps1": 6272 * 16 + 512 * 16, "fps2": 512 * 12 + 128 * 16}
WORKLOADS = {
    2: dict(baseline="configs[1]", rollout_T=50, desc="4096 tabletop PlanningProblems per GPU"),
    3: dict(baseline="configs[2]", rollout_T=70, desc="4096 cubby(+merged) + dresser PlanningProblems per GPU, collision check every step"),
    4: dict(baseline="configs[3]", rollout_T=70, desc="4096 mixed-scene PlanningProblems per GPU (tabletop / cubby / dresser thirds), "
                                                      "problem-sharded, one NCCL gather of the metrics table"),
}


def load_traffic():
    """per-problem DRAM bytes (read + write) of each stage's dominant kernel from the committed ncu --set full captures"""
    out = {}
    for name in ("r1_traffic.json", "r2_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            out.update(json.load(open(p)))
    return out


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
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
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


def _host_threads():
    import torch
    if torch.get_num_threads() == 1 and (os.cpu_count() or 1) > 2:   # torchrun pins OMP_NUM_THREADS=1: undo it for the CPU arm
        torch.set_num_threads(max(1, (os.cpu_count() or 2) // 2))      # (physical cores; SMT siblings only slow the GEMMs down)
    return torch.get_num_threads()


class CpuPort:
    """The CPU port of the same step (oracle/): torch-CPU MLPs on the host threads + C geometry.  Problems, weights and the
    t = 0 clouds are prepared once (untimed, like the GPU arm's resident inputs); run() times rollout + collision sweep."""

    def __init__(self, n_problems: int, seed: int, config: int = 2):
        from mpinets_b200 import scenes, franka
        from oracle import oracle as O
        self.threads = _host_threads()
        self.O, self.n, self.seed = O, n_problems, seed
        self.tables = franka.default_tables()
        self.sd = O.reference_state_dict(0)
        self.p = scenes.config_problems(config, n_problems)
        self.cloud = O.build_cloud(self.p["q0"], self.p["target"], self.p, self.tables, seed)
        self.qn = O.normalize(self.p["q0"], self.tables.joint_limits)

    def run(self, steps: int) -> float:
        t0 = time.perf_counter()
        traj = self.O.rollout(self.sd, self.cloud.copy(), self.qn, self.tables, steps, self.seed, emulate_bf16=False)
        self.O.sweep_flags(self.p, traj, self.tables)
        return time.perf_counter() - t0


def cpu_config0(reps: int = 25):
    """BASELINE configs[0] exactly as SURVEY section 8d words it: ONE tabletop problem -- build the 4096-point obstacle cloud, then
    the link-sphere SDF sweep of 50 poses (joint-space interpolation between two in-limit configurations), no network.
    Oracle port at 1 thread and on all host cores (independent problems over threads; the C oracle releases the GIL), median of
    `reps` repetitions after a warm-up."""
    from concurrent.futures import ThreadPoolExecutor
    from mpinets_b200 import scenes, franka
    from oracle import oracle as O
    tables = franka.default_tables()
    cores = os.cpu_count() or 1
    PER = 16                                    # problems per thread and call on the all-cores leg (amortises the Python call overhead)
    P = scenes.config_problems(2, cores * PER)
    arr = {k: v for k, v in P.items() if isinstance(v, np.ndarray)}
    w = np.linspace(0.0, 1.0, 50, dtype=np.float32)[None, :, None]
    traj = (P["q0"][:, None, :] * (1 - w) + P["q_goal"][:, None, :] * w).astype(np.float32)

    def job(lo, hi):
        p = {k: v[lo:hi] for k, v in arr.items()}
        t0 = time.perf_counter()
        O.sample_obstacles(p, 4096, 0x4D50694E, problem0=lo)
        t1 = time.perf_counter()
        O.sweep_flags(p, traj[lo:hi], tables)
        return t1 - t0, time.perf_counter() - t1

    job(0, 1)
    single = np.array([job(0, 1) for _ in range(reps)])
    b1, s1 = float(np.median(single[:, 0])), float(np.median(single[:, 1]))
    walls = []
    with ThreadPoolExecutor(cores) as ex:
        run_all = lambda: list(ex.map(lambda i: job(i * PER, (i + 1) * PER), range(cores)))
        run_all()
        for _ in range(reps):
            t0 = time.perf_counter()
            run_all()
            walls.append(time.perf_counter() - t0)
    wall = float(np.median(walls))
    n_prims = int((np.abs(P["cuboid_dims"][0]).min(axis=1) > 1e-8).sum() + ((np.abs(P["cylinder_radii"][0]).reshape(-1) > 1e-8) &
                                                                              (np.abs(P["cylinder_heights"][0]).reshape(-1) > 1e-8)).sum())
    return {"workload": "configs[0]: 1 tabletop problem, 4096-pt obstacle cloud build + 50-pose x %d-sphere x %d-primitive SDF sweep, no network"
                        % (tables.sphere_centers.shape[0], n_prims),
            "reps": reps, "kind": "port",
            "threads_1": {"build_cloud_ms": 1e3 * b1, "sweep_50_poses_ms": 1e3 * s1, "problems_per_s": 1.0 / (b1 + s1),
                          "pose_checks_per_s": 50.0 / s1},
            "all_cores": {"cores": cores, "problems_per_call_and_thread": PER, "wall_ms": 1e3 * wall, "problems_per_s": cores * PER / wall,
                          "pose_checks_per_s": 50.0 * cores * PER / wall},
            "reference_classes_note": "the real TorchCuboids/TorchCylinders.sdf_sequence (geometry.py:290-347,509-568) need /root/reference, "
                                      "which does not exist on the GPU box: timed by scripts/cpu_baseline_config0.py in the build container "
                                      "(profiles/r2_cpu_baseline_config0.json)"}


def cpu_baseline(seed: int, config: int, budget_s: float = 10.0, n_problems: int = 16, steps: int = 4, max_calls: int = 8):
    """bounded sample of the bench workload on the host cores: calls of (n_problems x steps) until ~budget_s of CPU work"""
    port = CpuPort(n_problems, seed, config)
    port.run(1)   # warm-up (thread pools, page faults)
    total, calls = 0.0, 0
    while calls < max_calls and (calls == 0 or total < budget_s):
        total += port.run(steps)
        calls += 1
    done = calls * n_problems * steps
    out = {"value": done / total, "unit": "env steps/s", "cores": port.threads, "kind": "port",
           "sample": f"{calls} x ({n_problems} problems x {steps} steps) of the same workload = {done} problem-steps in {total:.1f} s "
                     f"(cloud build excluded), torch-CPU fp32 MLPs + C oracle geometry"}
    try:
        out["config0"] = cpu_config0()
    except Exception as e:   # the baseline leg must not take the bench line down
        out["config0"] = {"error": repr(e)}
    return out


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path.  The reference cannot be imported
    (pointnet2_ops is CUDA-only; robofin/geometrout are not installable offline), so this times the oracle port on all host
    threads.  One bench step = one lock-step env step of a bounded sample of the workload (16 problems)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload or (2 if args.gpus == 1 else 4)
    if wl == 5:
        print(json.dumps({"impl": "reference", "unavailable": "the training step has no CPU reference arm (torch.autograd oracle is a test checker only)"}))
        return
    n = 16
    port = CpuPort(n, 1, wl)
    for _ in range(max(args.warmup, 1)):
        port.run(1)
    total = 0.0
    for _ in range(args.steps):
        total += port.run(1)
    value = n * args.steps / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "env steps/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": f"{WORKLOADS[wl]['baseline']}: {WORKLOADS[wl]['desc']}; 6272-pt clouds, lock-step policy rollout + SDF sweep",
                       "sample": f"{n} problems x 1 step per bench step"},
            "cpu_baseline": {"value": value, "unit": "env steps/s", "cores": port.threads, "kind": "port",
                             "sample": f"{n} problems x {args.steps} steps (cloud build excluded), torch-CPU MLPs + C oracle geometry"},
            "e2e": {"value": value, "unit": "env steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


_REAL_STDOUT = None


def emit(line: dict):
    """the ONE JSON line of the run, on the process's original stdout"""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def setup_dist():
    """torch.distributed over NCCL when launched by torchrun.  Everything libraries print on fd 1 during the run (NCCL's version banner
    and its NCCL_DEBUG lines) is sent to stderr, so that stdout stays the single JSON line and the communicator lines remain checkable."""
    global _REAL_STDOUT
    if not os.environ.get("NCCL_DEBUG_FILE"):
        os.environ["NCCL_DEBUG_FILE"] = "/dev/stderr"
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback for the product path)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        t = torch.ones(1, device="cuda")
        dist.all_reduce(t)   # communicator creation outside every timed region
        assert int(t.item()) == world
        print(f"[bench] rank {rank}/{world} on cuda:{local}: NCCL {'.'.join(map(str, torch.cuda.nccl.version()))} communicator up, "
              f"nranks={world}", file=sys.stderr, flush=True)
    return world, rank, local


def sa_work(k, B, mode, tiles_per_launch):
    """executed work of one SA launch: rows really pushed through the shared MLP (tiles x 128) vs the reference's 128 rows per group"""
    ratio = tiles_per_launch / float(SA_GROUPS[k] * B)            # 128-row tiles per group (1.0 = the reference formulation)
    issued = tiles_per_launch * MMA_FLOP_PER_TILE[mode][k] + (SA2_PRE_MMA_FLOP * B if (k == "sa2" and mode == "bf16x3") else 0)
    return {"tiles_per_group": ratio, "executed_flop": FLOP_PER_STEP[k] * B * ratio, "issued_mma_flop": issued}


def roofline_of(stages, per_stage, B, peaks, mode, traffic, tiles=None):
    """tiles: {"sa1": tiles per launch, "sa2": ...} from mpn_sa_tile_counts (None for the fp32 mode)"""
    total_stage_ms = sum(v["ms"] for v in stages.values())
    dom = max((k for k in per_stage if k in FLOP_PER_STEP or k in BYTES_PER_STEP), key=lambda k: stages[k]["ms"])
    tkey = dom + ("_x3" if mode == "bf16x3" else "")
    dom_traffic = traffic[tkey]["dram_bytes_per_problem"] * B if tkey in traffic else None
    if dom in FLOP_PER_STEP:
        sec = per_stage[dom] / 1000.0
        ref_flop = FLOP_PER_STEP[dom] * B
        peak = peaks["bf16_sustained"]
        if tiles and dom in SA_GROUPS and mode in MMA_FLOP_PER_TILE:
            w = sa_work(dom, B, mode, tiles[dom])
            flop, issued = w["executed_flop"], w["issued_mma_flop"]
        else:
            flop, issued = ref_flop, ref_flop * (X3_MMA_FACTOR.get(dom, 3.0) if mode == "bf16x3" else 1.0)
        ach = ref_flop / sec / 1e12
        roof = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                "traffic": dom_traffic, "peak_source": peaks["source"] + ", sustained bf16",
                "algorithmic_flop_per_launch": ref_flop, "avg_launch_ms": per_stage[dom],
                "share_of_step": stages[dom]["ms"] / total_stage_ms,
                "executed_flop_per_launch": flop, "executed_tflops": flop / sec / 1e12, "executed_frac_of_peak": flop / sec / 1e12 / peak,
                "issued_mma_flop_per_launch": issued, "issued_mma_frac_of_peak": issued / sec / 1e12 / peak}
        note = ["achieved = SURVEY 8(d)'s algorithmic flops of the stage (the reference formulation: all 128 rows of every ball-query group) "
                "x problems per launch / the launch time measured with CUDA events"]
        if tiles and dom in SA_GROUPS:
            roof["tiles_per_group"] = tiles[dom] / float(SA_GROUPS[dom] * B)
            note.append("the kernel does not execute all of that work: a group holds H <= 128 distinct neighbours (the rest are copies of the "
                        "first hit, which the max-pool ignores) and the distinct rows of several groups are packed into 128-row MMA tiles "
                        "(%.3f tiles per group, counted on the device) -- executed_* = the fp32 formulation of the rows really evaluated, "
                        "issued_mma_* = the bf16 MMA flops behind them, i.e. the tensor pipe's real load (its utilisation is "
                        "issued_mma_frac_of_peak, not frac)" % roof["tiles_per_group"])
        if mode == "bf16x3":
            note.append("the bf16x3 mode issues three bf16 MMAs per product (a_hi w_hi + a_lo w_hi + a_hi w_lo)")
        roof["note"] = "; ".join(note)
    else:
        ach = BYTES_PER_STEP[dom] * B / (per_stage[dom] / 1000.0) / 1e9
        roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "traffic": dom_traffic, "peak_source": peaks["source"],
                "algorithmic_bytes_per_launch": BYTES_PER_STEP[dom] * B, "avg_launch_ms": per_stage[dom],
                "share_of_step": stages[dom]["ms"] / total_stage_ms}
        if dom.startswith("fps"):
            roof["note"] = ("latency-bound kernel: %d dependent selection rounds per problem; the HBM figure only says that its traffic "
                            "(the cloud once) is not what limits it" % (511 if dom == "fps1" else 127))
    return roof


def tensor_kernels_of(stages, tiles, B, mode, peaks):
    """per tensor stage: ms per launch, algorithmic TFLOP/s (SURVEY 8(d)'s figure: all 128 rows per group), executed TFLOP/s (the fp32
    formulation of the distinct rows really evaluated) and the issued bf16 MMA rate (the tensor pipe's real load)"""
    out = {}
    for k in ("sa1", "sa2", "sa3", "fc"):
        v = stages.get(k)
        if not v or not v["launches"]:
            continue
        ms = v["ms"] / v["launches"]
        sec = ms / 1000.0
        ref = FLOP_PER_STEP[k] * B
        if tiles and k in SA_GROUPS and mode in MMA_FLOP_PER_TILE:
            w = sa_work(k, B, mode, tiles[k])
            flop, issued = w["executed_flop"], w["issued_mma_flop"]
        else:
            flop, issued = ref, ref * (X3_MMA_FACTOR[k] if mode == "bf16x3" else 1.0)
        out[k] = {"ms": ms, "TFLOPs": ref / sec / 1e12, "frac_of_bf16_sustained": ref / sec / 1e12 / peaks["bf16_sustained"],
                  "executed_tflops": flop / sec / 1e12, "executed_frac_of_bf16_sustained": flop / sec / 1e12 / peaks["bf16_sustained"],
                  "issued_mma_frac_of_bf16_sustained": issued / sec / 1e12 / peaks["bf16_sustained"]}
        if tiles and k in SA_GROUPS:
            out[k]["tiles_per_group"] = tiles[k] / float(SA_GROUPS[k] * B)
    return out


def run_training(args, world, rank, local):
    """--workload 5 (BASELINE configs[4]): one training step = forward with saved state + both losses + backward to all 19.07 M
    parameters (mpn_train_step_grads) -> NCCL all-reduce-mean of the flat gradient vector -> clip_grad_norm_(1.0) + Adam."""
    import torch
    import torch.distributed as dist
    from mpinets_b200 import scenes, _lib
    from mpinets_b200.engine import Engine
    from mpinets_b200.parallel import allreduce_mean_, broadcast_params_
    from oracle import oracle as O   # weights init only
    B = args.samples_per_gpu
    K, W = args.steps, max(args.warmup, 3)
    precision = "bf16" if args.precision in ("auto", "bf16x3") else args.precision
    prec = _lib.PRECISIONS[precision]
    eng = Engine(device=local)
    eng.load_state_dict(O.reference_state_dict(0))
    broadcast_params_(eng)
    p = scenes.config_problems(4, B, problem0=rank * B)
    host = {k: torch.from_numpy(np.ascontiguousarray(p[k])).pin_memory() for k in scenes.SCENE_KEYS + ("q0", "target")}
    d = {k: v.cuda(non_blocking=True) for k, v in host.items()}
    sc = {k: d[k] for k in scenes.SCENE_KEYS}
    cloud = eng.build_cloud(sc, d["q0"], d["target"], problem0=rank * B)
    qn = eng.normalize(d["q0"])
    gen = torch.Generator(device="cuda").manual_seed(rank)
    sup = torch.clamp(qn + 0.05 * torch.randn(qn.shape, generator=gen, device="cuda"), -1, 1)
    grads = torch.empty(eng.param_count, device="cuda")
    ev = lambda: torch.cuda.Event(enable_timing=True)

    def step(i, rec=None):
        e = [ev() for _ in range(4)]
        e[0].record()
        losses, _, _ = eng.train_step_grads(sc, cloud, qn, sup, grads=grads, precision=prec)
        e[1].record()
        allreduce_mean_(grads)
        e[2].record()
        eng.adam_step(grads, i + 1)
        e[3].record()
        if rec is not None:
            rec.append(e)
        return losses

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(W):
        step(i)
    barrier()
    clocks = ClockSampler(local)
    clocks.start()
    rec, l0 = [], eng.launch_count
    e0, e1 = ev(), ev()
    barrier()
    e0.record()
    for i in range(K):
        losses = step(W + i, rec)
    e1.record()
    barrier()
    clk = clocks.stop()
    launches = eng.launch_count - l0
    ph = np.array([[e[j].elapsed_time(e[j + 1]) for j in range(3)] for e in rec]).mean(axis=0)
    # e2e: host-resident batch -> H2D -> cloud build (the data loader's per-item CPU work in the reference) -> step -> D2H of the losses
    loss_h = torch.empty(2).pin_memory()
    barrier()
    x0, x1 = ev(), ev()
    x0.record()
    for i in range(K):
        d2 = {k: v.cuda(non_blocking=True) for k, v in host.items()}
        sc2 = {k: d2[k] for k in scenes.SCENE_KEYS}
        cloud2 = eng.build_cloud(sc2, d2["q0"], d2["target"], problem0=rank * B)
        qn2 = eng.normalize(d2["q0"])
        l2, _, _ = eng.train_step_grads(sc2, cloud2, qn2, sup, grads=grads, precision=prec)
        allreduce_mean_(grads)
        eng.adam_step(grads, W + K + i + 1)
        loss_h.copy_(l2, non_blocking=True)
    x1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1), x0.elapsed_time(x1), ph[0], ph[1], ph[2]], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t = t.cpu().numpy()
    if rank == 0:
        peaks = load_peaks()
        flop = 3 * FLOP_TOTAL * B   # forward + data gradients + weight gradients of the same contractions
        h2d = sum(v.numel() * v.element_size() for v in host.values())
        line = {"metric": "training samples/sec (PointNetEncoder fwd+bwd + CollisionLoss + BC loss, DDP all-reduce, clip + Adam)",
                "value": world * B * K / (t[0] / 1000.0), "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": float(t[0] / K), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": precision,
                "data": "synthetic", "gpu_launches": launches, "clocks": clk,
                "config": {"workload": "configs[4]: training step, %d samples per GPU (mixed scenes), encoder + heads fwd+bwd, "
                                       "CollisionAndBCLossContainer (1024 robot points, margin 0.03, weights 5/1), DDP all-reduce of 19.07 M "
                                       "fp32 gradients, clip 1.0 + Adam 1e-4" % B,
                           "samples_per_gpu": B, "global_batch": world * B, "parallelism": f"dp{world}",
                           "precision": precision + (" operands on tcgen05, fp32 master weights / accumulation / optimizer state" if precision == "bf16" else ""),
                           "l2": "inputs larger than L2 (clouds %d MB/GPU)" % (B * 6272 * 16 // 2 ** 20), "weights": "random init, seed 0"},
                "e2e": {"value": world * B * K / (t[1] / 1000.0), "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8,
                        "ms_total": float(t[1]), "note": "per step: pinned H2D of the problem SoA, GPU cloud build, training step, D2H of the losses"},
                "phases_ms": {"forward_backward": float(t[2]), "grad_allreduce": float(t[3]), "clip_adam_repack": float(t[4])},
                "roofline": {"kernel": "whole training step", "bound": "tensor", "achieved": flop / (t[0] / K / 1000.0) / 1e12,
                             "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": flop / (t[0] / K / 1000.0) / 1e12 / peaks["bf16_sustained"],
                             "traffic": None, "peak_source": peaks["source"] + ", sustained bf16",
                             "algorithmic_flop_per_launch": flop,
                             "note": "reference formulation: 3 x the forward's contraction flops per sample (forward + data gradients + weight "
                                     "gradients over all 128 rows of every group).  NOT a hardware rate: the forward evaluates only the distinct rows "
                                     "of a group (packed tiles, ~0.3-0.6 tiles per group) and the backward only the rows that won a channel of the "
                                     "max-pool (SA1 ~7 of 128, SA2 ~25 of 128), compacted over the chunk; the executed tensor work is a small "
                                     "fraction of this figure and the step is bound by HBM-side row traffic and launch count, not by the tensor pipe"},
                "losses": [float(x) for x in losses.cpu()], "params": 19068103}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", type=int, default=0, choices=[0, 2, 3, 4, 5])
    ap.add_argument("--precision", default=os.environ.get("MPN_BENCH_PRECISION", "auto"), choices=["auto", "bf16x3", "bf16", "fp32"])
    ap.add_argument("--problems-per-gpu", type=int, default=PROBLEMS_PER_GPU)
    ap.add_argument("--samples-per-gpu", type=int, default=8192)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the fast-mode / extra-workload / parity legs (profiling runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    world, rank, local = setup_dist()
    if args.workload == 5:
        return run_training(args, world, rank, local)
    import torch
    import torch.distributed as dist
    from mpinets_b200 import scenes, _lib
    from mpinets_b200.engine import Engine
    from mpinets_b200.parallel import gather_metrics, shard_range
    from oracle import oracle as O   # reference_state_dict (weights init), the parity checker and the cpu_baseline leg -- never timed as product

    wl = args.workload or (2 if world == 1 else 4)
    B, K, W = args.problems_per_gpu, args.steps, max(args.warmup, 3)
    eng = Engine(device=local)
    sd = O.reference_state_dict(0)
    eng.load_state_dict(sd)
    eng.reserve(B)
    precision = "bf16x3" if args.precision == "auto" else args.precision
    prec = _lib.PRECISIONS[precision]

    # ---- problems: host (pinned) SoA, shard = rank's contiguous block of problem indices
    lo, hi = shard_range(rank, world, world * B)
    assert (lo, hi) == (rank * B, (rank + 1) * B)

    def make_inputs(config):
        p = scenes.config_problems(config, B, problem0=lo)
        host = {k: torch.from_numpy(np.ascontiguousarray(p[k])).pin_memory() for k in scenes.SCENE_KEYS + ("q0", "target")}
        return p, host

    p, host = make_inputs(wl)
    h2d_bytes = sum(t.numel() * t.element_size() for t in host.values())

    def upload(h):
        return {k: v.cuda(non_blocking=True) for k, v in h.items()}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        t = torch.tensor([x], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_job(h, prec_, steps, warm, profile=False, sample_clocks=False):
        """W untimed warm-up steps, then exactly `steps` timed lock-step steps on resident inputs (+ the metrics gather)"""
        d = upload(h)
        sc = {k: d[k] for k in scenes.SCENE_KEYS}
        cloud = eng.build_cloud(sc, d["q0"], d["target"], problem0=rank * B)
        traj = torch.empty(B, steps + 1, 7, device="cuda")
        metrics = torch.empty(B, _lib.METRICS_COLS, device="cuda")
        if warm:
            eng.rollout(sc, cloud, d["q0"], d["target"], warm, check_every_step=True, precision=prec_)
            cloud = eng.build_cloud(sc, d["q0"], d["target"], problem0=rank * B)   # the timed rollout starts from t = 0 again
        barrier()
        clocks = ClockSampler(local) if sample_clocks else None
        if clocks:
            clocks.start()
        if profile:
            eng.profile(True)
            eng.sa_tile_counts(reset=True)
        l0 = eng.launch_count
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        eng.rollout(sc, cloud, d["q0"], d["target"], steps, check_every_step=True, precision=prec_, traj=traj, metrics=metrics)
        gathered = gather_metrics(metrics, world * B) if world > 1 else metrics   # the single collective: final metrics table
        ev1.record()
        barrier()
        out = {"ms": maxr(ev0.elapsed_time(ev1)), "launches": eng.launch_count - l0, "traj": traj, "metrics": metrics,
               "gathered": gathered, "d": d, "sc": sc}
        if profile:
            out["stages"] = eng.profile_read()
            eng.profile(False)
            t1, t2 = eng.sa_tile_counts(reset=True)
            n1, n2 = out["stages"]["sa1"]["launches"], out["stages"]["sa2"]["launches"]
            out["tiles"] = {"sa1": t1 / max(n1, 1), "sa2": t2 / max(n2, 1)} if prec_ != _lib.PREC_FP32 else None
        if clocks:
            out["clocks"] = clocks.stop()
        return out

    main_run = timed_job(host, prec, K, W, profile=True, sample_clocks=True)
    ms, stages, clk = main_run["ms"], main_run["stages"], main_run["clocks"]
    value = world * B * K / (ms / 1000.0)
    d, sc, traj, metrics = main_run["d"], main_run["sc"], main_run["traj"], main_run["metrics"]

    # ---- e2e: host buffers -> H2D -> cloud build -> K steps -> D2H traj + metrics, all inside the timed region
    traj_h = torch.empty(B, K + 1, 7).pin_memory()
    metrics_h = torch.empty(B, _lib.METRICS_COLS).pin_memory()
    bms = []
    for _ in range(3):   # the one-off cloud build (FK + 6272 rows per problem), timed on its own with CUDA events
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
    d2 = upload(host)
    sc2 = {k: d2[k] for k in scenes.SCENE_KEYS}
    cloud2 = eng.build_cloud(sc2, d2["q0"], d2["target"], problem0=rank * B)
    eng.rollout(sc2, cloud2, d2["q0"], d2["target"], K, check_every_step=True, precision=prec, traj=traj, metrics=metrics)
    if world > 1:
        gather_metrics(metrics, world * B)
    traj_h.copy_(traj, non_blocking=True)
    metrics_h.copy_(metrics, non_blocking=True)
    e1.record()
    barrier()
    e2e_ms = maxr(e0.elapsed_time(e1))
    e2e_value = world * B * K / (e2e_ms / 1000.0)
    d2h_bytes = traj_h.numel() * 4 + metrics_h.numel() * 4
    del d2, sc2, cloud2

    # ---- the same job in the bf16 throughput mode (every rank runs it: same barriers)
    fast = None
    if precision != "bf16" and not args.no_extra:
        fr = timed_job(host, _lib.PREC_BF16, K, W, profile=True)
        fast = {"dtype": "bf16", "value": world * B * K / (fr["ms"] / 1000.0), "unit": "env steps/s", "ms_per_step": fr["ms"] / K,
                "gpu_launches": fr["launches"], "stage_ms_per_step": {k: v["ms"] / K for k, v in fr["stages"].items() if v["launches"]},
                "tensor_kernels": tensor_kernels_of(fr["stages"], fr["tiles"], B, "bf16", load_peaks()),
                "collision_rate": float(fr["metrics"][:, 0].mean().item()), "_run": fr}

    # ---- configs[2] next to the N = 1 default: cubby + dresser, the full T = 70 rollout with the per-step check
    extra = {}
    if world == 1 and wl == 2 and not args.no_extra:
        p3, host3 = make_inputs(3)
        r3 = timed_job(host3, prec, 70, 1, profile=False)
        extra["configs[2]"] = {"workload": WORKLOADS[3]["desc"] + ", full 70-step rollout", "dtype": precision, "steps": 70,
                               "value": B * 70 / (r3["ms"] / 1000.0), "unit": "env steps/s", "ms_per_step": r3["ms"] / 70,
                               "collision_rate": float(r3["metrics"][:, 0].mean().item()),
                               "collision_flag_match": float((O.sweep_flags(p3, r3["traj"].cpu().numpy(), eng.tables)[0] ==
                                                              r3["metrics"][:, 0].cpu().numpy().astype(np.uint8)).mean())}
        del r3, host3
        # the N > 1 default is configs[3] (mixed scenes, denser SA2 neighbourhoods than tabletop): one GPU's shard of it, so that the
        # scaling efficiency of the N-GPU lines can be read against the SAME workload
        p4, host4 = make_inputs(4)
        r4 = timed_job(host4, prec, K, W, profile=False)
        extra["configs[3] shard"] = {"workload": WORKLOADS[4]["desc"] + " -- one GPU's 4096-problem shard (what every rank of the N > 1 default runs)",
                                     "dtype": precision, "steps": K, "value": B * K / (r4["ms"] / 1000.0), "unit": "env steps/s",
                                     "ms_per_step": r4["ms"] / K, "collision_rate": float(r4["metrics"][:, 0].mean().item())}
        del r4, host4

    # ---- match fields of the metric, against the CPU oracle, outside every timed region (rank 0's shard)
    parity = None
    if rank == 0 and not args.no_extra:
        _host_threads()
        th = traj.cpu().numpy()
        gflags = metrics[:, 0].cpu().numpy().astype(np.uint8)
        oflags, ofirst, _ = O.sweep_flags(p, th, eng.tables)                      # all B problems: GPU sweep vs CPU sweep of the same trajectories
        n_sub = 64
        sub = np.arange(0, B, max(1, B // n_sub))[:n_sub]
        subt = torch.from_numpy(sub).cuda()
        psub = {k: v[sub] for k, v in p.items() if isinstance(v, np.ndarray)}
        cloud0 = eng.build_cloud(sc, d["q0"], d["target"], problem0=rank * B)      # the t = 0 clouds (the rollout rewrote the robot rows)
        c_sub = cloud0[subt].contiguous()
        oc = np.concatenate([O.build_cloud(psub["q0"][j:j + 1], psub["target"][j:j + 1], {k: v[j:j + 1] for k, v in psub.items()}, eng.tables,
                                           eng.cfg.seed, problem0=rank * B + int(i)) for j, i in enumerate(sub)])
        cloud_match = float(np.mean([np.array_equal(c_sub[j].cpu().numpy(), oc[j]) for j in range(len(sub))]))
        fidx = eng.fps(c_sub, 512).cpu().numpy()
        ofidx = O.fps(oc, 512)
        qn_sub = O.normalize(psub["q0"], eng.tables.joint_limits)
        odq = O.policy_forward(sd, oc, qn_sub).numpy()
        qn_dev = torch.from_numpy(qn_sub).cuda()
        dq_main = eng.policy_forward(c_sub, qn_dev, prec).cpu().numpy()
        # the mode's rollout against the fp32 SIMT parity mode of the library on a 256-problem subset, K steps
        n256 = min(256, B)
        s256 = torch.arange(0, B, max(1, B // n256), device="cuda")[:n256]
        sc256 = {k: sc[k][s256].contiguous() for k in scenes.SCENE_KEYS}
        q256, t256 = d["q0"][s256].contiguous(), d["target"][s256].contiguous()
        c256 = cloud0[s256].contiguous()
        del cloud0
        _, m32 = eng.rollout(sc256, c256, q256, t256, K, check_every_step=True, precision=_lib.PREC_FP32)
        f32flags = m32[:, 0].cpu().numpy().astype(np.uint8)
        parity = {"oracle": "oracle/ (CPU restatement; C geometry + torch-CPU fp32 network), outside the timed regions",
                  "collision_flag_match": float((gflags == oflags).mean()), "first_collision_step_match": float((metrics[:, 1].cpu().numpy().astype(np.int32) == ofirst).mean()),
                  "collision_flag_problems": int(B),
                  "cloud_match": cloud_match, "fps_idx_match": float(np.mean([np.array_equal(fidx[j], ofidx[j]) for j in range(len(sub))])),
                  "dq_max_abs_err": float(np.abs(dq_main - odq).max()), "dq_tolerance": 1e-5, "subset_problems": int(len(sub)),
                  "flags_equal_fp32_mode_rollout": float((gflags[s256.cpu().numpy()] == f32flags).mean()), "fp32_mode_subset_problems": int(n256),
                  "rollout_steps": K}
        if fast is not None:
            dq_fast = eng.policy_forward(c_sub, qn_dev, _lib.PREC_BF16).cpu().numpy()
            fast["dq_max_abs_err"] = float(np.abs(dq_fast - odq).max())
            ff = fast["_run"]["metrics"][:, 0].cpu().numpy().astype(np.uint8)
            fast["flags_equal_fp32_mode_rollout"] = float((ff[s256.cpu().numpy()] == f32flags).mean())
            fast["flags_equal_parity_mode"] = float((ff == gflags).mean())
    if fast is not None:
        fast.pop("_run", None)

    if rank == 0:
        peaks = load_peaks()
        traffic = load_traffic()
        per_stage = {k: (v["ms"] / max(v["launches"], 1)) for k, v in stages.items() if v["launches"]}
        roof = roofline_of(stages, per_stage, B, peaks, precision, traffic, main_run.get("tiles"))
        tensor_kernels = tensor_kernels_of(stages, main_run.get("tiles"), B, precision, peaks)
        hbm_kernels = {}
        for k in ("sample_robot", "sweep", "fps1"):
            if k in per_stage:
                g = BYTES_PER_STEP[k] * B / (per_stage[k] / 1000.0) / 1e9
                hbm_kernels[k] = {"GBps": g, "frac_of_hbm_peak": g / peaks["hbm_gbs"], "avg_launch_ms": per_stage[k]}
        g = (BYTES_PER_STEP["build_cloud"] + 3276) * B / (build_ms / 1000.0) / 1e9
        hbm_kernels["build_cloud"] = {"GBps": g, "frac_of_hbm_peak": g / peaks["hbm_gbs"], "avg_launch_ms": build_ms,
                                      "note": "one-off per problem (FK kernel + cloud kernel + output allocation), outside the timed regions"}
        prec_desc = {"bf16x3": "bf16x3: split-bf16 operands (hi + lo, three MMAs per product) on tcgen05, fp32 accumulate -- the parity-grade mode",
                     "bf16": "bf16 operands on tcgen05, fp32 accumulate -- the throughput mode", "fp32": "fp32 SIMT FMA parity mode"}[precision]
        line = {
            "metric": METRIC, "value": value, "unit": "env steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": precision, "data": "synthetic",
            "config": {"workload": f"{WORKLOADS[wl]['baseline']}: {WORKLOADS[wl]['desc']}; 6272-pt clouds (2048 robot + 4096 obstacle + 128 target), "
                                   "lock-step policy rollout (FK + cloud resample + PointNet++ + delta-q + per-step SDF sweep)",
                       "problems_per_gpu": B, "global_problems": world * B, "parallelism": f"problem-sharded x{world}",
                       "scaling_note": ("N = 1 defaults to configs[1] (tabletop), N > 1 to configs[3] (mixed scenes, ~20 % more SA2 work per "
                                        "problem): the same-workload one-GPU rate is extra['configs[3] shard'] of the N = 1 line"),
                       "rollout_length_of_config": WORKLOADS[wl]["rollout_T"], "timed_steps": K,
                       "precision": prec_desc,
                       "l2": "inputs larger than L2 (clouds 411 MB/GPU per step)", "weights": "random init, seed 0"},
            "gpu_launches": main_run["launches"], "clocks": clk,
            "e2e": {"value": e2e_value, "unit": "env steps/s", "h2d_bytes_per_step": h2d_bytes / K, "d2h_bytes_per_step": d2h_bytes / K,
                    "ms_total": e2e_ms},
            "roofline": roof,
            "stage_ms_per_step": {k: v["ms"] / K for k, v in stages.items() if v["launches"]},
            "hbm_kernels": hbm_kernels, "tensor_kernels": tensor_kernels,
            "reference_formulation_flops_per_step": FLOP_TOTAL * B,
            "reference_formulation_tflops_whole_step": FLOP_TOTAL * B * K * world / (ms / 1000.0) / 1e12,
            "collision_rate": float(main_run["gathered"][:, 0].mean().item()),
        }
        if world > 1:
            line["nccl"] = {"version": ".".join(map(str, torch.cuda.nccl.version())), "nranks": world, "collectives_in_timed_region": 1,
                            "gathered_rows": int(main_run["gathered"].shape[0])}
        if parity is not None:
            line["parity"] = parity
            line["collision_flag_match"] = parity["collision_flag_match"]
            line["fps_idx_match"] = parity["fps_idx_match"]
            line["dq_max_abs_err"] = parity["dq_max_abs_err"]
        if fast is not None:
            line["fast_mode"] = fast
        if extra:
            line["extra"] = extra
        if not args.no_cpu_baseline and world == 1:   # rank 0 at N = 1 only: the other ranks must not wait on 10 s of CPU work
            line["cpu_baseline"] = cpu_baseline(eng.cfg.seed, wl)
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
