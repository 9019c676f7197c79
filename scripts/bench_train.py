#!/usr/bin/env python
"""Training-step measurement (BASELINE.json configs[4]: PointNetEncoder fwd+bwd + CollisionLoss, DDP), fp32 path.

    python scripts/bench_train.py --samples-per-gpu 1024 --steps 3 --warmup 1
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_train.py --gpus N ...

One step = mpn_train_step_grads (forward with saved state, both losses, backward to all 19.07 M parameters) -> NCCL
all-reduce-mean of the flat gradient vector (N > 1) -> clip_grad_norm_(1.0) + Adam.  Prints one JSON line on rank 0
(samples/s over all ranks, per-phase milliseconds from CUDA events, max over ranks).  Not the headline bench (bench.py)."""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--samples-per-gpu", type=int, default=1024)
    ap.add_argument("--precision", default="fp32", choices=["fp32", "bf16"])
    args = ap.parse_args()
    os.environ.pop("NCCL_DEBUG", None)   # NCCL's version banner goes to stdout; keep it to the JSON line
    import torch
    import torch.distributed as dist
    from mpinets_b200 import scenes, _lib
    from mpinets_b200.engine import Engine
    from mpinets_b200.parallel import allreduce_mean_
    from oracle import oracle as O   # weights init only

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    B = args.samples_per_gpu
    eng = Engine(device=local)
    eng.load_state_dict(O.reference_state_dict(0))
    p = scenes.config_problems(4, B, problem0=rank * B)
    d = {k: torch.from_numpy(np.ascontiguousarray(p[k])).cuda() for k in scenes.SCENE_KEYS + ("q0", "target")}
    sc = {k: d[k] for k in scenes.SCENE_KEYS}
    cloud = eng.build_cloud(sc, d["q0"], d["target"], problem0=rank * B)
    qn = eng.normalize(d["q0"])
    gen = torch.Generator(device="cuda").manual_seed(rank)
    sup = torch.clamp(qn + 0.05 * torch.randn(qn.shape, generator=gen, device="cuda"), -1, 1)
    grads = torch.empty(eng.param_count, device="cuda")
    prec = _lib.PREC_BF16 if args.precision == "bf16" else _lib.PREC_FP32

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def step(i, timed=None):
        e = [ev() for _ in range(5)]
        e[0].record()
        eng.train_step_grads(sc, cloud, qn, sup, need_grad=False, precision=prec)
        e[1].record()
        losses, _, _ = eng.train_step_grads(sc, cloud, qn, sup, grads=grads, precision=prec)
        e[2].record()
        allreduce_mean_(grads)
        e[3].record()
        eng.adam_step(grads, i + 1)
        e[4].record()
        if timed is not None:
            timed.append(e)
        return losses

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    rec, l0 = [], eng.launch_count
    for i in range(args.steps):
        losses = step(args.warmup + i, rec)
    torch.cuda.synchronize()
    launches = eng.launch_count - l0
    ph = np.array([[e[j].elapsed_time(e[j + 1]) for j in range(4)] for e in rec]).mean(axis=0)   # fwd-only, fwd+bwd, allreduce, adam
    t = torch.tensor([ph[1] + ph[2] + ph[3], ph[0], ph[1], ph[2], ph[3]], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t = t.cpu().numpy()
    if rank == 0:
        print(json.dumps({
            "metric": "training samples/sec (fwd + losses + bwd + DDP all-reduce + clip + Adam), " + args.precision, "value": world * B / (t[0] / 1000.0),
            "unit": "samples/s", "n_gpus": world, "samples_per_gpu": B, "steps": args.steps, "ms_per_step": float(t[0]),
            "phases_ms": {"forward_and_losses_only": float(t[1]), "forward_backward": float(t[2]), "grad_allreduce": float(t[3]),
                          "clip_adam_transposes": float(t[4])},
            "gpu_launches_per_step": launches / args.steps / 2,   # the forward-only probe doubles the forward launches
            "losses": [float(x) for x in losses.cpu()], "dtype": args.precision, "data": "synthetic (config-4 scene mix)",
            "params": 19068103}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
