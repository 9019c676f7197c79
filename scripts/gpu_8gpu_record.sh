#!/bin/bash
# 8-GPU records (one node): bench.py default at N = 8 (workload 4 = configs[3]: 32768 mixed problems, one NCCL gather) and the training
# step of configs[4] (8 x 8192 samples, DDP all-reduce)
mkdir -p gpurun_out
N=${1:-8}
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/bench${N}.err
echo "bench rc=$?"; grep -c "NCCL INFO" gpurun_out/bench${N}.err; grep -m2 "NVLS\|nranks" gpurun_out/bench${N}.err | cut -c1-200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --workload 5 --steps 3 --warmup 3 --samples-per-gpu 8192 > gpurun_out/r2_train_${N}gpu.json 2> gpurun_out/train${N}.err
echo "train rc=$?"; tail -2 gpurun_out/train${N}.err | cut -c1-300
python - <<PY
import json
for f in ("gpurun_out/r2_bench_${N}gpu.json", "gpurun_out/r2_train_${N}gpu.json"):
    try:
        d = json.load(open(f))
        print(f, {k: d.get(k) for k in ("value", "unit", "dtype", "n_gpus", "ms_per_step", "phases_ms", "nccl", "collision_flag_match", "dq_max_abs_err")}, (d.get("fast_mode") or {}).get("value"), (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
