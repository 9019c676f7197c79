#!/bin/bash
# weak-scaling record at N GPUs of one node: bench.py default (workload 4 = configs[3]: 4096 mixed problems per GPU, one NCCL gather)
mkdir -p gpurun_out
N=${1:-2}
NCCL_DEBUG=INFO timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/bench${N}.err
echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/r2_bench_${N}gpu.json"))
print({k: d.get(k) for k in ("value", "dtype", "n_gpus", "ms_per_step", "collision_flag_match", "dq_max_abs_err")}, d["e2e"]["value"], d["fast_mode"]["value"], d.get("nccl"))
PY
