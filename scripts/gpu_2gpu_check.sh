#!/bin/bash
# 2-GPU check of bench.py under torchrun: workload 4 (mixed scenes, NCCL gather) with NCCL_DEBUG=INFO on stderr, then the training workload
mkdir -p gpurun_out
N=${1:-2}
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_${N}gpu_a.json 2> gpurun_out/bench${N}.err
echo rc=$?
grep -c "NCCL INFO" gpurun_out/bench${N}.err
grep -m3 "nranks\|Init COMPLETE\|bench\]" gpurun_out/bench${N}.err
wc -l gpurun_out/r2_bench_${N}gpu_a.json
python - <<PY
import json; d=json.load(open("gpurun_out/r2_bench_${N}gpu_a.json"))
for k in ("value","dtype","n_gpus","ms_per_step","e2e","config","nccl","parity","fast_mode"): print(k, d.get(k))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload 5 --steps 3 --warmup 3 --samples-per-gpu 2048 > gpurun_out/r2_train_${N}gpu_a.json 2> gpurun_out/train${N}.err
echo rc=$?; tail -3 gpurun_out/train${N}.err; cat gpurun_out/r2_train_${N}gpu_a.json
