#!/bin/bash
# training-step record (configs[4]) at N GPUs of one node: bench.py --workload 5, SAMPLES per GPU (default 8192)
mkdir -p gpurun_out
N=${1:-1}; S=${2:-8192}
if [ "$N" = "1" ]; then
  timeout 900 python bench.py --workload 5 --steps 3 --warmup 3 --samples-per-gpu $S > gpurun_out/r2_train_${N}gpu_${S}.json 2> gpurun_out/train${N}.err
else
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --workload 5 --steps 3 --warmup 3 --samples-per-gpu $S > gpurun_out/r2_train_${N}gpu_${S}.json 2> gpurun_out/train${N}.err
fi
echo "train rc=$?"; tail -2 gpurun_out/train${N}.err | cut -c1-300
python - <<PY
import json
d = json.load(open("gpurun_out/r2_train_${N}gpu_${S}.json"))
print({k: d.get(k) for k in ("value", "unit", "n_gpus", "ms_per_step", "phases_ms", "gpu_launches")}, d["e2e"]["value"])
PY
