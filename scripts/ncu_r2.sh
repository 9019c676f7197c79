#!/bin/bash
# Round-2 ncu evidence (1 GPU): --set full of the x3 / bf16 encoder kernels at the BENCHMARKED size (4096 problems), then the launch
# list of the bench command and of one training step.  Reports land in gpurun_out/ (scratch); summaries are made here with
# scripts/ncu_summary.py.
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on \
  -k regex:"sa1x3_tc_kernel|sa2x3h_tc_kernel|gemm_tma_kernel|fps_pruned_kernel" -c 12 -o gpurun_out/r2_x3_kernels -f \
  python scripts/prof_step.py 4096 bf16x3 1 > gpurun_out/ncu_x3.log 2>&1
echo "x3 rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on \
  -k regex:"sa1t_tc_kernel|sa2w3_tc_kernel|gemm_tma_kernel" -c 8 -o gpurun_out/r2_bf16_kernels -f \
  python scripts/prof_step.py 4096 bf16 1 > gpurun_out/ncu_bf16.log 2>&1
echo "bf16 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "launches rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r2_train_launches_b.csv \
  python bench.py --workload 5 --steps 1 --warmup 1 --samples-per-gpu 2048 > gpurun_out/ncu_train.log 2>&1
echo "train launches rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches_bench.csv gpurun_out/r2_train_launches_b.csv
