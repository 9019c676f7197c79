#!/bin/bash
# compute-sanitizer over smoke() (one tiny rollout in the fp32, bf16x3 and bf16 modes: every kernel of the rollout path)
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_sanitize_${tool}_smoke.log 2>&1
  echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke\[" gpurun_out/r2_sanitize_${tool}_smoke.log | tail -6
done
