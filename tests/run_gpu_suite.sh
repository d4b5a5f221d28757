#!/bin/bash
# Runs every GPU test file in its own process (a trapped kernel poisons the CUDA context) and keeps the logs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
rc=0
for f in elementwise gemm attention model parallel step_end; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -x -q -rA -s -p no:cacheprovider > gpurun_out/t_$f.log 2>&1
  e=$?; echo "== test_gpu_$f exit $e"; [ $e -ne 0 ] && rc=1
  grep -E "passed|failed|error" gpurun_out/t_$f.log | tail -3
done
exit $rc
