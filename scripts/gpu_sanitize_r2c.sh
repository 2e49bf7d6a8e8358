#!/usr/bin/env bash
# compute-sanitizer racecheck over the rest of the GPU suite (shared-memory kernels of the HAL entry points, convolution, trace, circuit bootstrapping)
set -u
mkdir -p gpurun_out
CS="compute-sanitizer --print-limit 50 --error-exitcode 7 --tool racecheck"
for f in test_gpu_hal test_gpu_cnv test_gpu_trace test_gpu_circuit; do
  timeout 2400 $CS python -m pytest tests/$f.py -m gpu -q > gpurun_out/san_race_$f.log 2>&1
  echo "$f rc=$?"; grep -E "passed|failed|RACECHECK SUMMARY" gpurun_out/san_race_$f.log | tail -3
done
