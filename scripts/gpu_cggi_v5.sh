#!/usr/bin/env bash
# producer-less key ring (last warp out refills) in both whole-rotation CGGI kernels: parity, timing, racecheck
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cggi.py tests/test_gpu_bench_shapes.py tests/test_gpu_circuit.py -m gpu -q -x > gpurun_out/cggi_tests_v5.log 2>&1
echo "tests rc=$?" >> gpurun_out/cggi_tests_v5.log
tail -6 gpurun_out/cggi_tests_v5.log
CGGI_FL=both timeout 300 python scripts/cggi_bench.py 2>&1 | tee gpurun_out/cggi_bench_v5.log
timeout 900 compute-sanitizer --print-limit 50 --error-exitcode 7 --tool racecheck python -m pytest tests/test_gpu_cggi.py -m gpu -q -x -k "test_blind_rotate_matches_oracle or test_blind_rotate_tall_keys" > gpurun_out/san_race_cggi_v5.log 2>&1
echo "racecheck rc=$?"; grep -E "passed|failed|SUMMARY" gpurun_out/san_race_cggi_v5.log | tail -3
