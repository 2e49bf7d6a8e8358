#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cggi.py tests/test_gpu_bench_shapes.py tests/test_golden.py tests/test_gpu_circuit.py -m gpu -q -x > gpurun_out/cggi_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/cggi_tests.log
tail -25 gpurun_out/cggi_tests.log
timeout 300 python scripts/cggi_bench.py 2>&1 | tee gpurun_out/cggi_bench.log
