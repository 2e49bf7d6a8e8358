#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/gpu_tests.log
tail -6 gpurun_out/gpu_tests.log
timeout 600 python scripts/vmp_stream.py 2>&1 | tee gpurun_out/vmp_stream.log
