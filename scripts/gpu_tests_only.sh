#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q ${PYTEST_ARGS:-} > gpurun_out/gpu_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/gpu_tests.log
tail -40 gpurun_out/gpu_tests.log
