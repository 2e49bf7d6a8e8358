#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_gadget_primes.py tests/test_gpu_core.py tests/test_gpu_trace.py -m gpu -q -x > gpurun_out/gpu_gadget3.log 2>&1
echo "tests rc=$?" >> gpurun_out/gpu_gadget3.log
tail -30 gpurun_out/gpu_gadget3.log
timeout 300 python scripts/gadget_primes_perf.py 2>&1 | tail -5
