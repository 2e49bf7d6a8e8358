#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gadget_primes.py tests/test_gpu_core.py tests/test_gpu_trace.py -m gpu -q -x > gpurun_out/gpu_gadget4.log 2>&1
echo "tests rc=$?" >> gpurun_out/gpu_gadget4.log
tail -5 gpurun_out/gpu_gadget4.log
timeout 300 python scripts/gadget_primes_perf.py 2>&1 | tail -3
