#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gadget_primes.py tests/test_gpu_core.py -m gpu -q -x > gpurun_out/gpu_gadget4.log 2>&1
echo "tests rc=$?" >> gpurun_out/gpu_gadget4.log
tail -5 gpurun_out/gpu_gadget4.log
timeout 300 python scripts/gadget_primes_perf.py 2>&1 | tail -3
KS_PIN=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gadget_kernel -s 2 -c 1 -f -o gpurun_out/prof_gadget_r2b python scripts/ks_prof.py > gpurun_out/prof_gadget_r2b.log 2>&1
tail -3 gpurun_out/prof_gadget_r2b.log
