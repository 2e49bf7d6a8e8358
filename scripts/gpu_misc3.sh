#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_circuit.py tests/test_gpu_bench_shapes.py tests/test_gpu_hal.py -m gpu -q > gpurun_out/gpu_tests3.log 2>&1
echo "tests rc=$?" >> gpurun_out/gpu_tests3.log
tail -8 gpurun_out/gpu_tests3.log
KS_PIN=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:gadget_kernel -s 2 -c 1 -f -o gpurun_out/prof_gadget_r2 python scripts/ks_prof.py > gpurun_out/prof_gadget_r2.log 2>&1
tail -3 gpurun_out/prof_gadget_r2.log
