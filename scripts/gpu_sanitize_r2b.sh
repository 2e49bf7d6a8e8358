#!/usr/bin/env bash
# compute-sanitizer memcheck over the rest of the GPU suite (HAL entry points, convolution / tensor, circuit bootstrapping, goldens, serialisation)
set -u
mkdir -p gpurun_out
CS="compute-sanitizer --print-limit 50 --error-exitcode 7 --tool memcheck"
timeout 2400 $CS python -m pytest tests/test_gpu_hal.py tests/test_gpu_cnv.py tests/test_golden.py tests/test_serialize.py -m gpu -q -x > gpurun_out/san_mem_hal.log 2>&1
echo "hal rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/san_mem_hal.log | tail -3
timeout 2400 $CS python -m pytest tests/test_gpu_circuit.py -m gpu -q -x > gpurun_out/san_mem_circuit.log 2>&1
echo "circuit rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/san_mem_circuit.log | tail -3
timeout 1200 $CS python -m pytest tests/test_gpu_bench_shapes.py -m gpu -q -x -k "keyswitch_and_external or ckks" > gpurun_out/san_mem_shapes.log 2>&1
echo "shapes rc=$?"; grep -E "passed|failed|ERROR SUMMARY" gpurun_out/san_mem_shapes.log | tail -3
