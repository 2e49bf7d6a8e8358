#!/usr/bin/env bash
# full GPU suite + default bench line + reference arm (round 2, version 3)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_v3.log 2>&1
echo "tests rc=$?" >> gpurun_out/gpu_tests_v3.log
tail -12 gpurun_out/gpu_tests_v3.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_v3.err
head -c 1200 gpurun_out/bench_v3.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/ref_v3.json 2> gpurun_out/ref_v3.err
echo "ref rc=$?"
head -c 600 gpurun_out/ref_v3.json
