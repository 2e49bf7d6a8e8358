#!/usr/bin/env bash
# first GPU pass of round 2: full GPU suite (incl. the bench-shape parity tests), default bench line, launch list
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2a_tests.log
tail -15 gpurun_out/r2a_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/r2a_bench.err
head -c 1500 gpurun_out/r2a_bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err
