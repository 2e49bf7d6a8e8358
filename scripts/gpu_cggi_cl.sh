#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_cggi.py -m gpu -q -x -k "cluster_multicast" > gpurun_out/cggi_cl.log 2>&1
echo "tests rc=$?" >> gpurun_out/cggi_cl.log; tail -5 gpurun_out/cggi_cl.log
CGGI_FL=fft64 timeout 300 python scripts/cggi_bench.py 2>&1 | tail -2
PGB_CGGI_CLUSTER=2 CGGI_FL=fft64 timeout 300 python scripts/cggi_bench.py 2>&1 | tail -2
