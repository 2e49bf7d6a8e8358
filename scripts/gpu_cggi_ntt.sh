#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cggi.py tests/test_gpu_bench_shapes.py tests/test_golden.py tests/test_gpu_circuit.py -m gpu -q -x > gpurun_out/cggi_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/cggi_tests.log
tail -8 gpurun_out/cggi_tests.log
CGGI_FL=${CGGI_FL:-ntt120} timeout 300 python scripts/cggi_bench.py 2>&1 | tee gpurun_out/cggi_bench.log
if [ -n "${TAG:-}" ]; then
CGGI_FL=${CGGI_FL:-ntt120} timeout 600 ncu --set full --clock-control none --import-source on -k regex:cggi_fused -s 1 -c 1 -f -o gpurun_out/prof_cggi_${CGGI_FL:-ntt120}_${TAG} python scripts/cggi_prof.py > gpurun_out/prof_cggi.log 2>&1
tail -3 gpurun_out/prof_cggi.log
fi
