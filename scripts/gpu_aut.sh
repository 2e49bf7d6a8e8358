#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_trace.py tests/test_gpu_core.py tests/test_gpu_gadget_primes.py tests/test_gpu_circuit.py -m gpu -q -x > gpurun_out/gpu_aut.log 2>&1
echo "tests rc=$?" >> gpurun_out/gpu_aut.log; tail -4 gpurun_out/gpu_aut.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-cggi 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); a=d['aux']
for k in a:
    if 'automorphism' in k or 'trace' in k: print(k, round(a[k]))
for k in a:
    if 'circuit' in k: print(k, round(a[k]['value']))
print('ks', round(d['value']))"
