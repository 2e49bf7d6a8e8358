#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
for b in 256 1024 4096 16384; do
  timeout 300 python bench.py --batch $b --steps 20 --warmup 3 --no-cpu --no-aux --no-cggi 2> /dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('batch', d['config']['batch_per_gpu'], 'key-switches/s', round(d['value']), 'ms_per_step', round(d['ms_per_step'],4), 'roofline', round(d['roofline']['frac'],3))"
done | tee gpurun_out/batch_sizes.log
