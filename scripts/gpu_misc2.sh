#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
./scripts/_bin/mma_ntt_probe > gpurun_out/r2_mma_ntt_probe.json 2>&1; cat gpurun_out/r2_mma_ntt_probe.json
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r2b_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2b_bench.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'frac',d['roofline']['frac'])
c=d['cggi']
for k in ('fft64','ntt120'): print(k, c[k]['value'], c[k]['e2e']['value'], c[k]['roofline']['frac'], c[k]['launches_per_batch'])
print(c.get('cpu_baseline'))
print(d.get('cpu_baseline_scalar',{}).get('value'), d.get('cpu_baseline_avx',{}).get('value'))
PY
