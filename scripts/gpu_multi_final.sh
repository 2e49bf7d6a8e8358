#!/usr/bin/env bash
# multi-GPU bench lines of the final round-2 build: N = number of GPUs of the box (gpurun --gpus N)
set -u
N=${N:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_v4_n$N.json 2> gpurun_out/bench_v4_n$N.err
echo "bench rc=$?"; tail -c 400 gpurun_out/bench_v4_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_v4_n$N.json').read().strip().splitlines()[-1])
print('n_gpus',d['n_gpus'],'value',d['value'],'e2e',d['e2e']['value'])
c=d['cggi']
for k in ('fft64','ntt120'): print(k, c[k]['value'], c[k]['e2e']['value'])
PY
