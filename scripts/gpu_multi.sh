#!/usr/bin/env bash
# multi-GPU pass: N = number of GPUs of the box (gpurun --gpus N)
set -u
N=${N:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_core.py -m gpu -q -k "sharded" > gpurun_out/multi_tests_n$N.log 2>&1; tail -3 gpurun_out/multi_tests_n$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/e2e_probe_ranks.py > gpurun_out/e2e_probe_n$N.json 2> gpurun_out/e2e_probe_n$N.err
tail -1 gpurun_out/e2e_probe_n$N.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench rc=$?"; tail -c 400 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.json').read().strip().splitlines()[-1])
print('n_gpus',d['n_gpus'],'value',d['value'],'e2e',d['e2e']['value'])
c=d['cggi']
for k in ('fft64','ntt120'): print(k, c[k]['value'], c[k]['e2e']['value'])
PY
