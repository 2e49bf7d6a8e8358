timeout 600 python -m pytest tests/test_gpu_core.py -x -q -m gpu -k "gadget or collapsed or bench_shape" 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu --no-aux 2>&1 | tail -1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l); continue
    print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])
    for k,v in d['kernels'].items(): print(k, v['launches'], round(v['avg_ms'],4), round(v['share_of_step'],3))
"
