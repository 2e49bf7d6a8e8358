#!/usr/bin/env bash
# full GPU suite + smoke + default bench line + reference arm (round 2, final build)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_v4.log 2>&1
echo "tests rc=$?" >> gpurun_out/gpu_tests_v4.log
tail -6 gpurun_out/gpu_tests_v4.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v4.json 2> gpurun_out/bench_v4.err
echo "bench rc=$?"
tail -c 600 gpurun_out/bench_v4.err
head -c 300 gpurun_out/bench_v4.json; echo
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/ref_v4.json 2> gpurun_out/ref_v4.err
echo "ref rc=$?"
head -c 200 gpurun_out/ref_v4.json; echo
KS_BATCH=1024 timeout 120 python - <<'PY'
import os, sys, time
import numpy as np
sys.path.insert(0, ".")
import poulpy_b200 as pb
n, k = 4096, 18
rng = np.random.default_rng(1)
m = pb.Module(n, pb.NTT120)
pm = m.vmp_pmat_alloc(3, 1, 2, 4)
m.vmp_prepare(pm, m.mat_znx_from_numpy(rng.integers(-(1 << 17), 1 << 17, size=(3, 1, 4, 2, n), dtype=np.int64)))
m.gadget_key_pin(pm)
for B in (256, 1024, 4096):
    a = m.vec_znx_from_numpy(rng.integers(-(1 << 17), 1 << 17, size=(B, 3, 2, n), dtype=np.int64))
    r = m.vec_znx_alloc(2, 3, B)
    sc = None
    for _ in range(3):
        sc = m.glwe_keyswitch(r, k, a, k, pm, k, 1, sc)
    m.sync()
    ts = []
    for _ in range(20):
        t0 = time.perf_counter(); sc = m.glwe_keyswitch(r, k, a, k, pm, k, 1, sc); m.sync(); ts.append(time.perf_counter() - t0)
    print("batch", B, "key-switches/s", round(B / np.median(ts)))
PY
