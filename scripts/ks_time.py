"""Times the batched key-switch (bench shape) with CUDA events through the profiler: python scripts/ks_time.py [batch]"""
import os, sys, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poulpy_b200 as pb
n, k, B = 4096, 18, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
m = pb.Module(n, pb.NTT120)
rng = np.random.default_rng(1)
mat = rng.integers(-(1 << 17), 1 << 17, size=(3, 1, 4, 2, n), dtype=np.int64)
pm = m.vmp_pmat_alloc(3, 1, 2, 4)
m.vmp_prepare(pm, m.mat_znx_from_numpy(mat))
a = m.vec_znx_from_numpy(rng.integers(-(1 << 17), 1 << 17, size=(B, 3, 2, n), dtype=np.int64))
r = m.vec_znx_alloc(2, 3, B)
sc = None
for _ in range(3):
    sc = m.glwe_keyswitch(r, k, a, k, pm, k, 1, sc)
m.sync()
lib = pb.lib()
lib.pgb_profile_enable(m._h, 1)
for _ in range(10):
    sc = m.glwe_keyswitch(r, k, a, k, pm, k, 1, sc)
ms = (C.c_double * 7)(); cnt = (C.c_uint64 * 7)()
lib.pgb_profile_read(m._h, ms, cnt, 1)
tot = sum(ms)
print(os.environ.get("PGB_GADGET_MB"), os.environ.get("PGB_GADGET_CARVEOUT"), "gadget ms", round(ms[6] / max(1, cnt[6]), 4), "total/step", round(tot / 10, 4), "ks/s", round(B / (tot / 10) * 1e3))
