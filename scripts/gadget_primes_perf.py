"""Three-prime vs four-prime launch of the NTT120 gadget kernel (pinned key): external product at the BASELINE config (n = 2048, GGSW
3 x 2 x 2 x 3, batch 4096) and a key-switch with a three-limb key at n = 4096.  CUDA-event timings, never under a profiler."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poulpy_b200 as pb
from poulpy_b200 import hal as H


def timed(fn, m, iters=10):
    """median wall-clock ms of one synchronised call (batches of >= 1 ms: the ~10 us of launch + sync do not matter)"""
    for _ in range(3):
        fn()
    ms = []
    for _ in range(iters):
        m.sync()
        t0 = time.perf_counter()
        fn()
        m.sync()
        ms.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ms))


out = {}
rng = np.random.default_rng(3)
for name, n, rows, cols_in, cols_out, ksize, a_size, ext in (("external_product_n2048", 2048, 3, 2, 2, 3, 3, True),
                                                            ("keyswitch_n4096_key3", 4096, 3, 1, 2, 3, 3, False),
                                                            ("keyswitch_n4096_key4_headline", 4096, 3, 1, 2, 4, 3, False)):
    B, k = 4096, 18
    m = pb.Module(n, pb.NTT120)
    mat = rng.integers(-(1 << 17), 1 << 17, size=(rows, cols_in, ksize, cols_out, n), dtype=np.int64)
    pm = m.vmp_pmat_alloc(rows, cols_in, cols_out, ksize)
    m.vmp_prepare(pm, m.mat_znx_from_numpy(mat))
    m.gadget_key_pin(pm)
    a = m.vec_znx_from_numpy(rng.integers(-(1 << 17), 1 << 17, size=(B, a_size, 2, n), dtype=np.int64))
    r = m.vec_znx_alloc(2, 3, B)
    sc = [None]

    def run():
        if ext:
            sc[0] = m.glwe_external_product(r, k, a, k, pm, k, 1, sc[0])
        else:
            sc[0] = m.glwe_keyswitch(r, k, a, k, pm, k, 1, sc[0])

    res = {}
    for force in (0, 4):
        m.set_option(H.OPT_GADGET_PRIMES, force)
        ms = timed(run, m)
        res["primes_%d" % m.get_option(H.OPT_LAST_GADGET_PRIMES) + ("_forced" if force else "")] = {"ms_per_batch": ms, "per_s": B / ms * 1e3}
    out[name] = res
    print(name, json.dumps(res), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/gadget_primes_perf.json", "w"), indent=1)
