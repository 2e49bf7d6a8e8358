"""Profiling driver: a few batched key-switches at the bench shape (n=4096, NTT120; KS_FLAVOUR=fft64 for the FFT64 kernel).  Usage under ncu:
ncu --set full --clock-control none --import-source on -k regex:gadget_kernel -s 2 -c 1 -o gpurun_out/prof_gadget python scripts/ks_prof.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poulpy_b200 as pb

n, k, B = 4096, 18, int(os.environ.get("KS_BATCH", "2048"))
m = pb.Module(n, pb.FFT64 if os.environ.get("KS_FLAVOUR") == "fft64" else pb.NTT120)
rng = np.random.default_rng(1)
KS = 3 if os.environ.get("KS_KEY3") else 4  # KS_KEY3: a three-limb key (the three-prime launch when pinned)
mat = rng.integers(-(1 << 17), 1 << 17, size=(3, 1, KS, 2, n), dtype=np.int64)
pm = m.vmp_pmat_alloc(3, 1, 2, KS)
m.vmp_prepare(pm, m.mat_znx_from_numpy(mat))
if os.environ.get("KS_PIN"):
    m.gadget_key_pin(pm)
a = m.vec_znx_from_numpy(rng.integers(-(1 << 17), 1 << 17, size=(B, 3, 2, n), dtype=np.int64))
r = m.vec_znx_alloc(2, 3, B)
sc = None
for _ in range(4):
    if os.environ.get("KS_OP") == "aut":  # the automorphism epilogue: res = normalize(aut_5(ks(a)) + a)
        sc = m.glwe_automorphism_op(0, r, k, a, pm, k, 5, 1, sc)
    else:
        sc = m.glwe_keyswitch(r, k, a, k, pm, k, 1, sc)
m.sync()
print("done")
