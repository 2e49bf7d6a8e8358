"""Per-category device time of glwe_trace (n = 4096, base2k 18, 12 rounds) and of one glwe_automorphism / automorphism_op call."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poulpy_b200 as pb

n, B, k = 4096, 1024, 18
fl = pb.FFT64 if os.environ.get("KS_FLAVOUR") == "fft64" else pb.NTT120
rng = np.random.default_rng(1)
m = pb.Module(n, fl)
keys = []
for _ in range(12):
    pm = m.vmp_pmat_alloc(3, 1, 2, 4)
    m.vmp_prepare(pm, m.mat_znx_from_numpy(rng.integers(-(1 << 17), 1 << 17, size=(3, 1, 4, 2, n), dtype=np.int64)))
    m.gadget_key_pin(pm)
    keys.append(pm)
a = m.vec_znx_from_numpy(rng.integers(-(1 << 17), 1 << 17, size=(B, 3, 2, n), dtype=np.int64))
r = m.vec_znx_alloc(2, 3, B)
lib = pb.lib()
lib.pgb_profile_category_name.restype = C.c_char_p
sc = [None, None, None]


def f_aut():
    sc[0] = m.glwe_automorphism(r, k, a, k, keys[0], k, 5, 1, sc[0])


def f_op():
    sc[1] = m.glwe_automorphism_op(0, r, k, a, keys[0], k, 5, 1, sc[1])


def f_tr():
    sc[2] = m.glwe_trace_assign(r, k, 0, keys, k, 1, sc[2])


for name, fn in (("glwe_automorphism", f_aut), ("glwe_automorphism_op(add)", f_op), ("glwe_trace (12 rounds)", f_tr)):
    fn(); fn(); m.sync()
    lib.pgb_profile_enable(m._h, 1)
    for _ in range(3):
        fn()
    ms = (C.c_double * 7)(); cnt = (C.c_uint64 * 7)()
    lib.pgb_profile_read(m._h, ms, cnt, 1)
    lib.pgb_profile_enable(m._h, 0)
    print(name, "total ms/call", round(sum(ms) / 3, 3),
          {lib.pgb_profile_category_name(i).decode(): (round(ms[i] / 3, 3), cnt[i] // 3) for i in range(7) if cnt[i]}, flush=True)
