"""A few batched forward / inverse transforms for ncu: python scripts/dft_prof.py [fft64|ntt120] [log_n]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poulpy_b200 as pb
fl = pb.FFT64 if (len(sys.argv) > 1 and sys.argv[1] == "fft64") else pb.NTT120
log_n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
n, size = 1 << log_n, 8
m = pb.Module(n, fl)
B = max(1, (256 << 20) // (n * 2 * size * 8))
a = m.vec_znx_alloc(2, size, B)
a.buf.upload(np.random.default_rng(1).integers(-(1 << 17), 1 << 17, size=(n * 2 * size,), dtype=np.int64))
d = m.vec_znx_dft_alloc(2, size, B); big = m.vec_znx_big_alloc(2, size, B)
for _ in range(2):
    m.vec_znx_dft_apply(1, 0, d, 0, a, 0)
    m.vec_znx_idft_apply(big, 0, d, 0)
m.sync()
print("done")
