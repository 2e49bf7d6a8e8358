import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import poulpy_b200 as pb
lib = pb.lib()
rng = np.random.default_rng(5)
n, n_lwe, rank, block, k, B = 512, 60, 3, 3, 18, 592
m = pb.Module(n, pb.NTT120 if os.environ.get('CGGI_FL') == 'ntt120' else pb.FFT64)
cols = rank + 1
per = n * cols * cols * 2 * m.prep_bytes
brk_buf = pb.DevBuf(per * n_lwe)
mat = rng.integers(-(1 << 17), 1 << 17, size=(1, cols, 2, cols, n), dtype=np.int64)
one = pb.hal.VmpPMat(brk_buf, n, 1, cols, cols, 2)
m.vmp_prepare(one, m.mat_znx_from_numpy(mat))
for i in range(1, n_lwe):
    lib.pgb_memcpy_d2d(C.c_void_p(brk_buf.ptr + i * per), C.c_void_p(brk_buf.ptr), C.c_size_t(per))
m.gadget_key_pin(one)
xpa = m.cggi_x_pow_a()
lut = m.vec_znx_from_numpy(rng.integers(-(1 << 16), 1 << 16, size=(1, 1, n), dtype=np.int64))
lwe = rng.integers(-n, n, size=(B, n_lwe + 1), dtype=np.int64)
lwe_dev = pb.DevBuf(lwe.nbytes); lwe_dev.upload(lwe)
res = m.vec_znx_alloc(cols, 1, B)
m.cggi_blind_rotate(res, lwe_dev, n_lwe, lut, one, xpa, block, k)
m.sync()
m.cggi_blind_rotate(res, lwe_dev, n_lwe, lut, one, xpa, block, k)
m.sync()
