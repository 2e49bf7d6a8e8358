import sys, time, ctypes as C
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import poulpy_b200 as pb
lib = pb.lib()
rng = np.random.default_rng(5)
stream = torch.cuda.Stream()
import os
FLS = {"fft64": ((pb.FFT64, "fft64"),), "ntt120": ((pb.NTT120, "ntt120"),)}.get(os.environ.get("CGGI_FL", ""), ((pb.FFT64, "fft64"), (pb.NTT120, "ntt120")))
for fl, nm in FLS:
    for B in (592, 2368):
        n, n_lwe, rank, block, k = 512, 687, 3, 3, 18
        m = pb.Module(n, fl)
        m.set_stream(stream.cuda_stream)
        cols = rank + 1
        per = n * cols * cols * 2 * m.prep_bytes
        brk_buf = pb.DevBuf(per * n_lwe)
        mat = rng.integers(-(1 << 17), 1 << 17, size=(1, cols, 2, cols, n), dtype=np.int64)
        one = pb.hal.VmpPMat(brk_buf, n, 1, cols, cols, 2)
        m.vmp_prepare(one, m.mat_znx_from_numpy(mat))
        for i in range(1, n_lwe):
            lib.pgb_memcpy_d2d(C.c_void_p(brk_buf.ptr + i * per), C.c_void_p(brk_buf.ptr), C.c_size_t(per))
        xpa = m.cggi_x_pow_a()
        lut = m.vec_znx_from_numpy(rng.integers(-(1 << 16), 1 << 16, size=(1, 1, n), dtype=np.int64))
        lwe = rng.integers(-n, n, size=(B, n_lwe + 1), dtype=np.int64)
        lwe_dev = pb.DevBuf(lwe.nbytes); lwe_dev.upload(lwe)
        res = m.vec_znx_alloc(cols, 1, B)
        sc = [None]
        def br():
            sc[0] = m.cggi_blind_rotate(res, lwe_dev, n_lwe, lut, one, xpa, block, k, sc[0])
        br(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            l0 = m.launch_count
            e0.record(stream); br(); br(); e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        print(nm, "B", B, "ms", round(ms, 2), "bootstraps/s", round(B / ms * 1e3), "launches/call", (m.launch_count - l0) // 2)
