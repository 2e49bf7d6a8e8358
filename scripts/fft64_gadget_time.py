"""Times the FFT64 key-switch (n = 4096) and external product (n = 2048) with CUDA events, fused kernel vs unfused HAL sequence:
python scripts/fft64_gadget_time.py [batch]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poulpy_b200 as pb
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
k = 18
stream = torch.cuda.Stream()
rng = np.random.default_rng(1)
def timeit(f, reps=10):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(reps): f()
        e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for name, n, ext in (("keyswitch n=4096", 4096, False), ("ext product n=2048", 2048, True)):
    m = pb.Module(n, pb.FFT64); m.set_stream(stream.cuda_stream)
    if ext:
        mat = rng.integers(-(1 << 17), 1 << 17, size=(3, 2, 3, 2, n), dtype=np.int64)
        pm = m.vmp_pmat_alloc(3, 2, 2, 3)
    else:
        mat = rng.integers(-(1 << 17), 1 << 17, size=(3, 1, 4, 2, n), dtype=np.int64)
        pm = m.vmp_pmat_alloc(3, 1, 2, 4)
    m.vmp_prepare(pm, m.mat_znx_from_numpy(mat))
    a = m.vec_znx_alloc(2, 3, B)
    a.buf.upload(rng.integers(-(1 << 17), 1 << 17, size=(min(B, 64) * 3 * 2 * n,), dtype=np.int64))
    r = m.vec_znx_alloc(2, 3, B)
    fn = m.glwe_external_product if ext else m.glwe_keyswitch
    for env in (None, "1"):
        m.set_option(pb.hal.OPT_NO_FUSION, 1 if env else 0)
        sc = [None]
        def f(): sc[0] = fn(r, k, a, k, pm, k, 1, sc[0])
        ms = timeit(f)
        print(name, "unfused" if env else "fused  ", "ms/batch", round(ms, 4), "per s", round(B / ms * 1e3))
    m.set_option(pb.hal.OPT_NO_FUSION, 0)
