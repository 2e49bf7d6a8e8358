"""Key-switch with dsize = 2 (n = 4096, base2k = 18, a of 4 limbs, key (2, 1, 2, 4)): fused single kernel vs the limb-wise HAL sequence."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poulpy_b200 as pb
n, k, B, dsize = 4096, 18, 4096, 2
stream = torch.cuda.Stream()
rng = np.random.default_rng(1)
m = pb.Module(n, pb.FFT64 if os.environ.get("KS_FLAVOUR") == "fft64" else pb.NTT120); m.set_stream(stream.cuda_stream)
mat = rng.integers(-(1 << 17), 1 << 17, size=(2, 1, 4, 2, n), dtype=np.int64)
pm = m.vmp_pmat_alloc(2, 1, 2, 4)
m.vmp_prepare(pm, m.mat_znx_from_numpy(mat))
a = m.vec_znx_alloc(2, 4, B)
a.buf.upload(rng.integers(-(1 << 17), 1 << 17, size=(64 * 4 * 2 * n,), dtype=np.int64))
r = m.vec_znx_alloc(2, 4, B)
for env in (None, "1"):
    m.set_option(pb.hal.OPT_NO_FUSION, 1 if env else 0)
    sc = None
    for _ in range(3): sc = m.glwe_keyswitch(r, k, a, k, pm, k, dsize, sc)
    torch.cuda.synchronize()
    l0 = m.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(10): sc = m.glwe_keyswitch(r, k, a, k, pm, k, dsize, sc)
        e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print("dsize 2", "unfused" if env else "fused  ", "ms/batch", round(ms, 3), "key-switches/s", round(B / ms * 1e3), "launches/call", (m.launch_count - l0) // 10)
