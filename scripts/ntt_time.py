"""Forward / inverse limb rates of the stand-alone transforms at a few sizes: python scripts/ntt_time.py [ntt120|fft64]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poulpy_b200 as pb
fl = pb.FFT64 if (len(sys.argv) > 1 and sys.argv[1] == "fft64") else pb.NTT120
stream = torch.cuda.Stream(); rng = np.random.default_rng(5)
for log_n, size in ((10, 2), (11, 4), (12, 8), (13, 16)):
    n = 1 << log_n; m = pb.Module(n, fl); m.set_stream(stream.cuda_stream)
    B = (256 << 20) // (n * 2 * size * 8); a = m.vec_znx_alloc(2, size, B)
    a.buf.upload(rng.integers(-(1 << 17), 1 << 17, size=(n * 2 * size,), dtype=np.int64))
    d = m.vec_znx_dft_alloc(2, size, B); big = m.vec_znx_big_alloc(2, size, B)
    def fwd():
        for c in range(2): m.vec_znx_dft_apply(1, 0, d, c, a, c)
    def inv():
        for c in range(2): m.vec_znx_idft_apply(big, c, d, c)
    out = []
    for f in (fwd, inv):
        for _ in range(2): f()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(5): f()
            e1.record(stream)
        torch.cuda.synchronize()
        out.append(B * 2 * size / (e0.elapsed_time(e1) / 5 * 1e-3))
    print("log_n", log_n, "fwd limbs/s %.3g" % out[0], "inv limbs/s %.3g" % out[1])
