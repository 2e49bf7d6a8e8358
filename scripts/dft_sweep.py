"""NTT120 / FFT64 forward + inverse limb rates at log_n 14..16 (the large-n paths): python scripts/dft_sweep.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poulpy_b200 as pb
stream = torch.cuda.Stream()
rng = np.random.default_rng(5)
for fl, nm in ((pb.NTT120, "ntt120"), (pb.FFT64, "fft64")):
    for log_n in (13, 14, 15, 16):
        n, size = 1 << log_n, 8
        m = pb.Module(n, fl); m.set_stream(stream.cuda_stream)
        B = max(1, (256 << 20) // (n * 2 * size * 8))
        a = m.vec_znx_alloc(2, size, B)
        a.buf.upload(rng.integers(-(1 << 17), 1 << 17, size=(n * 2 * size,), dtype=np.int64))
        d = m.vec_znx_dft_alloc(2, size, B); big = m.vec_znx_big_alloc(2, size, B)
        def fwd():
            for c in range(2): m.vec_znx_dft_apply(1, 0, d, c, a, c)
        def inv():
            for c in range(2): m.vec_znx_idft_apply(big, c, d, c)
        res = []
        for f in (fwd, inv):
            for _ in range(2): f()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream)
                for _ in range(5): f()
                e1.record(stream)
            torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) / 5)
        limbs = B * 2 * size
        print(nm, "log_n", log_n, "fwd limbs/s %.3g (%.0f GB/s)" % (limbs / res[0] * 1e3, limbs * n * (8 + m.prep_bytes) / res[0] / 1e6),
              "inv limbs/s %.3g (%.0f GB/s)" % (limbs / res[1] * 1e3, limbs * n * (m.prep_bytes + m.big_bytes) / res[1] / 1e6), os.environ.get("PGB_NTT_TWO_KERNEL", ""))
