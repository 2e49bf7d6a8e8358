"""vmp_apply_dft_to_dft in the HBM-streaming regime (many products per launch, each with its own matrix), both flavours, per forced
output-polys-per-thread setting: python scripts/vmp_stream.py"""
import os, sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import poulpy_b200 as pb
lib = pb.lib()
stream = torch.cuda.Stream()
peak = 6541.5
shapes = ((12, 7, 1, 2, 8), (13, 15, 1, 2, 16), (14, 31, 1, 2, 32), (15, 14, 1, 2, 15))
for fl, nm in ((pb.FFT64, "fft64"), (pb.NTT120, "ntt120")):
    for (log_n, rows, cols_in, cols_out, size) in shapes:
        n = 1 << log_n
        m = pb.Module(n, fl); m.set_stream(stream.cuda_stream)
        s = m.prep_bytes
        R, Cc = rows * cols_in, cols_out * size
        byts = (R + R * Cc + Cc) * n * s
        count = int(max(1, min(4096, -(-(1 << 30) // byts))))
        pm_one = n * R * Cc * s
        pm = pb.DevBuf(pm_one * count)
        a = m.vec_znx_dft_alloc(cols_in, rows, count); r = m.vec_znx_dft_alloc(cols_out, size, count)
        rs, as_ = r.struct(), a.struct()
        ps = pb.hal._PM(pm.ptr, n, size, rows, cols_in, cols_out)
        out = []
        for cnt in (count, 1):
            bt = pb.hal._BT(cnt, r.batch_stride, a.batch_stride, pm_one)
            for ct in (0, 2, 4):
                m.set_option(pb.hal.OPT_VMP_CT, ct)
                def f(): pb.hal._check(lib.pgb_vmp_apply_dft_to_dft_batched(m._h, C.byref(rs), C.byref(as_), C.byref(ps), C.c_uint64(0), C.byref(bt)))
                for _ in range(2): f()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                it = 5 if cnt > 1 else 20
                with torch.cuda.stream(stream):
                    e0.record(stream)
                    for _ in range(it): f()
                    e1.record(stream)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / it
                out.append(f"cnt={cnt} ct={ct}: {byts * cnt / ms / 1e6:7.0f} GB/s ({byts * cnt / ms / 1e6 / peak:.2f})")
        print(nm, (log_n, rows, cols_out * size), " | ".join(out))
        del m, a, r, pm
        pb.hal.pool_trim()
