"""Phase timing of circuit bootstrapping (constant mode) at the reference bench shape (poulpy-bench circuit_bootstrapping.rs:47-129):
n=1024, n_lwe=574, block 7, rank 2, base2k 13, keys k=52 dnum=3, result k=26 dnum=2.  CBT_B = batch, CBT_FL = fft64|ntt120."""
import ctypes as C
import os
import time

import numpy as np

import poulpy_b200 as pb
from poulpy_b200 import circuit, hal

fl = pb.NTT120 if os.environ.get("CBT_FL") == "ntt120" else pb.FFT64
B = int(os.environ.get("CBT_B", "128"))
n, log_n, n_lwe, block, rank, K = 1024, 10, 574, 7, 2, 13
cols, ksz, kd, res_size, dnum_res = rank + 1, 4, 3, 2, 2
rng = np.random.default_rng(1)
m = pb.Module(n, fl)
lib = pb.lib()
per = n * cols * cols * kd * ksz * m.prep_bytes
brk_buf = pb.DevBuf(per * n_lwe)
one = hal.VmpPMat(brk_buf, n, kd, cols, cols, ksz)
m.vmp_prepare(one, m.mat_znx_from_numpy(rng.integers(-(1 << 12), 1 << 12, size=(kd, cols, ksz, cols, n), dtype=np.int64)))
for i in range(1, n_lwe):
    lib.pgb_memcpy_d2d(C.c_void_p(brk_buf.ptr + i * per), C.c_void_p(brk_buf.ptr), C.c_size_t(per))


def mk(count):
    out = []
    for _ in range(count):
        pm = m.vmp_pmat_alloc(kd, rank, cols, ksz)
        m.vmp_prepare(pm, m.mat_znx_from_numpy(rng.integers(-(1 << 12), 1 << 12, size=(kd, rank, ksz, cols, n), dtype=np.int64)))
        out.append(pm)
    return out


atk, tsk = mk(log_n), mk(rank)
lwe = rng.integers(-(1 << 12), 1 << 12, size=(B, 1, 1, n_lwe + 1), dtype=np.int64)
lwe_dev = pb.DevBuf(lwe.nbytes)
lwe_dev.upload(lwe)
xpa = m.cggi_x_pow_a()


def timed(name, fn, reps=2):
    fn()
    m.sync()
    t0 = time.perf_counter()
    l0 = m.launch_count
    for _ in range(reps):
        fn()
    m.sync()
    dt = (time.perf_counter() - t0) / reps
    print(f"{name:28s} {dt * 1e3:9.3f} ms  {B / dt:10.1f} /s  launches {(m.launch_count - l0) // reps}", flush=True)


timed("cbt_to_constant", lambda: circuit.circuit_bootstrap_to_constant(m, lwe_dev, B, n_lwe, 1, K, one, xpa, block, atk, tsk, K, rank, dnum_res, res_size, 1))
lut, drift = circuit.lookup_table_set(m, [0, 1 << K, 1, 2], K * dnum_res, K, 1)
acc = m.vec_znx_alloc(cols, ksz, B)
lwe_2n = m.cggi_mod_switch_2n(lwe_dev, B, n_lwe, 1, K, 2 * n, rot_left=True)
sc = [None, None, None]


def br():
    sc[0] = m.cggi_blind_rotate(acc, lwe_2n, n_lwe, lut, one, xpa, block, K, sc[0])


timed("blind_rotate", br)
tmp = m.vec_znx_alloc(cols, ksz, B)


def tr():
    sc[1] = m.glwe_trace_assign(tmp, K, 0, atk, K, 1, sc[1])


timed("trace", tr)
ggsw = pb.DevBuf(B * n * dnum_res * cols * cols * res_size * 8)


def ex():
    sc[2] = m.ggsw_expand_row(ggsw, B, dnum_res, rank, res_size, K, tsk, K, 1, sc[2])


timed("ggsw_expand_row", ex)

if os.environ.get("CBT_PROFILE"):
    import cProfile
    import pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(2):
        circuit.circuit_bootstrap_to_constant(m, lwe_dev, B, n_lwe, 1, K, one, xpa, block, atk, tsk, K, rank, dnum_res, res_size, 1)
    m.sync()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(25)

# per-category split of the blind rotation (the module's event profiler around every launch)
NCAT = 16
lib.pgb_profile_category_name.restype = C.c_char_p
lib.pgb_profile_enable(m._h, 1)
br()
m.sync()
pm_, pn_ = (C.c_double * NCAT)(), (C.c_uint64 * NCAT)()
lib.pgb_profile_read(m._h, pm_, pn_, 1)
lib.pgb_profile_enable(m._h, 0)
for i in range(NCAT):
    if pn_[i]:
        print(f"  {lib.pgb_profile_category_name(i).decode():16s} {pm_[i]:9.3f} ms  {pn_[i]:5d} launches")
