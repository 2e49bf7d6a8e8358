"""CPU baselines (oracle C port of poulpy-cpu-ref, kind = "port") for the rows of BASELINE.md, measured on the host cores of
the box it runs on.  Usage: python scripts/cpu_baselines.py > gpurun_out/cpu_baselines.json"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pyoracle as O  # noqa: E402


def timeit(fn, min_s=1.0):
    fn()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < min_s:
        fn()
        n += 1
    return (time.perf_counter() - t0) / n


def main():
    rng = np.random.default_rng(0)
    threads = O.num_threads()
    out = {"cores": threads, "cpu": open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0].strip(": \t")}
    u = lambda shape, k: rng.integers(-(1 << (k - 1)), 1 << (k - 1), size=shape, dtype=np.int64)
    # M1 / M1f: key-switch n=4096
    for fl, nm in ((O.NTT120, "ntt120"), (O.FFT64, "fft64")):
        m = O.OracleModule(4096, fl)
        pm = m.vmp_pmat_alloc(3, 1, 2, 4)
        m.vmp_prepare(pm, u((3, 1, 4, 2, 4096), 18))
        a1 = u((1, 3, 2, 4096), 18)
        r1 = np.zeros_like(a1)
        t1 = timeit(lambda: m.glwe_keyswitch_batch(r1, 18, a1, 18, pm, 18, 1, threads=1))
        aN = u((16 * threads, 3, 2, 4096), 18)
        rN = np.zeros_like(aN)
        tN = timeit(lambda: m.glwe_keyswitch_batch(rN, 18, aN, 18, pm, 18, 1, threads=threads), 3.0)
        out[f"M1_keyswitch_{nm}_n4096"] = {"per_s_1T": 1 / t1, "per_s_all_cores": aN.shape[0] / tN}
    # M3: external product n=2048
    for fl, nm in ((O.NTT120, "ntt120"), (O.FFT64, "fft64")):
        m = O.OracleModule(2048, fl)
        pm = m.vmp_pmat_alloc(3, 2, 2, 3)
        m.vmp_prepare(pm, u((3, 2, 3, 2, 2048), 18))
        a1 = u((1, 3, 2, 2048), 18)
        r1 = np.zeros_like(a1)
        t1 = timeit(lambda: m.glwe_external_product_batch(r1, 18, a1, 18, pm, 18, 1, threads=1))
        aN = u((32 * threads, 3, 2, 2048), 18)
        rN = np.zeros_like(aN)
        tN = timeit(lambda: m.glwe_external_product_batch(rN, 18, aN, 18, pm, 18, 1, threads=threads), 3.0)
        out[f"M3_external_product_{nm}_n2048"] = {"per_s_1T": 1 / t1, "per_s_all_cores": aN.shape[0] / tN}
    # M2: dft / idft limbs per second, single thread
    for fl, nm in ((O.NTT120, "ntt120"), (O.FFT64, "fft64")):
        for log_n in (10, 12, 14, 16):
            n = 1 << log_n
            m = O.OracleModule(n, fl)
            a = u((8, 1, n), 18)
            d = m.vec_znx_dft_alloc(1, 8)
            b = m.vec_znx_big_alloc(1, 8)
            tf = timeit(lambda: m.vec_znx_dft_apply(1, 0, d, 0, a, 0), 0.5) / 8
            ti = timeit(lambda: m.vec_znx_idft_apply(b, 0, d, 0), 0.5) / 8
            out[f"M2_dft_{nm}_log_n={log_n}"] = {"fwd_limbs_per_s_1T": 1 / tf, "inv_limbs_per_s_1T": 1 / ti}
    # M5: vmp alone (reference layout: 32 B / 8 B per coefficient)
    for (log_n, rows, cols_in, cols_out, size) in ((12, 7, 1, 2, 8), (13, 15, 1, 2, 16)):
        n = 1 << log_n
        m = O.OracleModule(n, O.NTT120)
        pm = m.vmp_pmat_alloc(rows, cols_in, cols_out, size)
        m.vmp_prepare(pm, u((rows, cols_in, size, cols_out, n), 18))
        a = u((rows, cols_in, n), 18)
        d = m.vec_znx_dft_alloc(cols_in, rows)
        m.vec_znx_dft_apply(1, 0, d, 0, a, 0)
        r = m.vec_znx_dft_alloc(cols_out, size)
        t = timeit(lambda: m.vmp_apply_dft_to_dft(r, d, pm, 0), 0.5)
        byts = (rows * cols_in + rows * cols_in * cols_out * size + cols_out * size) * n * 32
        out[f"M5_vmp_ntt120_log_n={log_n}_rows={rows}_size={size}"] = {"ms_1T": t * 1e3, "gbs_reference_layout_1T": byts / t / 1e9}
    # M4: CGGI blind rotation n=512, n_lwe=687, rank 3, block 3 (FFT64 as in the reference's bench), single thread, one bootstrap
    n, n_lwe, rank, k = 512, 687, 3, 18
    m = O.OracleModule(n, O.FFT64)
    cols = rank + 1
    proto = m.vmp_pmat_alloc(1, cols, cols, 2)
    m.vmp_prepare(proto, u((1, cols, 2, cols, n), 18))
    brk = [proto] * n_lwe
    xpa = m.cggi_x_pow_a()
    lut = u((1, 1, n), 17)
    lwe = rng.integers(-n, n, size=n_lwe + 1, dtype=np.int64)
    res = np.zeros((1, cols, n), dtype=np.int64)
    t = timeit(lambda: m.cggi_blind_rotate_block_binary(res, lwe, lut, brk, xpa, 3, k), 2.0)
    out["M4_cggi_fft64_n512_nlwe687"] = {"bootstraps_per_s_1T": 1 / t, "bootstraps_per_s_all_cores_extrapolated": threads / t}
    # M6: CKKS ct x ct multiplication (glwe_tensor_apply + glwe_tensor_relinearize), N = 2^15, base2k = 52, 14 limbs, single thread
    n, k, size = 1 << 15, 52, 14
    m = O.OracleModule(n, O.NTT120)
    tsk = m.vmp_pmat_alloc(size, 1, 2, size + 1)
    m.vmp_prepare(tsk, u((size, 1, size + 1, 2, n), 52))
    a, b = u((size, 2, n), 52), u((size, 2, n), 52)
    tensor, res = np.zeros((size, 3, n), dtype=np.int64), np.zeros((size, 2, n), dtype=np.int64)

    def mul():
        m.glwe_tensor_apply(size * k, tensor, k, a, size * k, b, size * k, k)
        m.glwe_tensor_relinearize(res, k, tensor, k, tsk, k, 1)

    t = timeit(mul, 2.0)
    out["M6_ckks_mul_ntt120_n32768"] = {"mul_per_s_1T": 1 / t, "mul_per_s_all_cores_extrapolated": threads / t}
    # N4: glwe_automorphism and glwe_trace at the key-switch shape (n = 4096, rank 1, 3 limbs, key 3 rows x 4 limbs), single thread
    for fl, nm in ((O.NTT120, "ntt120"), (O.FFT64, "fft64")):
        m = O.OracleModule(4096, fl)
        keys = []
        for _ in range(12):
            pm = m.vmp_pmat_alloc(3, 1, 2, 4)
            m.vmp_prepare(pm, u((3, 1, 4, 2, 4096), 18))
            keys.append(pm)
        a1, r1 = u((3, 2, 4096), 18), np.zeros((3, 2, 4096), dtype=np.int64)
        ta = timeit(lambda: m.glwe_automorphism(r1, 18, a1, 18, keys[0], 18, 5), 1.0)
        tt = timeit(lambda: m.glwe_trace_assign(r1, 18, 0, keys, 18, 1), 1.0)
        out[f"N4_automorphism_trace_{nm}_n4096"] = {"automorphisms_per_s_1T": 1 / ta, "traces_per_s_1T": 1 / tt,
                                                    "all_cores_extrapolated": {"automorphisms_per_s": threads / ta, "traces_per_s": threads / tt}}
    # N4c: the blind rotation of circuit bootstrapping at the reference's bench shape (n = 1024, n_lwe = 574, block 7, rank 2, dnum 3,
    # 4-limb keys) = ~90 % of a circuit bootstrap, FFT64, single thread, one matrix replicated
    n, n_lwe = 1024, 574
    m = O.OracleModule(n, O.FFT64)
    pm = m.vmp_pmat_alloc(3, 3, 3, 4)
    m.vmp_prepare(pm, u((3, 3, 4, 3, n), 13))
    brk = [pm] * n_lwe
    lut, lwe = u((4, 1, n), 13), rng.integers(-n, n, size=n_lwe + 1, dtype=np.int64)
    xo, res = m.cggi_x_pow_a(), np.zeros((4, 3, n), dtype=np.int64)
    t = timeit(lambda: m.cggi_blind_rotate_block_binary(res, lwe, lut, brk, xo, 7, 13), 2.0)
    out["N4c_circuit_bootstrap_blind_rotation_fft64_n1024_nlwe574"] = {"per_s_1T": 1 / t, "per_s_all_cores_extrapolated": threads / t}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
