"""Circuit bootstrapping, constant mode: host orchestration of
`circuit_bootstrap_core(to_exponent = false, ...)` (poulpy-bin-fhe/src/circuit_bootstrapping/circuit.rs:219-380) with
extension_factor = 1 over the device-resident entry points of the C ABI:

    LUT f[j * alpha + i] = j * 2^(res_base2k * (dnum - 1 - i))            circuit.rs:278-299, lut.rs:271-338 (lookup_table_set)
    acc  = blind_rotate(lwe, LUT)                                           pgb_cggi_mod_switch_2n_batched + pgb_cggi_blind_rotate_batched
    for i in 0..dnum:  GGSW.at(i, 0) = trace(acc);  acc *= X^{-gap}         pgb_glwe_trace_assign_batched, pgb_vec_znx_rotate_batched
    GGSW columns 1..rank = ggsw_expand_row(GGSW, tsk)                       pgb_ggsw_expand_row_batched

Everything stays on the device; the host only sequences the launches, as `poulpy-bin-fhe` does above `Module<B>`.  All layouts share one
base2k here (the reference converts between the BRK / ATK / result layouts with glwe_normalize when they differ); the exponent mode
(post_process with glwe_pack) and extension_factor > 1 (extended blind rotation) are not ported.
"""
import ctypes as C

import numpy as np

from . import hal


def lookup_table_set(module: "hal.Module", f, k, base2k):
    """LookupTable::set for extension_factor = 1 (lut.rs:271-338) -> (VecZnx(1 col, ceil(k / base2k) limbs) on the device, drift)."""
    n = module.n
    f = [int(x) for x in f]
    assert 0 < len(f) <= n
    limbs = -(-k // base2k)
    scale = 1 << (base2k - k % base2k) if k % base2k else 1
    step = (n + len(f) // 2) // len(f)  # usize::div_round (lut.rs:243-247)
    last = np.zeros(n, dtype=np.int64)
    for i, fi in enumerate(f):
        last[i * step:(i + 1) * step] = fi * scale
    full = np.zeros((limbs, 1, n), dtype=np.int64)
    full[limbs - 1, 0] = last
    lut = module.vec_znx_from_numpy(full)
    module.vec_znx_normalize_assign(base2k, lut, 0)
    drift = step >> 1
    out = module.vec_znx_alloc(1, limbs)
    module.vec_znx_rotate(-drift, out, 0, lut, 0)  # lookup_table_rotate(-drift): extension_factor = 1 -> one negacyclic rotation
    return out, drift


def _glwe_copy(module, dst_ptr, dst_stride, dst_size, src_ptr, src_stride, src_size, cols, batch):
    """glwe_copy (vec_znx_copy per column): the first min(size) limbs, the rest of dst zeroed; limb-major containers, so that is one
    contiguous run per ciphertext."""
    n, lib = module.n, hal.lib()
    width = min(dst_size, src_size) * cols * n * 8
    hal._check(lib.pgb_memcpy_d2d_strided(module._h, C.c_void_p(dst_ptr), C.c_uint64(dst_stride), C.c_void_p(src_ptr), C.c_uint64(src_stride),
                                          C.c_uint64(width), C.c_uint64(batch)))
    if dst_size > src_size:
        module.sync()
        for b in range(batch):
            hal._check(lib.pgb_memset(C.c_void_p(dst_ptr + b * dst_stride + width), 0, C.c_size_t((dst_size - src_size) * cols * n * 8)))


def circuit_bootstrap_to_constant(module: "hal.Module", lwe_dev: "hal.DevBuf", batch, n_lwe, lwe_size, lwe_base2k, brk: "hal.VmpPMat",
                                  x_pow_a, block_size, atk, tsk, base2k, rank, dnum_res, res_size, log_domain, dsize_atk=1, dsize_tsk=1):
    """-> DevBuf holding `batch` GGSW MatZnx(dnum_res, rank+1, rank+1, res_size) (base2k digits).
    lwe_dev: (batch, lwe_size, 1, n_lwe + 1) int64; brk: the n_lwe prepared GGSWs stored consecutively (VmpPMat of the first);
    atk: log_n prepared automorphism keys, atk[i] for Module.trace_galois_element(i); tsk: rank prepared tensor keys."""
    n, cols = module.n, rank + 1
    assert res_size * 0 == 0 and base2k * (dnum_res - 1) < 64 and log_domain + base2k * max(dnum_res - 1, 0) < 64  # circuit.rs:266-276
    alpha = 1 << (dnum_res - 1).bit_length() if dnum_res > 1 else 1
    f = [0] * ((1 << log_domain) * alpha)
    for j in range(1 << log_domain):
        for i in range(dnum_res):
            f[j * alpha + i] = j * (1 << (base2k * (dnum_res - 1 - i)))
    lut, drift = lookup_table_set(module, f, base2k * dnum_res, base2k)
    # blind rotation over the BRK layout (k = brk.max_k: brk.size limbs)
    lwe_2n = module.cggi_mod_switch_2n(lwe_dev, batch, n_lwe, lwe_size, lwe_base2k, 2 * n, rot_left=True)
    acc_size = brk.size
    acc = module.vec_znx_alloc(cols, acc_size, batch)
    module.cggi_blind_rotate(acc, lwe_2n, n_lwe, lut, brk, x_pow_a, block_size, base2k)
    gap = 2 * drift
    assert gap > 0
    ggsw_stride = n * dnum_res * cols * cols * res_size * 8
    ggsw = hal.DevBuf(batch * ggsw_stride)
    tmp_size = max(acc_size, res_size)
    tmp = module.vec_znx_alloc(cols, tmp_size, batch)
    acc2 = module.vec_znx_alloc(cols, acc_size, batch)
    scratch = None
    for i in range(dnum_res):
        # glwe_trace(res.at(i, 0), 0, acc, atk) (glwe_trace.rs:91-127): copy, trace in place, copy out
        _glwe_copy(module, tmp.buf.ptr, tmp.batch_stride, tmp_size, acc.buf.ptr, acc.batch_stride, acc_size, cols, batch)
        scratch = module.glwe_trace_assign(tmp, base2k, 0, atk, base2k, dsize_atk, scratch)
        row_ptr = ggsw.ptr + (i * cols + 0) * (n * cols * res_size * 8)
        _glwe_copy(module, row_ptr, ggsw_stride, res_size, tmp.buf.ptr, tmp.batch_stride, tmp_size, cols, batch)
        if i + 1 < dnum_res:  # glwe_rotate_assign(-gap)
            r, a = acc2.struct(), acc.struct()
            bt = hal._BT(batch, acc2.batch_stride, acc.batch_stride, 0)
            for c in range(cols):
                hal._check(hal.lib().pgb_vec_znx_rotate_batched(module._h, C.c_int64(-gap), C.byref(r), C.c_uint64(c), C.byref(a), C.c_uint64(c),
                                                                C.byref(bt)))
            acc, acc2 = acc2, acc
    module.ggsw_expand_row(ggsw, batch, dnum_res, rank, res_size, base2k, tsk, base2k, dsize_tsk)
    module.sync()
    return ggsw
