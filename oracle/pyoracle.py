"""ctypes binding of the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py.  Nothing under poulpy_b200/ imports this module.

The class `OracleModule` mirrors the method names of the reference's `Module<B>` HAL API
(poulpy-hal/src/api/*.rs) so that parity tests read like poulpy-hal/src/test_suite/*.rs.

Containers are numpy arrays in the reference's limb-major / column-minor layout:
  VecZnx              int64   (size, cols, n)
  VecZnxDft  NTT120   uint64  (size, cols, n, 4)   q120b, lazy residues (compare modulo Q[k])
             FFT64    float64 (size, cols, n)      [re(m) | im(m)] per limb
  VecZnxBig  NTT120   uint64  (size, cols, n, 2)   little-endian i128 (lo, hi)
             FFT64    int64   (size, cols, n)
  SvpPPol             like one VecZnxDft limb per column: (cols, n, 4) / (cols, n)
  VmpPMat             opaque flat buffer of n*rows*cols_in*cols_out*size ScalarPrep
  MatZnx              int64   (rows, cols_in, size, cols_out, n)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

Q = (1073479681, 1071513601, 1070727169, 1068236801)
OMEGA = (1070907127, 315046632, 309185662, 846468380)
CRT_CST = (43599465, 292938863, 594011630, 140177212)
NTT120, FFT64 = 0, 1


def build(force: bool = False) -> str:
    """Compile oracle/*.c into liboracle.so (gcc only; the reference is Rust and cannot be built here)."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h", ".inc")) or f == "Makefile"]
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _VZ(C.Structure):
    _fields_ = [("data", C.c_void_p), ("n", C.c_size_t), ("cols", C.c_size_t), ("size", C.c_size_t)]


class _PP(C.Structure):
    _fields_ = [("data", C.c_void_p), ("n", C.c_size_t), ("cols", C.c_size_t)]


class _PM(C.Structure):
    _fields_ = [("data", C.c_void_p), ("n", C.c_size_t), ("rows", C.c_size_t), ("cols_in", C.c_size_t),
                ("cols_out", C.c_size_t), ("size", C.c_size_t)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_ntt120_new.restype = C.c_void_p
        _lib.orc_fft64_new.restype = C.c_void_p
        _lib.orc_ntt120_new.argtypes = [C.c_size_t]
        _lib.orc_fft64_new.argtypes = [C.c_size_t]
        _lib.orc_ntt120_free.argtypes = [C.c_void_p]
        _lib.orc_fft64_free.argtypes = [C.c_void_p]
        _lib.orc_fft64_omg.restype = C.POINTER(C.c_double)
        _lib.orc_ntt120_reduc_h.restype = C.c_uint64
        _lib.orc_ntt120_bbc_h.restype = C.c_uint64
        _lib.orc_ntt120_fwd_levels.restype = C.c_size_t
        _lib.orc_ntt120_inv_levels.restype = C.c_size_t
    return _lib


def _p(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


def _sz(x):
    return C.c_size_t(int(x))


def _vz(a: np.ndarray, size=None):
    """(size, cols, n[, k]) array -> struct; `size` may shrink the view (set_size)."""
    s = a.shape[0] if size is None else size
    return _VZ(a.ctypes.data, a.shape[2], a.shape[1], s)


def _pp(a: np.ndarray):
    return _PP(a.ctypes.data, a.shape[1], a.shape[0])


class VmpPMat:
    def __init__(self, n, rows, cols_in, cols_out, size, prep_bytes):
        self.n, self.rows, self.cols_in, self.cols_out, self.size = n, rows, cols_in, cols_out, size
        self.data = np.zeros(n * rows * cols_in * cols_out * size * prep_bytes, dtype=np.uint8)

    def struct(self):
        return _PM(self.data.ctypes.data, self.n, self.rows, self.cols_in, self.cols_out, self.size)


def i128_to_int(big: np.ndarray) -> np.ndarray:
    """(…, 2) uint64 (lo, hi) -> object array of Python ints (signed)."""
    lo = big[..., 0].astype(object)
    hi = big[..., 1].astype(np.int64).astype(object)
    return hi * (1 << 64) + lo


def int_to_i128(vals) -> np.ndarray:
    vals = np.asarray(vals, dtype=object)
    out = np.zeros(vals.shape + (2,), dtype=np.uint64)
    flat = vals.reshape(-1)
    o = out.reshape(-1, 2)
    for i, v in enumerate(flat):
        v = int(v) & ((1 << 128) - 1)
        o[i, 0] = v & 0xFFFFFFFFFFFFFFFF
        o[i, 1] = v >> 64
    return out


class OracleModule:
    """Mirror of `Module<NTT120Ref>` / `Module<FFT64Ref>` restricted to the hot path."""

    def __init__(self, n: int, flavour: int):
        self.n, self.flavour = n, flavour
        L = lib()
        self._h = C.c_void_p(L.orc_ntt120_new(n) if flavour == NTT120 else L.orc_fft64_new(n))
        self._pfx = "orc_ntt120_" if flavour == NTT120 else "orc_fft64_"

    def __del__(self):
        try:
            (lib().orc_ntt120_free if self.flavour == NTT120 else lib().orc_fft64_free)(self._h)
        except Exception:
            pass

    # --- sizes / allocation (poulpy-hal/src/layouts/module.rs:44-70) -------------------------------
    @property
    def prep_bytes(self):
        return 32 if self.flavour == NTT120 else 8

    def vec_znx_alloc(self, cols, size):
        return np.zeros((size, cols, self.n), dtype=np.int64)

    def vec_znx_dft_alloc(self, cols, size):
        if self.flavour == NTT120:
            return np.zeros((size, cols, self.n, 4), dtype=np.uint64)
        return np.zeros((size, cols, self.n), dtype=np.float64)

    def vec_znx_big_alloc(self, cols, size):
        if self.flavour == NTT120:
            return np.zeros((size, cols, self.n, 2), dtype=np.uint64)
        return np.zeros((size, cols, self.n), dtype=np.int64)

    def svp_ppol_alloc(self, cols):
        if self.flavour == NTT120:
            return np.zeros((cols, self.n, 4), dtype=np.uint64)
        return np.zeros((cols, self.n), dtype=np.float64)

    def vmp_pmat_alloc(self, rows, cols_in, cols_out, size):
        return VmpPMat(self.n, rows, cols_in, cols_out, size, self.prep_bytes)

    def _f(self, name):
        return getattr(lib(), self._pfx + name)

    # --- vec_znx_dft (poulpy-hal/src/api/vec_znx_dft.rs) -------------------------------------------
    def vec_znx_dft_apply(self, step, offset, res, res_col, a, a_col, res_size=None, a_size=None):
        r, av = _vz(res, res_size), _vz(a, a_size)
        self._f("vec_znx_dft_apply")(self._h, _sz(step), _sz(offset), C.byref(r), _sz(res_col), C.byref(av), _sz(a_col))

    def vec_znx_idft_apply(self, res, res_col, a, a_col, res_size=None, a_size=None):
        r, av = _vz(res, res_size), _vz(a, a_size)
        self._f("vec_znx_idft_apply")(self._h, C.byref(r), _sz(res_col), C.byref(av), _sz(a_col))

    def vec_znx_idft_apply_tmpa(self, res, res_col, a, a_col):
        r, av = _vz(res), _vz(a)
        self._f("vec_znx_idft_apply_tmpa")(self._h, C.byref(r), _sz(res_col), C.byref(av), _sz(a_col))

    def vec_znx_idft_apply_consume(self, a):
        """Consumes `a` in place and returns the VecZnxBig view of the same memory."""
        av = _vz(a)
        self._f("vec_znx_idft_apply_consume")(self._h, C.byref(av))
        size, cols, n = a.shape[:3]
        if self.flavour == NTT120:
            flat = a.reshape(-1)[: size * cols * n * 2]
            return flat.reshape(size, cols, n, 2)
        return a.view(np.int64)

    def _dft2(self, name, res, res_col, a, a_col, res_size=None, a_size=None):
        r, av = _vz(res, res_size), _vz(a, a_size)
        self._f(name)(C.byref(r), _sz(res_col), C.byref(av), _sz(a_col))

    def _dft3(self, name, res, res_col, a, a_col, b, b_col):
        r, av, bv = _vz(res), _vz(a), _vz(b)
        self._f(name)(C.byref(r), _sz(res_col), C.byref(av), _sz(a_col), C.byref(bv), _sz(b_col))

    def vec_znx_dft_add_into(self, res, res_col, a, a_col, b, b_col):
        self._dft3("vec_znx_dft_add_into", res, res_col, a, a_col, b, b_col)

    def vec_znx_dft_sub(self, res, res_col, a, a_col, b, b_col):
        self._dft3("vec_znx_dft_sub", res, res_col, a, a_col, b, b_col)

    def vec_znx_dft_add_assign(self, res, res_col, a, a_col, **kw):
        self._dft2("vec_znx_dft_add_assign", res, res_col, a, a_col, **kw)

    def vec_znx_dft_sub_assign(self, res, res_col, a, a_col, **kw):
        self._dft2("vec_znx_dft_sub_assign", res, res_col, a, a_col, **kw)

    def vec_znx_dft_sub_negate_assign(self, res, res_col, a, a_col, **kw):
        self._dft2("vec_znx_dft_sub_negate_assign", res, res_col, a, a_col, **kw)

    def vec_znx_dft_copy(self, step, offset, res, res_col, a, a_col, res_size=None, a_size=None):
        r, av = _vz(res, res_size), _vz(a, a_size)
        self._f("vec_znx_dft_copy")(_sz(step), _sz(offset), C.byref(r), _sz(res_col), C.byref(av), _sz(a_col))

    def vec_znx_dft_zero(self, res, res_col):
        r = _vz(res)
        self._f("vec_znx_dft_zero")(C.byref(r), _sz(res_col))

    # --- svp (poulpy-hal/src/api/svp_ppol.rs) ------------------------------------------------------
    def svp_prepare(self, res, res_col, a, a_col):
        """a: ScalarZnx int64 (cols, n)."""
        r = _pp(res)
        s = _PP(a.ctypes.data, a.shape[1], a.shape[0])
        self._f("svp_prepare")(self._h, C.byref(r), _sz(res_col), C.byref(s), _sz(a_col))

    def svp_apply_dft_to_dft(self, res, res_col, a, a_col, b, b_col):
        r, pv, bv = _vz(res), _pp(a), _vz(b)
        self._f("svp_apply_dft_to_dft")(self._h, C.byref(r), _sz(res_col), C.byref(pv), _sz(a_col), C.byref(bv), _sz(b_col))

    def svp_apply_dft_to_dft_assign(self, res, res_col, a, a_col):
        r, pv = _vz(res), _pp(a)
        self._f("svp_apply_dft_to_dft_assign")(self._h, C.byref(r), _sz(res_col), C.byref(pv), _sz(a_col))

    # --- vmp (poulpy-hal/src/api/vmp_pmat.rs) ------------------------------------------------------
    def vmp_prepare(self, pmat: VmpPMat, mat: np.ndarray):
        rows, cols_in, size, cols_out, n = mat.shape
        assert (rows, cols_in, cols_out, size, n) == (pmat.rows, pmat.cols_in, pmat.cols_out, pmat.size, pmat.n)
        ms = _PM(mat.ctypes.data, n, rows, cols_in, cols_out, size)
        ps = pmat.struct()
        self._f("vmp_prepare")(self._h, C.byref(ps), C.byref(ms))

    def vmp_apply_dft_to_dft(self, res, a, pmat: VmpPMat, limb_offset=0, res_size=None, a_size=None):
        r, av, ps = _vz(res, res_size), _vz(a, a_size), pmat.struct()
        self._f("vmp_apply_dft_to_dft")(self._h, C.byref(r), C.byref(av), C.byref(ps), _sz(limb_offset))

    # --- vec_znx_big (poulpy-hal/src/api/vec_znx_big.rs) -------------------------------------------
    def vec_znx_big_add_small_assign(self, res, res_col, a, a_col):
        self._dft2("vec_znx_big_add_small_assign", res, res_col, a, a_col)

    def vec_znx_big_normalize(self, res, res_base2k, res_offset, res_col, a, a_base2k, a_col, op=0):
        r, av = _vz(res), _vz(a)
        self._f("vec_znx_big_normalize")(C.byref(r), _sz(res_base2k), C.c_int64(res_offset), _sz(res_col), C.byref(av),
                                         _sz(a_base2k), _sz(a_col), C.c_int(op))

    def vec_znx_big_normalize_add_assign(self, *a):
        self.vec_znx_big_normalize(*a, op=1)

    def vec_znx_big_normalize_sub_assign(self, *a):
        self.vec_znx_big_normalize(*a, op=-1)

    # --- compositions --------------------------------------------------------------------------------
    def glwe_keyswitch(self, res, res_base2k, a, a_base2k, key: VmpPMat, key_base2k, dsize=1):
        r, av, ks = _vz(res), _vz(a), key.struct()
        lib().orc_glwe_keyswitch(C.c_int(self.flavour), self._h, C.byref(r), _sz(res_base2k), C.byref(av), _sz(a_base2k),
                                 C.byref(ks), _sz(key_base2k), _sz(dsize))

    def glwe_external_product(self, res, res_base2k, a, a_base2k, ggsw: VmpPMat, ggsw_base2k, dsize=1):
        r, av, ks = _vz(res), _vz(a), ggsw.struct()
        lib().orc_glwe_external_product(C.c_int(self.flavour), self._h, C.byref(r), _sz(res_base2k), C.byref(av),
                                        _sz(a_base2k), C.byref(ks), _sz(ggsw_base2k), _sz(dsize))

    def glwe_keyswitch_batch(self, res, res_base2k, a, a_base2k, key: VmpPMat, key_base2k, dsize=1, threads=0):
        """res/a: (batch, size, cols, n) int64."""
        B, a_size, a_cols, n = a.shape
        ks = key.struct()
        lib().orc_glwe_keyswitch_batch(C.c_int(self.flavour), self._h, _p(res), _sz(res.shape[1]), _sz(res_base2k), _p(a),
                                       _sz(a_size), _sz(a_base2k), _sz(n), _sz(a_cols - 1), _sz(res.shape[2] - 1),
                                       C.byref(ks), _sz(key_base2k), _sz(dsize), _sz(B), C.c_int(threads))

    def glwe_external_product_batch(self, res, res_base2k, a, a_base2k, ggsw: VmpPMat, ggsw_base2k, dsize=1, threads=0):
        B, a_size, a_cols, n = a.shape
        ks = ggsw.struct()
        lib().orc_glwe_external_product_batch(C.c_int(self.flavour), self._h, _p(res), _sz(res.shape[1]), _sz(res_base2k),
                                              _p(a), _sz(a_size), _sz(a_base2k), _sz(n), _sz(a_cols - 1), C.byref(ks),
                                              _sz(ggsw_base2k), _sz(dsize), _sz(B), C.c_int(threads))

    # --- bivariate convolution (poulpy-hal/src/api/convolution.rs; CnvPVecL/R kept in the VecZnxDft layout) ---------------------
    def cnv_prepare(self, res, a, mask=-1):
        """cnv_prepare_left / cnv_prepare_right / cnv_prepare_self: res = VecZnxDft-shaped array (size, cols, n[, 4])."""
        r, av = _vz(res), _vz(a)
        self._f("cnv_prepare")(self._h, C.byref(r), C.byref(av), C.c_int64(mask))

    def cnv_apply_dft(self, cnv_offset, res, res_col, a, a_col, b, b_col):
        r, av, bv = _vz(res), _vz(a), _vz(b)
        if self.flavour == NTT120:
            lib().orc_ntt120_cnv_apply_dft(self._h, _sz(cnv_offset), C.byref(r), _sz(res_col), C.byref(av), _sz(a_col), C.byref(bv), _sz(b_col))
        else:
            lib().orc_fft64_cnv_apply_dft(_sz(cnv_offset), C.byref(r), _sz(res_col), C.byref(av), _sz(a_col), C.byref(bv), _sz(b_col))

    def cnv_pairwise_apply_dft(self, cnv_offset, res, res_col, a, b, col_i, col_j):
        r, av, bv = _vz(res), _vz(a), _vz(b)
        if self.flavour == NTT120:
            lib().orc_ntt120_cnv_pairwise_apply_dft(self._h, _sz(cnv_offset), C.byref(r), _sz(res_col), C.byref(av), C.byref(bv), _sz(col_i), _sz(col_j))
        else:
            lib().orc_fft64_cnv_pairwise_apply_dft(_sz(cnv_offset), C.byref(r), _sz(res_col), C.byref(av), C.byref(bv), _sz(col_i), _sz(col_j))

    def cnv_by_const_apply(self, cnv_offset, res, res_col, a, a_col, b):
        r, av = _vz(res), _vz(a)
        b = np.ascontiguousarray(b, dtype=np.int64)
        self._f("cnv_by_const_apply")(_sz(cnv_offset), C.byref(r), _sz(res_col), C.byref(av), _sz(a_col), _p(b), _sz(len(b)))

    def glwe_tensor_apply(self, cnv_offset, res, res_base2k, a, a_effective_k, b, b_effective_k, ab_base2k):
        r, av, bv = _vz(res), _vz(a), _vz(b)
        lib().orc_glwe_tensor_apply(C.c_int(self.flavour), self._h, _sz(cnv_offset), C.byref(r), _sz(res_base2k), C.byref(av),
                                    _sz(a_effective_k), C.byref(bv), _sz(b_effective_k), _sz(ab_base2k))

    def glwe_tensor_relinearize(self, res, res_base2k, a, a_base2k, tsk: VmpPMat, key_base2k, dsize=1):
        r, av, ks = _vz(res), _vz(a), tsk.struct()
        lib().orc_glwe_tensor_relinearize(C.c_int(self.flavour), self._h, C.byref(r), _sz(res_base2k), C.byref(av), _sz(a_base2k),
                                          C.byref(ks), _sz(key_base2k), _sz(dsize))

    def glwe_automorphism(self, res, res_base2k, a, a_base2k, key: VmpPMat, key_base2k, p, dsize=1):
        r, av, ks = _vz(res), _vz(a), key.struct()
        lib().orc_glwe_automorphism(C.c_int(self.flavour), self._h, C.byref(r), _sz(res_base2k), C.byref(av), _sz(a_base2k), C.byref(ks),
                                    _sz(key_base2k), C.c_int64(p), _sz(dsize))

    def glwe_automorphism_add_assign(self, res, res_base2k, key: VmpPMat, key_base2k, p, dsize=1):
        """res (size, cols, n) += automorphism_p(key-switch(res)) -- automorphism/glwe_ct.rs:142-183."""
        r, ks = _vz(res), key.struct()
        lib().orc_glwe_automorphism_add_assign(C.c_int(self.flavour), self._h, C.byref(r), _sz(res_base2k), C.byref(ks), _sz(key_base2k),
                                               C.c_int64(p), _sz(dsize))

    def glwe_automorphism_op(self, op, res, res_base2k, a, key: VmpPMat, key_base2k, p, dsize=1):
        """op 0: res = aut(ks(a)) + a, 1: aut(ks(a)) - a, 2: a - aut(ks(a)) (automorphism/glwe_ct.rs:95-275); res may be a."""
        r, av, ks = _vz(res), _vz(a), key.struct()
        lib().orc_glwe_automorphism_op(C.c_int(self.flavour), self._h, C.c_int(op), C.byref(r), _sz(res_base2k), C.byref(av), C.byref(ks),
                                       _sz(key_base2k), C.c_int64(p), _sz(dsize))

    def glwe_trace_assign(self, res, res_base2k, skip, keys, key_base2k, dsize=1):
        """keys: list of log_n prepared automorphism keys, keys[i] for trace_galois_element(i, n) (glwe_trace.rs:129-175)."""
        arr = (C.POINTER(_PM) * len(keys))(*[C.pointer(k.struct()) for k in keys])
        r = _vz(res)
        lib().orc_glwe_trace_assign(C.c_int(self.flavour), self._h, C.byref(r), _sz(res_base2k), _sz(skip), arr, _sz(key_base2k), _sz(dsize))

    def ggsw_expand_row(self, ggsw, res_base2k, tsk, tsk_base2k, dsize=1):
        """ggsw: int64 (dnum, rank+1, size, rank+1, n) with the column-0 GLWEs filled; tsk: list of rank prepared keys
        (conversion/gglwe_to_ggsw.rs:116-268)."""
        dnum, cols, size, cols2, n = ggsw.shape
        assert cols == cols2 and len(tsk) == cols - 1
        arr = (C.POINTER(_PM) * len(tsk))(*[C.pointer(k.struct()) for k in tsk])
        lib().orc_ggsw_expand_row(C.c_int(self.flavour), self._h, _p(ggsw), _sz(n), _sz(dnum), _sz(cols - 1), _sz(size), _sz(res_base2k), arr,
                                  _sz(tsk_base2k), _sz(dsize))

    def cggi_x_pow_a(self):
        res = self.svp_ppol_alloc(2 * self.n)
        r = _pp(res)
        lib().orc_cggi_x_pow_a(C.c_int(self.flavour), self._h, C.byref(r))
        return res

    def cggi_blind_rotate_block_binary(self, res, lwe_2n, lut, brk, x_pow_a, block_size, base2k):
        """brk: list of VmpPMat (one per LWE coefficient)."""
        n_lwe = len(brk)
        arr = (_PM * n_lwe)(*[b.struct() for b in brk])
        r, lv, xp = _vz(res), _vz(lut), _pp(x_pow_a)
        lwe_2n = np.ascontiguousarray(lwe_2n, dtype=np.int64)
        lib().orc_cggi_blind_rotate_block_binary(C.c_int(self.flavour), self._h, C.byref(r), _p(lwe_2n), _sz(n_lwe),
                                                 C.byref(lv), arr, C.byref(xp), _sz(block_size), _sz(base2k))

    def cggi_blind_rotate_block_binary_batch(self, res, lwe_2n, lut, brk, x_pow_a, block_size, base2k, threads=0):
        """res: (batch, size, cols, n) int64; lwe_2n: (batch, n_lwe + 1); ciphertexts spread over `threads` host threads (0 = all)."""
        n_lwe = len(brk)
        arr = (_PM * n_lwe)(*[b.struct() for b in brk])
        lv, xp = _vz(lut), _pp(x_pow_a)
        lwe_2n = np.ascontiguousarray(lwe_2n, dtype=np.int64)
        batch, size, cols, n = res.shape
        assert res.flags["C_CONTIGUOUS"] and lwe_2n.shape == (batch, n_lwe + 1)
        lib().orc_cggi_blind_rotate_block_binary_batch(C.c_int(self.flavour), self._h, _p(res), _sz(n), _sz(cols), _sz(size), _p(lwe_2n),
                                                       _sz(n_lwe), C.byref(lv), arr, C.byref(xp), _sz(block_size), _sz(base2k), _sz(batch),
                                                       C.c_int(threads))

    def cggi_blind_rotate_block_binary_extended(self, res, lwe_2n, luts, brk, x_pow_a, block_size, base2k):
        """execute_block_binary_extended (algorithm.rs:121-273); luts: list of extension_factor arrays (size, 1, n) = LookupTable.data,
        lwe_2n mod-switched to 2 * n * extension_factor."""
        n_lwe, ext = len(brk), len(luts)
        arr = (_PM * n_lwe)(*[b.struct() for b in brk])
        larr = (_VZ * ext)(*[_vz(l) for l in luts])
        r, xp = _vz(res), _pp(x_pow_a)
        lwe_2n = np.ascontiguousarray(lwe_2n, dtype=np.int64)
        lib().orc_cggi_blind_rotate_block_binary_extended(C.c_int(self.flavour), self._h, C.byref(r), _p(lwe_2n), _sz(n_lwe), larr, _sz(ext),
                                                          arr, C.byref(xp), _sz(block_size), _sz(base2k))

    def cggi_blind_rotate_standard(self, res, res_base2k, lwe_2n, lut, brk, brk_base2k):
        """execute_standard (block_size == 1); brk: list of VmpPMat (one GGSW per LWE coefficient)."""
        n_lwe = len(brk)
        arr = (_PM * n_lwe)(*[b.struct() for b in brk])
        r, lv = _vz(res), _vz(lut)
        lwe_2n = np.ascontiguousarray(lwe_2n, dtype=np.int64)
        lib().orc_cggi_blind_rotate_standard(C.c_int(self.flavour), self._h, C.byref(r), _sz(res_base2k), _p(lwe_2n), _sz(n_lwe),
                                             C.byref(lv), arr, _sz(brk_base2k))


# --- free functions --------------------------------------------------------------------------------
def vec_znx_normalize(res, res_base2k, res_offset, res_col, a, a_base2k, a_col):
    r, av = _vz(res), _vz(a)
    lib().orc_vec_znx_normalize(C.byref(r), _sz(res_base2k), C.c_int64(res_offset), _sz(res_col), C.byref(av),
                                _sz(a_base2k), _sz(a_col), C.c_int(0))


def vec_znx_rotate(p, res, res_col, a, a_col):
    r, av = _vz(res), _vz(a)
    lib().orc_vec_znx_rotate(C.c_int64(p), C.byref(r), _sz(res_col), C.byref(av), _sz(a_col))


def vec_znx_automorphism(p, res, res_col, a, a_col):
    r, av = _vz(res), _vz(a)
    lib().orc_vec_znx_automorphism(C.c_int64(p), C.byref(r), _sz(res_col), C.byref(av), _sz(a_col))


def vec_znx_normalize_assign(base2k, res, res_col):
    r = _vz(res)
    lib().orc_vec_znx_normalize_assign(_sz(base2k), C.byref(r), _sz(res_col))


def vec_znx_mul_xp_minus_one_assign(p, res, res_col):
    r = _vz(res)
    lib().orc_vec_znx_mul_xp_minus_one_assign(C.c_int64(p), C.byref(r), _sz(res_col))


def mod_switch_2n(two_n_domain, lwe, base2k, rot_left=True):
    """lwe: VecZnx (size, 1, n_lwe + 1)."""
    res = np.zeros(lwe.shape[2], dtype=np.int64)
    lv = _vz(lwe)
    lib().orc_mod_switch_2n(_sz(two_n_domain), _p(res), C.byref(lv), _sz(base2k), C.c_int(1 if rot_left else 0))
    return res


def ntt120_set_simd(on: bool):
    """Select the AVX2 four-primes-per-__m256i data path (poulpy-cpu-avx style) for the NTT120 butterflies and bbc products."""
    lib().orc_ntt120_set_simd(C.c_int(1 if on else 0))


def num_threads():
    return int(lib().orc_num_threads())


# raw leaf access for the KATs
def ntt120_b_from_znx64(x):
    x = np.ascontiguousarray(x, dtype=np.int64)
    res = np.zeros((len(x), 4), dtype=np.uint64)
    lib().orc_ntt120_b_from_znx64(_sz(len(x)), _p(res), _p(x))
    return res


def ntt120_b_to_znx128(x):
    x = np.ascontiguousarray(x, dtype=np.uint64)
    res = np.zeros((x.shape[0], 2), dtype=np.uint64)
    lib().orc_ntt120_b_to_znx128(_sz(x.shape[0]), _p(res), _p(x))
    return i128_to_int(res)


def vec_znx_rsh_assign(base2k, k, res, res_col):
    r = _vz(res)
    lib().orc_vec_znx_rsh_assign(_sz(base2k), _sz(k), C.byref(r), _sz(res_col))


def trace_galois_element(i, n):
    lib().orc_trace_galois_element.restype = C.c_int64
    return int(lib().orc_trace_galois_element(_sz(i), _sz(n)))
