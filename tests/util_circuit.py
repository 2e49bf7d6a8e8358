"""The GLWE-operations interface of poulpy_b200/circuit.py (DeviceGlweOps) over the oracle, shared by the CPU pins and the GPU parity tests."""
import numpy as np

from oracle import pyoracle as O


class _Ct:
    def __init__(self, arr):
        self.arr = arr


class OracleGlweOps:
    """The interface of circuit.DeviceGlweOps over the oracle (numpy batches of (size, cols, n) GLWEs): the packing / post-processing
    sequencing of poulpy_b200/circuit.py is data independent, so the same functions drive both sides."""

    def __init__(self, o, atk, base2k, cols, size, batch, n, dsize=1):
        self.o, self.atk, self.k, self.cols, self.size, self.batch, self.n, self.dsize = o, atk, base2k, cols, size, batch, n, dsize
        self.log_n = n.bit_length() - 1

    def new(self):
        return _Ct(np.zeros((self.batch, self.size, self.cols, self.n), dtype=np.int64))

    def copy(self, dst, src):
        dst.arr[...] = src.arr

    def rotate(self, k, dst, src):
        for b in range(self.batch):
            for c in range(self.cols):
                O.vec_znx_rotate(k, dst.arr[b], c, src.arr[b], c)

    def rotate_assign(self, k, ct):
        tmp = self.new()
        self.rotate(k, tmp, ct)
        ct.arr = tmp.arr

    def add_assign(self, res, a):
        res.arr = res.arr + a.arr  # wrapping int64

    def sub_assign(self, res, a):
        res.arr = res.arr - a.arr

    def sub(self, res, a, b):
        res.arr = a.arr - b.arr

    def rsh1(self, ct):
        for b in range(self.batch):
            for c in range(self.cols):
                O.vec_znx_rsh_assign(self.k, 1, ct.arr[b], c)

    def normalize_assign(self, ct):
        for b in range(self.batch):
            for c in range(self.cols):
                O.vec_znx_normalize_assign(self.k, ct.arr[b], c)

    def automorphism_assign(self, ct, i):
        p = O.trace_galois_element(i, self.n)
        out = np.zeros_like(ct.arr)
        for b in range(self.batch):
            self.o.glwe_automorphism(out[b], self.k, ct.arr[b], self.k, self.atk[i], self.k, p, self.dsize)
        ct.arr = out

    def automorphism_add_assign(self, ct, i):
        for b in range(self.batch):
            self.o.glwe_automorphism_op(0, ct.arr[b], self.k, ct.arr[b], self.atk[i], self.k, O.trace_galois_element(i, self.n), self.dsize)

    def automorphism_sub_negate(self, res, a, i):
        for b in range(self.batch):
            self.o.glwe_automorphism_op(2, res.arr[b], self.k, a.arr[b], self.atk[i], self.k, O.trace_galois_element(i, self.n), self.dsize)

    def trace(self, res, skip, a):
        res.arr[...] = a.arr
        for b in range(self.batch):
            self.o.glwe_trace_assign(res.arr[b], self.k, skip, self.atk, self.k, self.dsize)
