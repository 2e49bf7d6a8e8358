"""CPU pin of the packing sequencing (poulpy_b200/circuit.py: glwe_pack / pack_internal, restating poulpy-core/src/glwe_packing.rs:15-170)
driven over the oracle: with a zero mask every key-switch term vanishes, so the result is plain integer arithmetic on the bodies and the
meaning of the operation can be checked exactly -- coefficient i * 2^g of the output carries the constant coefficient of input i, every
coefficient that is not a multiple of 2^g is cleared by the final partial trace."""
from fractions import Fraction

import numpy as np

from oracle import pyoracle as O
from poulpy_b200 import circuit
from util import fill_uniform
from util_circuit import OracleGlweOps, _Ct


def _torus(limbs, K):
    return sum(Fraction(int(v), 1 << ((j + 1) * K)) for j, v in enumerate(limbs))


def test_glwe_pack_places_constant_coefficients():
    n, K, size, log_n = 64, 16, 4, 6
    rng = np.random.default_rng(11)
    for fl in (O.NTT120, O.FFT64):
        o = O.OracleModule(n, fl)
        atk = []
        for _ in range(log_n):
            pm = o.vmp_pmat_alloc(size, 1, 2, size + 1)
            o.vmp_prepare(pm, fill_uniform(rng, (size, 1, size + 1, 2, n), K))
            atk.append(pm)
        ops = OracleGlweOps(o, atk, K, 2, size, 1, n)
        for g in (3, 4):
            cts, consts = {}, {}
            for i in range(n >> g):
                a = np.zeros((1, size, 2, n), dtype=np.int64)
                a[0, :, 0] = fill_uniform(rng, (size, n), K)  # body only: zero mask
                cts[i << g] = _Ct(a)
                consts[i << g] = _torus(a[0, :, 0, 0], K)
            res = ops.new()
            circuit.glwe_pack(ops, res, cts, g)
            assert not res.arr[0, :, 1].any()
            tol = Fraction(4 * log_n, 1 << (size * K))
            for c in range(n):
                got = _torus(res.arr[0, :, 0, c], K)
                err = got - (consts[c] if c in consts else 0)
                err -= round(err)
                assert abs(err) <= tol, (fl, g, c, float(err))
