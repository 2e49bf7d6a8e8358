"""CPU pins of the oracle's trace restatement: vec_znx_rsh_assign (reference/vec_znx/shift.rs:186-243),
glwe_automorphism_add_assign (poulpy-core/src/automorphism/glwe_ct.rs:142-183) and glwe_trace_assign (poulpy-core/src/glwe_trace.rs:
129-175).  The reference's own tests are statistical (decrypt + noise); here the restatement is pinned by exact properties: the torus
value after a shift, the zero-mask case (the key-switch term vanishes, so the result is the plain integer formula), the projection onto
the constant coefficient, and the reference's cross-backend procedure (FFT64 oracle == NTT120 oracle on normalised outputs)."""
from fractions import Fraction

import numpy as np

from oracle import pyoracle as O
from util import fill_uniform


def _torus(v, K, i):
    return sum(Fraction(int(v[j, i]), 1 << ((j + 1) * K)) for j in range(v.shape[0]))


def test_rsh_assign_torus_value():
    """value(rsh_k(a)) == value(a) / 2^k (mod 1) up to the last limb, digits normalised -- for every shift with at most one limb step and
    for whole-limb shifts.  (With two or more limb steps and a partial shift the reference's in-place loop zeroes a limb it has just
    written, shift.rs:235-242; that order is restated verbatim and excluded from this property.)"""
    rng, n = np.random.default_rng(0), 8
    for K in (4, 7, 12, 18, 52):
        for size in (1, 2, 3, 5):
            for k in (0, 1, 3, K - 1, K, 2 * K, 3 * K):
                steps = -(-k // K)
                if steps > size or (steps >= 2 and k % K):
                    continue
                a = fill_uniform(rng, (size, 2, n), K)
                r = a.copy()
                O.vec_znx_rsh_assign(K, k, r, 1)
                assert np.array_equal(r[:, 0], a[:, 0])
                assert np.all(r[:, 1] >= -(1 << (K - 1))) and np.all(r[:, 1] < (1 << (K - 1)))
                for i in range(n):
                    err = _torus(r[:, 1], K, i) - _torus(a[:, 1], K, i) / (1 << k)
                    err -= round(err)
                    assert abs(err) <= Fraction(1, 1 << (size * K)), (K, size, k)


def _keys(o, rng, count, dnum, rank, size, k):
    out = []
    for _ in range(count):
        pm = o.vmp_pmat_alloc(dnum, rank, rank + 1, size)
        o.vmp_prepare(pm, fill_uniform(rng, (dnum, rank, size, rank + 1, o.n), k))
        out.append(pm)
    return out


def _automorphism(p, a):
    n = a.shape[-1]
    out = np.zeros_like(a)
    for i in range(n):
        k = (i * p) % (2 * n)
        if k < n:
            out[..., k] = a[..., i]
        else:
            out[..., k - n] = -a[..., i]
    return out


def test_automorphism_add_assign_zero_mask():
    """With a zero mask the gadget product is exactly zero: res_body <- normalize(aut_p(body) + body), mask stays zero."""
    n, K, size = 64, 14, 3
    rng = np.random.default_rng(1)
    for fl in (O.NTT120, O.FFT64):
        o = O.OracleModule(n, fl)
        (key,) = _keys(o, rng, 1, size, 1, 4, K)
        for p in (-1, 5, 25):
            res = np.zeros((size, 2, n), dtype=np.int64)
            res[:, 0] = fill_uniform(rng, (size, n), K)
            body = res[:, 0].copy()
            o.glwe_automorphism_add_assign(res, K, key, K, p)
            assert not res[:, 1].any()
            tot = body + _automorphism(p, body)  # per limb, then one carry pass from the last limb (vec_znx_big_normalize, equal base2k)
            want, carry = np.zeros_like(tot), np.zeros(n, dtype=np.int64)
            for j in range(size - 1, -1, -1):
                v = tot[j] + carry
                d = ((v + (1 << (K - 1))) % (1 << K)) - (1 << (K - 1))
                carry = (v - d) >> K
                want[j] = d
            assert np.array_equal(res[:, 0], want), (fl, p)


def test_trace_projects_onto_constant_coefficient():
    """Zero mask: log_n rounds of (halve, add the automorphism) leave (1/n) * sum over the Galois orbit = the constant coefficient."""
    n, K, size, log_n = 32, 16, 4, 5
    rng = np.random.default_rng(2)
    assert [O.trace_galois_element(i, n) for i in range(log_n)] == [-1, 5, 25, 625 % 64, pow(5, 8, 64)]
    for fl in (O.NTT120, O.FFT64):
        o = O.OracleModule(n, fl)
        keys = _keys(o, rng, log_n, size, 1, size + 1, K)
        res = np.zeros((size, 2, n), dtype=np.int64)
        res[:, 0] = fill_uniform(rng, (size, n), K)
        before = [_torus(res[:, 0], K, i) for i in range(n)]
        o.glwe_trace_assign(res, K, 0, keys, K)
        tol = Fraction(log_n + 1, 1 << (size * K))
        for i in range(n):
            got = _torus(res[:, 0], K, i)
            err = got - (before[0] if i == 0 else 0)
            err -= round(err)
            assert abs(err) <= tol, (fl, i, float(err))


def test_trace_cross_backend():
    """poulpy-cpu-ref/src/tests.rs:47-141 procedure: the FFT64 and the NTT120 restatements agree exactly on the normalised result."""
    n, K, size, log_n = 64, 12, 3, 6
    rng = np.random.default_rng(3)
    for rank, dsize, res_k, skip in ((1, 1, K, 0), (2, 1, K, 2), (1, 2, K, 1), (1, 1, K - 1, 0)):
        mats = [fill_uniform(rng, (-(-size // dsize), rank, size + 1, rank + 1, n), K) for _ in range(log_n)]
        a = fill_uniform(rng, (size, rank + 1, n), res_k)
        outs = []
        for fl in (O.NTT120, O.FFT64):
            o = O.OracleModule(n, fl)
            keys = []
            for mt in mats:
                pm = o.vmp_pmat_alloc(mt.shape[0], rank, rank + 1, size + 1)
                o.vmp_prepare(pm, mt)
                keys.append(pm)
            r = a.copy()
            o.glwe_trace_assign(r, res_k, skip, keys, K, dsize)
            outs.append(r)
        assert np.array_equal(outs[0], outs[1]), (rank, dsize, res_k, skip)


def test_ggsw_expand_row_pins():
    """conversion/gglwe_to_ggsw.rs:116-268.  Zero mask: the gadget products vanish, so GLWE (row, col) is zero except for its column `col`,
    which carries the column-0 body (already normalised: unchanged) -- the '+ M[i] on the diagonal' of the reference's comment; and the
    cross-backend procedure on random inputs."""
    n, K, dnum, size = 64, 12, 2, 3
    rng = np.random.default_rng(4)
    for rank in (1, 2):
        mats = [fill_uniform(rng, (size, rank, size + 1, rank + 1, n), K) for _ in range(rank)]
        outs = []
        for fl in (O.NTT120, O.FFT64):
            o = O.OracleModule(n, fl)
            tsk = []
            for mt in mats:
                pm = o.vmp_pmat_alloc(size, rank, rank + 1, size + 1)
                o.vmp_prepare(pm, mt)
                tsk.append(pm)
            g = np.zeros((dnum, rank + 1, size, rank + 1, n), dtype=np.int64)
            body = fill_uniform(np.random.default_rng(5), (dnum, size, n), K)
            g[:, 0, :, 0] = body
            o.ggsw_expand_row(g, K, tsk, K)
            for col in range(1, rank + 1):
                for j in range(rank + 1):
                    assert np.array_equal(g[:, col, :, j], body if j == col else np.zeros_like(body)), (fl, col, j)
            r = fill_uniform(np.random.default_rng(6), (dnum, rank + 1, size, rank + 1, n), K)
            o.ggsw_expand_row(r, K, tsk, K)
            outs.append(r)
        assert np.array_equal(outs[0], outs[1]), rank
