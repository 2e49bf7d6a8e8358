"""Pins the oracle's bivariate convolution (HalImpl::cnv_*) and GLWE tensoring by the reference's own test procedures, restated:
poulpy-hal/src/test_suite/convolution.rs:21-245 compares cnv_apply_dft / cnv_pairwise_apply_dft / cnv_by_const_apply (after idft and
normalize) with `bivariate_convolution_naive` (:247-296): limb (a_limb + b_limb + 1 - k) += naive negacyclic product, then
vec_znx_normalize_assign.  Both flavours must also agree with each other (poulpy-cpu-ref/src/tests.rs cross-backend procedure)."""
import numpy as np
import pytest

from oracle import pyoracle as O
from util import fill_uniform


def negacyclic(a, b):
    n = len(a)
    full = np.convolve(a.astype(object), b.astype(object))
    res = full[:n].copy()
    res[: n - 1] -= full[n:]
    return res


def naive_bivariate(base2k, k, res_size, a, b):
    """a, b: (size, n) int64 -> (res_size, n) int64 (test_suite/convolution.rs:247-296)."""
    n = a.shape[1]
    res = np.zeros((res_size, n), dtype=object)
    for ai in range(a.shape[0]):
        for bi in range(b.shape[0]):
            limb = ai + bi + 1
            if k <= 0:
                limb += -k
            elif limb >= k:
                limb -= k
            else:
                continue
            if limb < res_size:
                res[limb] += negacyclic(a[ai], b[bi])
    wrapped = ((res + (1 << 63)) % (1 << 64) - (1 << 63)).astype(np.int64)  # the reference accumulates in (wrapping) i64
    out = wrapped.reshape(res_size, 1, n).copy()
    O.vec_znx_normalize_assign(base2k, out, 0)
    return out[:, 0]


def _norm(m, res_size, base2k, big):
    out = np.zeros((res_size, 1, m.n), dtype=np.int64)
    m.vec_znx_big_normalize(out, base2k, 0, 0, big, base2k, 0)
    return out[:, 0]


@pytest.mark.parametrize("flavour", [O.NTT120, O.FFT64])
def test_convolution_matches_naive(flavour):
    """test_suite/convolution.rs:85-158 with n = 16, a_size = b_size = 5 (the reference uses 15), base2k = 12, 11-bit digits."""
    n, base2k, a_size, b_size = 16, 12, 5, 5
    res_size = a_size + b_size
    m = O.OracleModule(n, flavour)
    rng = np.random.default_rng(41)
    a, b = fill_uniform(rng, (a_size, 2, n), 11), fill_uniform(rng, (b_size, 2, n), 11)
    ap, bp = m.vec_znx_dft_alloc(2, a_size), m.vec_znx_dft_alloc(2, b_size)
    m.cnv_prepare(ap, a)
    m.cnv_prepare(bp, b)
    for a_col in range(2):
        for b_col in range(2):
            for off in range(res_size):
                rd = m.vec_znx_dft_alloc(1, res_size)
                m.cnv_apply_dft(off, rd, 0, ap, a_col, bp, b_col)
                got = _norm(m, res_size, base2k, m.vec_znx_idft_apply_consume(rd))
                want = naive_bivariate(base2k, off + 1, res_size, a[:, a_col], b[:, b_col])
                assert np.array_equal(got, want), (a_col, b_col, off)


@pytest.mark.parametrize("flavour", [O.NTT120, O.FFT64])
def test_convolution_pairwise_and_by_const(flavour):
    """test_suite/convolution.rs:160-245 (pairwise = convolution of the column sums) and :21-83 (by_const)."""
    n, base2k, a_size, b_size = 16, 12, 4, 3
    res_size = a_size + b_size
    m = O.OracleModule(n, flavour)
    rng = np.random.default_rng(43)
    a, b = fill_uniform(rng, (a_size, 3, n), 10), fill_uniform(rng, (b_size, 3, n), 10)
    ap, bp = m.vec_znx_dft_alloc(3, a_size), m.vec_znx_dft_alloc(3, b_size)
    m.cnv_prepare(ap, a)
    m.cnv_prepare(bp, b)
    for (i, j) in ((0, 1), (1, 2), (0, 2), (1, 1)):
        for off in (0, 1, 3, res_size - 1):
            rd = m.vec_znx_dft_alloc(1, res_size)
            m.cnv_pairwise_apply_dft(off, rd, 0, ap, bp, i, j)
            got = _norm(m, res_size, base2k, m.vec_znx_idft_apply_consume(rd))
            sa = a[:, i] + a[:, j] if i != j else a[:, i]
            sb = b[:, i] + b[:, j] if i != j else b[:, i]
            assert np.array_equal(got, naive_bivariate(base2k, off + 1, res_size, sa, sb)), (i, j, off)
    bc = fill_uniform(rng, (b_size,), 10)
    bvec = np.zeros((b_size, n), dtype=np.int64)
    bvec[:, 0] = bc
    for off in range(res_size):
        big = m.vec_znx_big_alloc(1, res_size)
        m.cnv_by_const_apply(off, big, 0, a, 1, bc)
        got = _norm(m, res_size, base2k, big)
        assert np.array_equal(got, naive_bivariate(base2k, off + 1, res_size, a[:, 1], bvec)), off


def test_prepare_mask_and_short_results():
    """The last active limb is ANDed with the mask (convolution.rs:86-93); res shorter than a.size + b.size - 1 truncates."""
    n, base2k = 16, 12
    rng = np.random.default_rng(47)
    a, b = fill_uniform(rng, (3, 1, n), 11), fill_uniform(rng, (2, 1, n), 11)
    mask = -1 << 5
    am = a.copy()
    am[2] &= mask
    for flavour in (O.NTT120, O.FFT64):
        m = O.OracleModule(n, flavour)
        ap, bp = m.vec_znx_dft_alloc(1, 3), m.vec_znx_dft_alloc(1, 2)
        m.cnv_prepare(ap, a, mask)
        m.cnv_prepare(bp, b)
        rd = m.vec_znx_dft_alloc(1, 3)
        m.cnv_apply_dft(1, rd, 0, ap, 0, bp, 0)
        got = _norm(m, 3, base2k, m.vec_znx_idft_apply_consume(rd))
        # exact: un-normalised limbs k = 1, 2, 3 of the full product, normalised over 3 limbs
        full = np.zeros((3, n), dtype=object)
        for k in range(3):
            for j in range(2):
                if 0 <= k + 1 - j < 3:
                    full[k] += negacyclic(am[k + 1 - j, 0], b[j, 0])
        want = full.astype(np.int64).reshape(3, 1, n).copy()
        O.vec_znx_normalize_assign(base2k, want, 0)
        assert np.array_equal(got, want[:, 0]), flavour


@pytest.mark.parametrize("rank", [1, 2])
@pytest.mark.parametrize("cnv_offset", [5, 12, 20, 30])
def test_tensor_apply_cross_flavour_and_definition(rank, cnv_offset):
    """glwe_tensor_apply (operations/glwe.rs:699-818): the NTT120 and FFT64 oracles agree, and every tensor column equals its
    definition assembled from exact integer convolutions: c(i,i) = N(a_i x b_i), c(i,j) = N((a_i+a_j) x (b_i+b_j)) - c(i,i) - c(j,j),
    with N = big_normalize(res_base2k, cnv_offset_lo) of the limbs [cnv_offset_hi, ...) of the bivariate product."""
    n, ab, res_k, size = 16, 12, 12, 3
    cols = rank + 1
    rng = np.random.default_rng(53 + rank + cnv_offset)
    a, b = fill_uniform(rng, (size, cols, n), ab), fill_uniform(rng, (size, cols, n), ab)
    outs = []
    for flavour in (O.NTT120, O.FFT64):
        m = O.OracleModule(n, flavour)
        res = fill_uniform(rng, (size, cols * (cols + 1) // 2, n), ab)  # garbage: every column is overwritten
        m.glwe_tensor_apply(cnv_offset, res, res_k, a, size * ab, b, size * ab, ab)
        outs.append(res)
    assert np.array_equal(outs[0], outs[1])
    m = O.OracleModule(n, O.NTT120)
    if cnv_offset < ab:
        hi, lo = 0, -(ab - cnv_offset % ab)
    else:
        hi, lo = cnv_offset // ab - 1, cnv_offset % ab
    ob = lo % ab
    dft_size = min(2 * size - hi, -(-(size * res_k + ob) // ab))

    def N(x, y):  # exact limbs k_abs = k + hi of the bivariate product as an i128 big, then the reference's normalize
        big = np.zeros((dft_size, 1, n, 2), dtype=np.uint64)
        for k in range(dft_size):
            acc = np.zeros(n, dtype=object)
            for j in range(size):
                if 0 <= k + hi - j < size:
                    acc += negacyclic(x[k + hi - j], y[j])
            for i in range(n):
                v = int(acc[i]) % (1 << 128)
                big[k, 0, i, 0], big[k, 0, i, 1] = v & ((1 << 64) - 1), v >> 64
        out = np.zeros((size, 1, n), dtype=np.int64)
        m.vec_znx_big_normalize(out, res_k, lo, 0, big, ab, 0)
        return out[:, 0]

    diag = [N(a[:, i], b[:, i]) for i in range(cols)]
    for i in range(cols):
        ci = i * cols - i * (i + 1) // 2
        assert np.array_equal(outs[0][:, ci + i], diag[i]), ("diag", i)
        for j in range(i + 1, cols):
            want = N(a[:, i] + a[:, j], b[:, i] + b[:, j]) - diag[i] - diag[j]
            assert np.array_equal(outs[0][:, ci + j], want), ("pair", i, j)


@pytest.mark.parametrize("flavour", [O.NTT120, O.FFT64])
def test_tensor_relinearize_with_trivial_keys(flavour):
    """glwe_tensor_relinearize (operations/glwe.rs:545-610) with noiseless gadget keys: a key whose row (limb j, pair p) holds the
    constant 1 at limb j of output column c sends the quadratic column cols + p onto column c, so
    res[c] = normalize(a[c] + sum_{p -> c} a[cols + p]); a zero key leaves normalize(a[:cols])."""
    n, k, size, rank = 16, 12, 3, 2
    cols, pairs = rank + 1, rank * (rank + 1) // 2
    rng = np.random.default_rng(61 + flavour)
    m = O.OracleModule(n, flavour)
    a = fill_uniform(rng, (size, cols + pairs, n), k - 2)
    target = [0, 2, 1]  # pair p -> output column
    mat = np.zeros((size, pairs, size, cols, n), dtype=np.int64)
    for j in range(size):
        for p in range(pairs):
            mat[j, p, j, target[p], 0] = 1
    tsk = m.vmp_pmat_alloc(size, pairs, cols, size)
    m.vmp_prepare(tsk, mat)
    res = fill_uniform(rng, (size, cols, n), k)
    m.glwe_tensor_relinearize(res, k, a, k, tsk, k)
    want = a[:, :cols].copy()
    for p in range(pairs):
        want[:, target[p]] += a[:, cols + p]
    for c in range(cols):
        O.vec_znx_normalize_assign(k, want, c)
    assert np.array_equal(res, want)
