"""Pins the CPU oracle against the reference's own known-answer tests.

The reference (Rust) cannot be built in this image, so these restate its unit tests:
  * poulpy-cpu-ref/src/reference/ntt120/primes.rs:80-90           Primes30 constants
  * poulpy-cpu-ref/src/reference/ntt120/ntt.rs:843-870           ntt_intt_identity
  * poulpy-cpu-ref/src/reference/ntt120/ntt.rs:877-911           ntt_convolution -> [3, 10, 8, 0, ...]
  * poulpy-cpu-ref/src/reference/ntt120/ntt.rs:175-345           reduction split h = 47 and level bit sizes
  * poulpy-cpu-ref/src/reference/ntt120/mat_vec.rs:184-212       bbc split h = 25
  * poulpy-cpu-ref/src/reference/vec_znx/normalize.rs:427-633    normalize torus-value property tests
  * poulpy-cpu-avx/src/fft64/reim/fft_avx2_fma.rs:241-353        FFT round trip / convolution bound
plus an independent big-integer negacyclic schoolbook check of the NTT120 pipeline.
"""
import ctypes as C
from fractions import Fraction

import numpy as np
import pytest

from oracle import pyoracle as O
from util import bitrev, fill_uniform, negacyclic_mul

L = O.lib()


def test_primes30_constants():
    assert O.Q == tuple((1 << 30) - c * (1 << 17) + 1 for c in (2, 17, 23, 42))
    total = 1
    for q in O.Q:
        total *= q
    for k, q in enumerate(O.Q):
        assert pow(O.OMEGA[k], 1 << 16, q) == q - 1          # primitive 2^17-th root
        assert (total // q) * O.CRT_CST[k] % q == 1          # CRT_CST = (Q/Q_k)^-1 mod Q_k
    libq = (C.c_uint32 * 4).in_dll(L, "ORC_Q")
    libo = (C.c_uint32 * 4).in_dll(L, "ORC_OMEGA")
    libc = (C.c_uint32 * 4).in_dll(L, "ORC_CRT_CST")
    assert tuple(libq) == O.Q and tuple(libo) == O.OMEGA and tuple(libc) == O.CRT_CST


def test_ntt_table_metadata():
    m = O.OracleModule(4096, O.NTT120)
    bs = (C.c_uint64 * 20)()
    rd = (C.c_int * 20)()
    nl = L.orc_ntt120_fwd_levels(m._h, bs, rd, C.c_size_t(20))
    got = [(int(bs[i]), int(rd[i])) for i in range(nl)]
    want = [(63, 0), (64, 0), (56, 1), (60, 0), (62, 0), (63, 0), (64, 0), (56, 1), (60, 0), (62, 0), (63, 0), (64, 0), (49, 1)]
    assert got == want
    assert L.orc_ntt120_reduc_h(m._h) == 47
    assert L.orc_ntt120_bbc_h(m._h) == 25


@pytest.mark.parametrize("log_n", range(1, 9))
def test_ntt_intt_identity(log_n):
    n = 1 << log_n
    m = O.OracleModule(n, O.NTT120)
    coeffs = np.array([(i * 7 + 3) % 201 - 100 for i in range(n)], dtype=np.int64)
    d = O.ntt120_b_from_znx64(coeffs)
    orig = d.copy()
    L.orc_ntt120_ntt(m._h, O._p(d))
    L.orc_ntt120_intt(m._h, O._p(d))
    for k, q in enumerate(O.Q):
        assert np.array_equal(orig[:, k] % np.uint64(q), d[:, k] % np.uint64(q))


def test_ntt_convolution_kat():
    n = 8
    m = O.OracleModule(n, O.NTT120)
    a = np.array([1, 2, 0, 0, 0, 0, 0, 0], dtype=np.int64)
    b = np.array([3, 4, 0, 0, 0, 0, 0, 0], dtype=np.int64)
    da, db = O.ntt120_b_from_znx64(a), O.ntt120_b_from_znx64(b)
    L.orc_ntt120_ntt(m._h, O._p(da))
    L.orc_ntt120_ntt(m._h, O._p(db))
    dc = np.zeros_like(da)
    for i in range(n):
        for k, q in enumerate(O.Q):
            dc[i, k] = (int(da[i, k]) % q) * (int(db[i, k]) % q) % q
    L.orc_ntt120_intt(m._h, O._p(dc))
    assert list(O.ntt120_b_to_znx128(dc)) == [3, 10, 8, 0, 0, 0, 0, 0]


@pytest.mark.parametrize("n", [2, 16, 128])
def test_ntt_frequency_order(n):
    """out[bitrev(j)] = sum_i a_i * w^(i(2j+1)) mod Q_k."""
    m = O.OracleModule(n, O.NTT120)
    rng = np.random.default_rng(n)
    x = fill_uniform(rng, n, 64)
    d = O.ntt120_b_from_znx64(x)
    L.orc_ntt120_ntt(m._h, O._p(d))
    bits = n.bit_length() - 1
    for k, q in enumerate(O.Q):
        w = pow(O.OMEGA[k], (1 << 16) // n, q)
        for j in range(n):
            v = sum(int(x[i]) * pow(w, i * (2 * j + 1), q) for i in range(n)) % q
            assert v == int(d[bitrev(j, bits), k]) % q


@pytest.mark.parametrize("n,base2k", [(16, 12), (64, 18), (64, 52)])
def test_ntt120_product_vs_bigint_schoolbook(n, base2k):
    """dft -> svp (pointwise) -> idft -> i128 equals the centred negacyclic product mod Q."""
    m = O.OracleModule(n, O.NTT120)
    rng = np.random.default_rng(7 * n + base2k)
    a = fill_uniform(rng, (1, 1, n), base2k)
    s = fill_uniform(rng, (1, n), base2k)
    a_dft = m.vec_znx_dft_alloc(1, 1)
    m.vec_znx_dft_apply(1, 0, a_dft, 0, a, 0)
    pp = m.svp_ppol_alloc(1)
    m.svp_prepare(pp, 0, s, 0)
    r_dft = m.vec_znx_dft_alloc(1, 1)
    m.svp_apply_dft_to_dft(r_dft, 0, pp, 0, a_dft, 0)
    big = m.vec_znx_big_alloc(1, 1)
    m.vec_znx_idft_apply(big, 0, r_dft, 0)
    got = list(O.i128_to_int(big[0, 0]))
    total = O.Q[0] * O.Q[1] * O.Q[2] * O.Q[3]
    want = []
    for v in negacyclic_mul(a[0, 0], s[0]):
        v %= total
        want.append(v - total if v >= (total + 1) // 2 else v)
    assert got == want
    # consume variant gives the same bytes
    big2 = m.vec_znx_idft_apply_consume(r_dft.copy())
    assert np.array_equal(big2, big)


def _torus(limbs, base2k):
    return sum(Fraction(int(v), 1 << ((j + 1) * base2k)) for j, v in enumerate(limbs))


def _reduce(x):
    r = x - (x.numerator // x.denominator)
    return r - 1 if r >= Fraction(1, 2) else r


def _torus_err(have, want):
    e = abs(_reduce(have) - _reduce(want))
    return min(e, 1 - e)


@pytest.mark.parametrize("base2k", [1, 2, 7, 12, 18, 31, 50, 51])
def test_normalize_inter_base2k_property(base2k):
    """reference/vec_znx/normalize.rs:538-633 with Fractions instead of dashu floats."""
    n, prec = 8, 128
    rng = np.random.default_rng(base2k)
    size = -(-prec // base2k)
    for offset in range(-prec, prec + 1, base2k + 1):
        want = fill_uniform(rng, (size, 1, n), 60)
        have = fill_uniform(rng, (size, 1, n), 60)
        O.vec_znx_normalize(have, base2k, offset, 0, want, base2k, 0)
        for i in range(n):
            w = _torus(want[:, 0, i], base2k) * (Fraction(2) ** offset)
            h = _torus(have[:, 0, i], base2k)
            assert _torus_err(h, w) <= Fraction(1, 1 << (size * base2k)), (base2k, offset, i)
        # i128 twin (NTT120 big) must agree bit for bit with the i64 path on sign-extended input
        big = O.int_to_i128(want.astype(object))
        have128 = fill_uniform(rng, (size, 1, n), 60)
        O.OracleModule(n, O.NTT120).vec_znx_big_normalize(have128, base2k, offset, 0, big, base2k, 0)
        assert np.array_equal(have128, have), (base2k, offset)


@pytest.mark.parametrize("in_base2k,out_base2k", [(1, 3), (5, 2), (12, 17), (17, 12), (18, 19), (51, 50), (50, 51), (7, 51), (51, 7), (52, 18)])
def test_normalize_cross_base2k_property(in_base2k, out_base2k):
    """reference/vec_znx/normalize.rs:427-536."""
    n, prec = 8, 128
    rng = np.random.default_rng(100 * in_base2k + out_base2k)
    k = in_base2k
    for offset in [-prec, -(prec - 1), -(prec - k), -(k + 1), k, -(k - 1), 0, k - 1, k, k + 1, prec - k, prec - 1, prec]:
        in_size = -(-prec // in_base2k)
        in_prec = in_size * in_base2k
        out_size = -(-in_prec // out_base2k)
        min_prec = min(in_prec, out_size * out_base2k)
        want = fill_uniform(rng, (in_size, 1, n), min(60, 63))
        have = fill_uniform(rng, (out_size, 1, n), 60)
        O.vec_znx_normalize(have, out_base2k, offset, 0, want, in_base2k, 0)
        for i in range(n):
            w = _torus(want[:, 0, i], in_base2k) * (Fraction(2) ** offset)
            h = _torus(have[:, 0, i], out_base2k)
            assert _torus_err(h, w) <= Fraction(2, 1 << min_prec), (offset, i)
        big = O.int_to_i128(want.astype(object))
        have128 = fill_uniform(rng, (out_size, 1, n), 60)
        O.OracleModule(n, O.NTT120).vec_znx_big_normalize(have128, out_base2k, offset, 0, big, in_base2k, 0)
        assert np.array_equal(have128, have), offset


@pytest.mark.parametrize("log_m", range(1, 14))
def test_fft_roundtrip_and_evaluation(log_m):
    """FFT KATs: ifft(fft(x))/m == x, and out[p] = a(zeta^(4*bitrev(p)+1)), zeta = exp(2*pi*i/(2n)).
    Tolerance precedent: poulpy-cpu-avx/src/fft64/reim/fft_avx2_fma.rs:318-353, 2^-(53-log_m-1) on (0,2] ramps."""
    m = 1 << log_m
    n = 2 * m
    md = O.OracleModule(n, O.FFT64)
    ramp = np.array([(i + 1) / m for i in range(n)], dtype=np.float64)
    d = ramp.copy()
    L.orc_fft64_fft(md._h, O._p(d))
    if log_m <= 8:
        z = ramp[:m] + 1j * ramp[m:]
        for p in range(m):
            e = 4 * bitrev(p, log_m) + 1
            root = np.exp(2j * np.pi * e / (2 * n))
            want = np.sum(z * root ** np.arange(m))
            assert abs(want - (d[p] + 1j * d[m + p])) <= 1e-9 * m
    L.orc_fft64_ifft(md._h, O._p(d))
    assert np.max(np.abs(d / m - ramp)) <= 2.0 ** -(53 - log_m - 1) * 4


def test_avx2_data_path_is_bit_identical_to_the_scalar_port():
    """The "cpu-avx-style" leaves (four primes per __m256i: poulpy-cpu-avx/src/ntt120/ntt.rs:81-110, mat_vec_avx.rs) compute the same lazy
    u64 in every lane as the scalar restatement of poulpy-cpu-ref: forward / inverse NTT values, vmp results and a whole key-switch and
    blind rotation agree bit for bit (not just modulo Q_k), for every n the reduction schedule changes at."""
    from util import fill_uniform
    rng = np.random.default_rng(77)
    try:
        for n in (8, 64, 1024, 4096, 16384):
            o = O.OracleModule(n, O.NTT120)
            a = fill_uniform(rng, (3, 2, n), 40)
            mat = fill_uniform(rng, (3, 1, 4, 2, n), 18)
            got = []
            for simd in (False, True):
                O.ntt120_set_simd(simd)
                d = o.vec_znx_dft_alloc(2, 3)
                for c in range(2):
                    o.vec_znx_dft_apply(1, 0, d, c, a, c)
                fwd = np.array(d, copy=True)
                pm = o.vmp_pmat_alloc(3, 1, 2, 4)
                o.vmp_prepare(pm, mat)
                a1 = np.ascontiguousarray(d[:, :1])
                r = o.vec_znx_dft_alloc(2, 4)
                o.vmp_apply_dft_to_dft(r, a1, pm, 0)
                prod = np.array(r, copy=True)
                big = o.vec_znx_idft_apply_consume(r)
                res = np.zeros((3, 2, n), dtype=np.int64)
                o.glwe_keyswitch(res, 18, fill_uniform(np.random.default_rng(5), (3, 2, n), 18), 18, pm, 18, 1)
                got.append((fwd, prod, np.array(big, copy=True), res))
            for x, y in zip(*got):
                assert np.array_equal(x, y), n
    finally:
        O.ntt120_set_simd(False)
