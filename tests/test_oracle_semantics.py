"""The oracle's COMPOSITIONS pinned independently of oracle/core.c ("next round" item 3a / 3b), CPU only.

(a) Big-integer model: for a whole key-switch and a whole external product at n = 64 the torus value of every normalised output column
    equals -- up to the rounding of the final normalisation -- the exact rational  sum_rows a_row (*) key_row  (+ body) computed with
    Python integers from the reference's call sequence (tests/semantics.py), for dsize 1 and 2, equal and mixed base2k, truncating and
    extending result sizes.  A mis-restated row order, limb offset, size rule or body term shows up as an O(1) torus error.
(b) L4, noiseless keys (reference: poulpy-core/src/test_suite/keyswitch/glwe_ct.rs:132-156 decrypts and bounds the noise):
    with a noise-free key-switching key from s to s', key-switching a ciphertext of phase m under s gives a ciphertext whose phase under
    s' is m -- EXACTLY, up to the final rounding; likewise an external product by a noise-free GGSW(X^e) multiplies the phase by X^e."""
from fractions import Fraction

import numpy as np
import pytest

import semantics as S
from oracle import pyoracle as O
from util import fill_uniform, negacyclic_mul

N = 64


def _prep(o, mat):
    dnum, cols_in, size, cols_out, _ = mat.shape
    pm = o.vmp_pmat_alloc(dnum, cols_in, cols_out, size)
    o.vmp_prepare(pm, mat)
    return pm


def _to_key_base(a, a_k, key_k):
    """glwe_normalize of the input into the key's base2k when they differ (keyswitching/glwe.rs:92-100): leaf operation, pinned by the
    normalize property tests of test_oracle_kat.py."""
    if a_k == key_k:
        return a
    size = -(-a.shape[0] * a_k // key_k)
    out = np.zeros((size, a.shape[1], a.shape[2]), dtype=np.int64)
    for c in range(a.shape[1]):
        O.vec_znx_normalize(out, key_k, 0, c, a, a_k, c)
    return out


CASES = [
    # dsize, a_k, key_k, res_k, a_size, key_size, res_size
    (1, 12, 12, 12, 3, 4, 4),   # no truncation: exact
    (1, 12, 12, 12, 3, 4, 2),   # truncating normalisation
    (1, 12, 12, 12, 2, 3, 5),   # result longer than the key
    (2, 12, 12, 12, 4, 5, 5),   # two digit groups
    (2, 12, 12, 12, 5, 4, 3),   # a_size not a multiple of dsize, truncation
    (1, 10, 12, 12, 4, 4, 4),   # input converted into the key's base first
    (1, 12, 12, 9, 3, 4, 5),    # output in another base
    (2, 11, 13, 10, 4, 4, 4),   # everything mixed
]


@pytest.mark.parametrize("fl", [O.NTT120, O.FFT64])
@pytest.mark.parametrize("dsize,a_k,key_k,res_k,a_size,key_size,res_size", CASES)
def test_keyswitch_equals_bigint_model(fl, dsize, a_k, key_k, res_k, a_size, key_size, res_size):
    rng = np.random.default_rng(500 + dsize + a_k + res_size)
    o = O.OracleModule(N, fl)
    for rank_in, rank_out in ((1, 1), (2, 1), (1, 2)):
        a = fill_uniform(rng, (a_size, rank_in + 1, N), a_k)
        in_size = -(-a_size * a_k // key_k)
        dnum = -(-in_size // dsize)
        key = fill_uniform(rng, (dnum, rank_in, key_size, rank_out + 1, N), key_k)
        res = fill_uniform(rng, (res_size, rank_out + 1, N), res_k)  # garbage pre-fill
        o.glwe_keyswitch(res, res_k, a, a_k, _prep(o, key), key_k, dsize)
        ain = _to_key_base(a, a_k, key_k)
        want = S.keyswitch_torus(ain.tolist(), key.tolist(), key_k, dsize)
        exact = res_size * res_k >= key_size * key_k
        for c in range(rank_out + 1):
            S.assert_normalised_equals(res[:, c, :].tolist(), res_k, want[c], None if exact else res_size * res_k, ("ks", rank_in, rank_out, c),
                                       balanced=res_k == key_k)


@pytest.mark.parametrize("fl", [O.NTT120, O.FFT64])
@pytest.mark.parametrize("dsize,a_k,key_k,res_k,a_size,key_size,res_size", CASES)
def test_external_product_equals_bigint_model(fl, dsize, a_k, key_k, res_k, a_size, key_size, res_size):
    rng = np.random.default_rng(600 + dsize + a_k + res_size)
    o = O.OracleModule(N, fl)
    for rank in (1, 2):
        a = fill_uniform(rng, (a_size, rank + 1, N), a_k)
        in_size = -(-a_size * a_k // key_k)
        dnum = -(-in_size // dsize)
        ggsw = fill_uniform(rng, (dnum, rank + 1, key_size, rank + 1, N), key_k)
        res = fill_uniform(rng, (res_size, rank + 1, N), res_k)
        o.glwe_external_product(res, res_k, a, a_k, _prep(o, ggsw), key_k, dsize)
        ain = _to_key_base(a, a_k, key_k)
        want = S.external_product_torus(ain.tolist(), ggsw.tolist(), key_k, dsize)
        exact = res_size * res_k >= key_size * key_k
        for c in range(rank + 1):
            S.assert_normalised_equals(res[:, c, :].tolist(), res_k, want[c], None if exact else res_size * res_k, ("ep", rank, c),
                                       balanced=res_k == key_k)


# ---- L4: noiseless keys ---------------------------------------------------------------------------------------------------------------
def _ternary(rng, n):
    return [int(x) for x in rng.integers(-1, 2, size=n)]


def noiseless_ksk(rng, s_in, s_out, k, dnum, key_size, dsize=1):
    """GGLWE key-switching key from the secrets s_in to s_out with zero noise (poulpy-core/src/encryption/gglwe.rs: row d, input column
    ci encrypts s_in[ci] * 2^(-(d+1) * dsize * k) under s_out with a uniform mask) -> int64 (dnum, rank_in, key_size, rank_out + 1, n)."""
    n = len(s_out[0])
    key = np.zeros((dnum, len(s_in), key_size, len(s_out) + 1, n), dtype=np.int64)
    for d in range(dnum):
        limb = (d + 1) * dsize - 1  # the limb that carries weight 2^(-(d+1) dsize k)
        for ci, si in enumerate(s_in):
            msg = [[0] * n for _ in range(key_size)]
            if limb < key_size:
                msg[limb] = list(si)
            masks = [fill_uniform(rng, (key_size, n), k).tolist() for _ in s_out]
            row = S.noiseless_glwe_row(msg, masks, s_out, k, key_size)
            key[d, ci] = np.array(row, dtype=np.int64)
    return key


@pytest.mark.parametrize("fl", [O.NTT120, O.FFT64])
@pytest.mark.parametrize("dsize", [1, 2])
def test_noiseless_keyswitch_preserves_the_phase(fl, dsize):
    k, a_size = 12, 4
    key_size = a_size + dsize  # room for the lowest gadget digit
    rng = np.random.default_rng(700 + fl + dsize)
    o = O.OracleModule(N, fl)
    for rank_in, rank_out in ((1, 1), (2, 1), (1, 2)):
        s_in = [_ternary(rng, N) for _ in range(rank_in)]
        s_out = [_ternary(rng, N) for _ in range(rank_out)]
        dnum = -(-a_size // dsize)
        key = noiseless_ksk(rng, s_in, s_out, k, dnum, key_size, dsize)
        a = fill_uniform(rng, (a_size, rank_in + 1, N), k)  # any ciphertext: its phase under s_in is the "message"
        want = S.phase(a.tolist(), s_in, k)
        for res_size in (key_size, a_size):
            res = np.zeros((res_size, rank_out + 1, N), dtype=np.int64)
            o.glwe_keyswitch(res, k, a, k, _prep(o, key), k, dsize)
            got = S.phase(res.tolist(), s_out, k)
            # exact when nothing is truncated; otherwise every column is rounded at 2^-(res_size k), and the mask columns are multiplied by
            # a ternary secret of at most n non-zero coefficients
            tol = Fraction(0) if res_size == key_size else Fraction(1 + N * rank_out, 1 << (res_size * k))
            for i, (g, w) in enumerate(zip(got, want)):
                assert abs(S.centred_mod1(g - w)) <= tol, (rank_in, rank_out, res_size, i)


@pytest.mark.parametrize("fl", [O.NTT120, O.FFT64])
def test_noiseless_external_product_multiplies_the_phase(fl):
    """GGSW(m) with zero noise (poulpy-core/src/encryption/ggsw.rs:62-120 with glwe_encrypt_sk_internal, encryption/glwe.rs:426-510: row d,
    column 0 has phase m * 2^(-(d+1)k), column c > 0 has phase m * s_{c-1} * 2^(-(d+1)k)), so sum_c a_c (*) row_c has phase
    m * phase(a); m = X^e: the phase comes out rotated."""
    k, size, rank, e = 12, 4, 2, 5
    rng = np.random.default_rng(800 + fl)
    o = O.OracleModule(N, fl)
    s = [_ternary(rng, N) for _ in range(rank)]
    m = [0] * N
    m[e] = 1
    ggsw = np.zeros((size, rank + 1, size, rank + 1, N), dtype=np.int64)
    for d in range(size):
        for c in range(rank + 1):
            pt = m if c == 0 else negacyclic_mul(m, s[c - 1])  # phase(a) = a_0 + sum a_c s_c, so column c must carry m s_{c-1}
            msg = [[0] * N for _ in range(size)]
            msg[d] = list(pt)
            masks = [fill_uniform(rng, (size, N), k).tolist() for _ in s]
            ggsw[d, c] = np.array(S.noiseless_glwe_row(msg, masks, s, k, size), dtype=np.int64)
    a = fill_uniform(rng, (size, rank + 1, N), k)
    res = np.zeros((size, rank + 1, N), dtype=np.int64)
    o.glwe_external_product(res, k, a, k, _prep(o, ggsw), k, 1)
    ph = S.phase(a.tolist(), s, k)
    want = [(-ph[i - e + N] if i < e else ph[i - e]) for i in range(N)]  # X^e * phase, negacyclic
    got = S.phase(res.tolist(), s, k)
    for i, (g, w) in enumerate(zip(got, want)):
        assert S.centred_mod1(g - w) == 0, i
