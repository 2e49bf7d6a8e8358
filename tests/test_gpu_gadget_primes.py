"""The NTT120 single-kernel gadget product on THREE primes (ntt120_gadget.cu, NP = 3): for a pinned key whose bound the host knows, the
integers of the collapsed-key product stay below Q[0] Q[1] Q[2] / 2, three residues reconstruct them exactly and the results are the
reference's bit for bit (poulpy-core/src/keyswitching/glwe.rs:207-239, external_product/glwe.rs:197-271, automorphism/glwe_ct.rs:51-275).
Every case is compared with the oracle AND with the four-prime launch of the same call; `OPT_LAST_GADGET_PRIMES` says which launch ran."""
import numpy as np
import pytest

import poulpy_b200 as pb
from poulpy_b200 import hal as H
from oracle import pyoracle as O
from util import fill_uniform

pytestmark = pytest.mark.gpu


def _pinned_key(g, o, mat):
    dnum, cols_in, size, cols_out, _ = mat.shape
    pg, po = g.vmp_pmat_alloc(dnum, cols_in, cols_out, size), o.vmp_pmat_alloc(dnum, cols_in, cols_out, size)
    g.vmp_prepare(pg, g.mat_znx_from_numpy(mat))
    o.vmp_prepare(po, mat)
    g.gadget_key_pin(pg)
    return pg, po


def _keyswitch_both(g, pg, a, res_shape, k, prefill):
    """-> (result of the default launch, primes it used, result of the forced four-prime launch)"""
    out = []
    for force in (0, 4):
        g.set_option(H.OPT_GADGET_PRIMES, force)
        res = g.vec_znx_from_numpy(prefill)
        g.glwe_keyswitch(res, k, g.vec_znx_from_numpy(a), k, pg, k)
        g.sync()
        out.append((g.vec_znx_to_numpy(res).reshape(res_shape), g.get_option(H.OPT_LAST_GADGET_PRIMES)))
    g.set_option(H.OPT_GADGET_PRIMES, 0)
    assert out[1][1] == 4
    return out[0][0], out[0][1], out[1][0]


# (n, rank_in, rank_out, a_size, key_size, res_size, base2k, primes expected)
SHAPES = [
    (4096, 1, 1, 3, 3, 3, 18, 3),   # the headline key-switch
    (4096, 1, 1, 3, 3, 2, 18, 3),   # fewer output limbs than key limbs
    (4096, 1, 1, 3, 3, 5, 18, 3),   # more (zero-filled)
    (2048, 2, 1, 3, 3, 3, 18, 3),   # R = 6
    (2048, 1, 2, 2, 3, 3, 17, 3),   # three output columns
    (1024, 1, 3, 2, 3, 3, 16, 3),   # four output columns
    (1024, 3, 1, 2, 2, 2, 18, 3),
    (1024, 1, 1, 4, 5, 5, 12, 3),   # S K = 60: two words of digits
    (1024, 1, 1, 1, 6, 6, 11, 3),   # S K = 66: three words of digits on three primes
    (1024, 1, 1, 1, 2, 2, 30, 4),   # wide digits: the key alone uses up the three-prime range
    (1024, 1, 1, 1, 1, 2, 32, 3),   # one 32-bit digit: the 64-bit digit path on three primes
    (4096, 1, 1, 3, 4, 4, 18, 4),   # S K = 72: beyond the three-prime bound, four primes as before
    (2048, 1, 1, 2, 3, 3, 25, 4),
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "n%d_r%d%d_a%d_s%d_o%d_k%d_p%d" % s)
def test_keyswitch_three_primes(shape):
    n, rank_in, rank_out, a_size, key_size, res_size, k, primes = shape
    g, o = pb.Module(n, pb.NTT120), O.OracleModule(n, pb.NTT120)
    rng = np.random.default_rng(9000 + n + 10 * k + key_size)
    batch = 160 if n == 4096 else 37   # 160 > 148 resident clusters: the persistent loop runs twice for some clusters
    mat = fill_uniform(rng, (a_size, rank_in, key_size, rank_out + 1, n), k)
    pg, po = _pinned_key(g, o, mat)
    a = fill_uniform(rng, (batch, a_size, rank_in + 1, n), k)
    a[3] = fill_uniform(rng, a[3].shape, 61)   # beyond any bound: flagged for the per-limb route
    a[7, 0, 1, 5] = 1 << (k + 3)               # slightly beyond normalised digits
    a[9, :, 1:, :] = -(1 << (k - 1))           # every mask digit at the extreme of the balanced range
    a[11] = 0
    prefill = fill_uniform(rng, (batch, res_size, rank_out + 1, n), k)
    want = prefill.copy()
    o.glwe_keyswitch_batch(want, k, a, k, po, k)
    got, used, got4 = _keyswitch_both(g, pg, a, want.shape, k, prefill)
    assert used == primes, (used, primes)
    bad = [b for b in range(batch) if not np.array_equal(got[b], want[b])]
    assert not bad, (bad[:10], len(bad))
    assert np.array_equal(got4, want)


@pytest.mark.parametrize("n", [1024, 4096])
def test_three_primes_worst_case_magnitudes(n):
    """Key digits and input digits all at -2^(K-1) (and the sign patterns that make every product add up): the largest integers the
    bound admits, |v| ~ R n 2^(2K-2) 2^((S-1)K) -- still reconstructed exactly from three residues."""
    k, a_size, key_size = 18, 3, 3
    g, o = pb.Module(n, pb.NTT120), O.OracleModule(n, pb.NTT120)
    ext = -(1 << (k - 1))
    mat = np.full((a_size, 1, key_size, 2, n), ext, dtype=np.int64)
    mat[:, :, :, 1, 1:] = -ext - 1              # column 1: x^0 negative, the rest positive: the negacyclic wrap aligns the signs
    pg, po = _pinned_key(g, o, mat)
    a = np.full((4, a_size, 2, n), ext, dtype=np.int64)
    a[1, :, 1, 1::2] = -ext - 1
    a[2, :, 1, :] = np.where(np.arange(n) < n // 2, ext, -ext - 1)
    a[3, :, 1, :] = np.random.default_rng(5).choice([ext, -ext - 1], size=(a_size, n))
    prefill = np.zeros((4, key_size, 2, n), dtype=np.int64)
    want = prefill.copy()
    o.glwe_keyswitch_batch(want, k, a, k, po, k)
    got, used, got4 = _keyswitch_both(g, pg, a, want.shape, k, prefill)
    assert used == 3
    assert np.array_equal(got, want) and np.array_equal(got4, want)


def test_unpinned_key_stays_on_four_primes():
    """Without a pinned key the host does not know the key's bound before the launch: four primes, as in round 1."""
    n, k = 2048, 18
    g, o = pb.Module(n, pb.NTT120), O.OracleModule(n, pb.NTT120)
    rng = np.random.default_rng(77)
    mat = fill_uniform(rng, (3, 1, 3, 2, n), k)
    pg, po = g.vmp_pmat_alloc(3, 1, 2, 3), o.vmp_pmat_alloc(3, 1, 2, 3)
    g.vmp_prepare(pg, g.mat_znx_from_numpy(mat))
    o.vmp_prepare(po, mat)
    a = fill_uniform(rng, (6, 3, 2, n), k)
    want = np.zeros((6, 3, 2, n), dtype=np.int64)
    o.glwe_keyswitch_batch(want, k, a, k, po, k)
    res = g.vec_znx_alloc(2, 3, 6)
    g.glwe_keyswitch(res, k, g.vec_znx_from_numpy(a), k, pg, k)
    g.sync()
    assert g.get_option(H.OPT_LAST_GADGET_PRIMES) == 4
    assert np.array_equal(g.vec_znx_to_numpy(res), want)
    g.gadget_key_pin(pg)
    g.glwe_keyswitch(res, k, g.vec_znx_from_numpy(a), k, pg, k)
    g.sync()
    assert g.get_option(H.OPT_LAST_GADGET_PRIMES) == 3
    assert np.array_equal(g.vec_znx_to_numpy(res), want)
    # a key with wider coefficients in the same pinned buffer: the cached bound is dropped with the key, four primes again
    mat2 = fill_uniform(rng, (3, 1, 3, 2, n), 40)
    g.vmp_prepare(pg, g.mat_znx_from_numpy(mat2))
    o.vmp_prepare(po, mat2)
    want = np.zeros((6, 3, 2, n), dtype=np.int64)
    o.glwe_keyswitch_batch(want, k, a, k, po, k)
    g.glwe_keyswitch(res, k, g.vec_znx_from_numpy(a), k, pg, k)
    g.sync()
    assert g.get_option(H.OPT_LAST_GADGET_PRIMES) == 4
    assert np.array_equal(g.vec_znx_to_numpy(res), want)


@pytest.mark.parametrize("rank", [1, 2])
def test_external_product_three_primes(rank):
    n, k, batch = 2048, 18, 21
    g, o = pb.Module(n, pb.NTT120), O.OracleModule(n, pb.NTT120)
    rng = np.random.default_rng(9100 + rank)
    a_size = 3 if rank == 1 else 2
    mat = fill_uniform(rng, (a_size, rank + 1, 3, rank + 1, n), k)
    pg, po = _pinned_key(g, o, mat)
    a = fill_uniform(rng, (batch, a_size, rank + 1, n), k)
    a[2] = fill_uniform(rng, a[2].shape, 50)
    want = np.zeros((batch, 3, rank + 1, n), dtype=np.int64)
    o.glwe_external_product_batch(want, k, a, k, po, k)
    for force, primes in ((0, 3), (4, 4)):
        g.set_option(H.OPT_GADGET_PRIMES, force)
        res = g.vec_znx_alloc(rank + 1, 3, batch)
        g.glwe_external_product(res, k, g.vec_znx_from_numpy(a), k, pg, k)
        g.sync()
        assert g.get_option(H.OPT_LAST_GADGET_PRIMES) == primes
        assert np.array_equal(g.vec_znx_to_numpy(res), want), force


@pytest.mark.parametrize("n", [1024, 4096])
def test_automorphism_family_three_primes(n):
    """The automorphism epilogues of the gadget kernel (glwe_automorphism, _add / _sub / _sub_negate, _add_assign) on three primes."""
    k, batch = 18, 6
    g, o = pb.Module(n, pb.NTT120), O.OracleModule(n, pb.NTT120)
    rng = np.random.default_rng(9200 + n)
    mat = fill_uniform(rng, (3, 1, 3, 2, n), k)
    pg, po = _pinned_key(g, o, mat)
    a = fill_uniform(rng, (batch, 3, 2, n), k)
    for p in (5, -1, 2 * n - 3):
        want = np.zeros((batch, 3, 2, n), dtype=np.int64)
        res = g.vec_znx_alloc(2, 3, batch)
        g.glwe_automorphism(res, k, g.vec_znx_from_numpy(a), k, pg, k, p)
        g.sync()
        assert g.get_option(H.OPT_LAST_GADGET_PRIMES) == 3
        for b in range(batch):
            o.glwe_automorphism(want[b], k, a[b], k, po, k, p)
        assert np.array_equal(g.vec_znx_to_numpy(res), want), p
        for op in (0, 1, 2):
            want = np.zeros((batch, 3, 2, n), dtype=np.int64)
            res = g.vec_znx_alloc(2, 3, batch)
            g.glwe_automorphism_op(op, res, k, g.vec_znx_from_numpy(a), pg, k, p)
            g.sync()
            for b in range(batch):
                o.glwe_automorphism_op(op, want[b], k, a[b], po, k, p)
            assert np.array_equal(g.vec_znx_to_numpy(res), want), (p, op)
        acc = a.copy()
        acc_g = g.vec_znx_from_numpy(acc)
        g.glwe_automorphism_add_assign(acc_g, k, pg, k, p)
        g.sync()
        for b in range(batch):
            o.glwe_automorphism_add_assign(acc[b], k, po, k, p)
        assert np.array_equal(g.vec_znx_to_numpy(acc_g), acc), p


def test_pinned_random_shapes():
    """The randomised sweep of test_gadget_kernel_random_shapes with every key pinned: both prime counts occur; all bit for bit."""
    n = 1024
    g, o = pb.Module(n, pb.NTT120), O.OracleModule(n, pb.NTT120)
    rng = np.random.default_rng(4343)
    seen = set()
    for trial in range(36):
        k = int(rng.choice([4, 8, 11, 13, 16, 18, 25, 31]))
        rank_in, rank_out = int(rng.integers(1, 4)), int(rng.integers(1, 4))
        a_size, key_size, res_size = int(rng.integers(1, 5)), int(rng.integers(1, 6)), int(rng.integers(1, 7))
        batch = int(rng.integers(1, 10))
        ext = trial % 3 == 2
        cols_in = rank_in + 1 if ext else rank_in
        cols_out = rank_in + 1 if ext else rank_out + 1
        mat = fill_uniform(rng, (a_size, cols_in, key_size, cols_out, n), k)
        pg, po = _pinned_key(g, o, mat)
        a = fill_uniform(rng, (batch, a_size, rank_in + 1, n), k)
        want = fill_uniform(rng, (batch, res_size, cols_out, n), 8)
        res_g = g.vec_znx_from_numpy(want)
        before = g.get_option(H.OPT_LAST_GADGET_PRIMES)
        g.set_option(H.OPT_LAST_GADGET_PRIMES, 0)
        if ext:
            g.glwe_external_product(res_g, k, g.vec_znx_from_numpy(a), k, pg, k)
            o.glwe_external_product_batch(want, k, a, k, po, k)
        else:
            g.glwe_keyswitch(res_g, k, g.vec_znx_from_numpy(a), k, pg, k)
            o.glwe_keyswitch_batch(want, k, a, k, po, k)
        g.sync()
        seen.add(g.get_option(H.OPT_LAST_GADGET_PRIMES))
        got = g.vec_znx_to_numpy(res_g).reshape(want.shape)
        assert np.array_equal(got, want), (trial, ext, k, rank_in, rank_out, a_size, key_size, res_size, batch, before)
        g.gadget_key_unpin(pg)
    assert {3, 4} <= seen, seen
