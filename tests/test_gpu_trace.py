"""GPU parity tests of the trace and its pieces: vec_znx_rsh_assign, vec_znx_big_automorphism(_assign),
glwe_automorphism_add_assign (poulpy-core/src/automorphism/glwe_ct.rs:142-183) and glwe_trace_assign (poulpy-core/src/glwe_trace.rs:
129-175), both flavours, equal and mixed base2k, dsize 1 and 2 -- normalised outputs bit for bit against the oracle."""
import numpy as np
import pytest

import poulpy_b200 as pb
from oracle import pyoracle as O
from util import fill_uniform

pytestmark = pytest.mark.gpu
FLAVOURS = [pb.NTT120, pb.FFT64]


def _key(g, o, rng, dnum, cols_in, cols_out, size, k):
    mat = fill_uniform(rng, (dnum, cols_in, size, cols_out, g.n), k)
    pg, po = g.vmp_pmat_alloc(dnum, cols_in, cols_out, size), o.vmp_pmat_alloc(dnum, cols_in, cols_out, size)
    g.vmp_prepare(pg, g.mat_znx_from_numpy(mat))
    o.vmp_prepare(po, mat)
    return pg, po


def test_vec_znx_rsh_assign():
    n = 64
    g = pb.Module(n, pb.FFT64)
    rng = np.random.default_rng(7)
    for K in (4, 12, 18, 52):
        for size in (1, 2, 3, 5):
            for k in (0, 1, 3, K - 1, K, K + 1, 2 * K, 2 * K + 3):
                if -(-k // K) > size:
                    continue
                a = fill_uniform(rng, (3, size, 2, n), min(K + 3, 62))  # batch of 3, unnormalised digits
                want = a.copy()
                for b in range(3):
                    O.vec_znx_rsh_assign(K, k, want[b], 1)
                v = g.vec_znx_from_numpy(a)
                g.vec_znx_rsh_assign(K, k, v, 1)
                g.sync()
                assert np.array_equal(g.vec_znx_to_numpy(v), want), (K, size, k)
    with pytest.raises(pb.PoulpyError):
        g.vec_znx_rsh_assign(4, 9, g.vec_znx_alloc(1, 2), 0)  # three limb steps on a two-limb vector


@pytest.mark.parametrize("fl", FLAVOURS)
def test_vec_znx_big_automorphism(fl):
    n = 128
    g = pb.Module(n, fl)
    rng = np.random.default_rng(8)
    # a big built from a small vector: big limbs = sign-extended i64
    a = fill_uniform(rng, (3, 2, n), 50)
    av = g.vec_znx_from_numpy(a)
    big = g.vec_znx_big_alloc(2, 3)
    for c in range(2):
        g.vec_znx_big_from_small(big, c, av, c)
    out = g.vec_znx_big_alloc(2, 4)
    for p in (-1, 5, 25, 2 * n - 1, 3):
        g.vec_znx_big_automorphism(p, out, 1, big, 0)
        res = g.vec_znx_alloc(2, 4)
        g.vec_znx_big_normalize(res, 60, 0, 1, out, 60, 1)  # digits < 2^59: normalisation at base 2^60 returns the values
        want = np.zeros((4, n), dtype=np.int64)
        for j in range(3):
            O.lib().orc_znx_automorphism(O.C.c_int64(p), O._p(want[j]), O._p(np.ascontiguousarray(a[j, 0])), O._sz(n))
        assert np.array_equal(g.vec_znx_to_numpy(res)[:, 1], want), p
        g.vec_znx_big_automorphism_assign(p, big, 1)
        g.vec_znx_big_normalize(res, 60, 0, 0, big, 60, 1)
        for j in range(3):
            O.lib().orc_znx_automorphism(O.C.c_int64(p), O._p(want[j]), O._p(np.ascontiguousarray(a[j, 1])), O._sz(n))
        assert np.array_equal(g.vec_znx_to_numpy(res)[:3, 0], want[:3]), ("assign", p)
        a[:, 1] = want[:3]


@pytest.mark.parametrize("fl", FLAVOURS)
@pytest.mark.parametrize("dsize", [1, 2])
def test_glwe_automorphism_add_assign(fl, dsize):
    n, batch = 256, 3
    g, o = pb.Module(n, fl), O.OracleModule(n, fl)
    rng = np.random.default_rng(300 + dsize + fl)
    b = 12 if fl == pb.FFT64 else 40
    for rank in (1, 2):
        for (res_k, key_k) in ((b, b), (b - 2, b)):
            size, key_size = 4, 5
            dnum = -(-size // dsize)
            pg, po = _key(g, o, rng, dnum, rank, rank + 1, key_size, key_k)
            want = fill_uniform(rng, (batch, size, rank + 1, n), res_k)
            res_g = g.vec_znx_from_numpy(want)
            for p in (-1, 5):
                g.glwe_automorphism_add_assign(res_g, res_k, pg, key_k, p, dsize)
                g.sync()
                for bi in range(batch):
                    o.glwe_automorphism_add_assign(want[bi], res_k, po, key_k, p, dsize)
                assert np.array_equal(g.vec_znx_to_numpy(res_g), want), (rank, res_k, key_k, p)


@pytest.mark.parametrize("fl", FLAVOURS)
def test_glwe_trace_assign(fl):
    n, batch, log_n = 64, 2, 6
    g, o = pb.Module(n, fl), O.OracleModule(n, fl)
    rng = np.random.default_rng(400 + fl)
    b = 12 if fl == pb.FFT64 else 30
    assert [g.trace_galois_element(i) for i in range(log_n)] == [O.trace_galois_element(i, n) for i in range(log_n)]
    assert g.trace_galois_element(0) == -1 and g.trace_galois_element(1) == 5 and g.trace_galois_element(3) == pow(5, 4, 2 * n)
    for rank, dsize in ((1, 1), (2, 1), (1, 2)):
        size, key_size = 4, 5
        dnum = -(-size // dsize)
        keys = [_key(g, o, rng, dnum, rank, rank + 1, key_size, b) for _ in range(log_n)]
        for res_k, skip in ((b, 0), (b, 3), (b - 1, 2), (b, log_n)):
            want = fill_uniform(rng, (batch, size, rank + 1, n), res_k)
            res_g = g.vec_znx_from_numpy(want)
            g.glwe_trace_assign(res_g, res_k, skip, [k[0] for k in keys], b, dsize)
            g.sync()
            for bi in range(batch):
                o.glwe_trace_assign(want[bi], res_k, skip, [k[1] for k in keys], b, dsize)
            assert np.array_equal(g.vec_znx_to_numpy(res_g), want), (rank, dsize, res_k, skip)


@pytest.mark.parametrize("fl", FLAVOURS)
@pytest.mark.parametrize("dsize", [1, 2])
def test_ggsw_expand_row(fl, dsize):
    """ggsw_expand_row (conversion/gglwe_to_ggsw.rs:116-268): columns 1..rank of a batch of GGSWs from their column-0 GLWEs and the tensor
    keys, equal and mixed base2k, ranks 1..2 -- every GLWE of the result bit for bit against the oracle, column 0 untouched."""
    n, batch = 256, 3
    g, o = pb.Module(n, fl), O.OracleModule(n, fl)
    rng = np.random.default_rng(500 + dsize + fl)
    b = 12 if fl == pb.FFT64 else 30
    for rank in (1, 2):
        for (res_k, tsk_k) in ((b, b), (b - 2, b)):
            dnum, size, tsk_size = 3, 3, 4
            conv = -(-size * res_k // tsk_k)
            tsk = [_key(g, o, rng, -(-conv // dsize), rank, rank + 1, tsk_size, tsk_k) for _ in range(rank)]
            want = fill_uniform(rng, (batch, dnum, rank + 1, size, rank + 1, n), res_k)  # garbage in columns >= 1 on both sides
            buf = pb.DevBuf(want.nbytes)
            buf.upload(want)
            g.ggsw_expand_row(buf, batch, dnum, rank, size, res_k, [t[0] for t in tsk], tsk_k, dsize)
            g.sync()
            col0 = want[:, :, 0].copy()
            for bi in range(batch):
                o.ggsw_expand_row(want[bi], res_k, [t[1] for t in tsk], tsk_k, dsize)
            got = buf.download(np.int64, want.shape)
            assert np.array_equal(got[:, :, 0], col0)
            assert np.array_equal(got, want), (rank, res_k, tsk_k)


@pytest.mark.parametrize("fl", FLAVOURS)
def test_glwe_automorphism_op_family(fl):
    """glwe_automorphism_add / _sub / _sub_negate (automorphism/glwe_ct.rs:95-275), out of place and in place, equal and mixed base2k."""
    n, batch = 256, 3
    g, o = pb.Module(n, fl), O.OracleModule(n, fl)
    rng = np.random.default_rng(600 + fl)
    b = 12 if fl == pb.FFT64 else 40
    for rank in (1, 2):
        for (res_k, key_k) in ((b, b), (b - 2, b)):
            size, key_size = 3, 4
            pg, po = _key(g, o, rng, size, rank, rank + 1, key_size, key_k)
            for op in (0, 1, 2):
                a = fill_uniform(rng, (batch, size, rank + 1, n), res_k)
                want = fill_uniform(rng, (batch, size, rank + 1, n), res_k)
                res_g, a_g = g.vec_znx_from_numpy(want), g.vec_znx_from_numpy(a)
                g.glwe_automorphism_op(op, res_g, res_k, a_g, pg, key_k, 5)
                g.sync()
                for bi in range(batch):
                    o.glwe_automorphism_op(op, want[bi], res_k, a[bi], po, key_k, 5)
                assert np.array_equal(g.vec_znx_to_numpy(res_g), want), (rank, res_k, key_k, op)
                # in place
                g.glwe_automorphism_op(op, a_g, res_k, a_g, pg, key_k, -1)
                g.sync()
                for bi in range(batch):
                    o.glwe_automorphism_op(op, a[bi], res_k, a[bi], po, key_k, -1)
                assert np.array_equal(g.vec_znx_to_numpy(a_g), a), ("assign", rank, res_k, key_k, op)


@pytest.mark.parametrize("fl", FLAVOURS)
@pytest.mark.parametrize("n", [1024, 4096])
def test_automorphism_single_kernel_route(n, fl):
    """log_n in 10..12 (FFT64: 9..12), one base2k: key-switch + X -> X^p + (add | sub | sub_negate) of the input run as ONE launch of the gadget
    kernel with the automorphism epilogue (ntt120_gadget.cu); the launch count proves the route, the oracle the bits.  Out of place, in
    place (staged), dsize 2 (digit groups folded into the collapsed key) and a base2k whose integers leave the collapsed-key bound (every
    ciphertext flagged on the device -> the limb-wise sequence redoes the batch from the untouched input)."""
    g, o = pb.Module(n, fl), O.OracleModule(n, fl)
    rng = np.random.default_rng(700 + n + fl)
    batch = 3
    if fl == pb.NTT120:
        cases = ((1, 18, 3, 4, 1, True), (2, 18, 3, 4, 1, True), (1, 18, 4, 4, 2, True), (1, 30, 3, 3, 1, False))
    else:  # FFT64: no bound to flag; rank 2 at n = 4096 does not fit the kernel's shared memory (either route must agree)
        cases = ((1, 12, 3, 4, 1, True), (2, 12, 3, 4, 1, True if n == 1024 else None), (1, 12, 4, 4, 2, True))
    for rank, k, size, key_size, dsize, fused in cases:
        dnum = -(-size // dsize)
        pg, po = _key(g, o, rng, dnum, rank, rank + 1, key_size, k)
        for op, p in ((0, 5), (1, -1), (2, 2 * n - 3), (0, 2 * n - 1)):
            a = fill_uniform(rng, (batch, size, rank + 1, n), k)
            want = fill_uniform(rng, (batch, size, rank + 1, n), k)
            res_g, a_g = g.vec_znx_from_numpy(want), g.vec_znx_from_numpy(a)
            sc = g.glwe_automorphism_op(op, res_g, k, a_g, pg, k, p, dsize)
            g.sync()
            l0 = g.launch_count
            g.glwe_automorphism_op(op, res_g, k, a_g, pg, k, p, dsize, sc)
            g.sync()
            launches = g.launch_count - l0
            assert fused is None or (launches <= 6) == fused, (rank, k, dsize, launches)
            for bi in range(batch):
                o.glwe_automorphism_op(op, want[bi], k, a[bi], po, k, p, dsize)
            assert np.array_equal(g.vec_znx_to_numpy(res_g), want), (rank, k, dsize, op, p)
            assert np.array_equal(g.vec_znx_to_numpy(a_g), a)
            g.glwe_automorphism_op(op, a_g, k, a_g, pg, k, p, dsize, sc)  # in place
            g.sync()
            assert np.array_equal(g.vec_znx_to_numpy(a_g), want), ("assign", rank, k, dsize, op, p)


@pytest.mark.parametrize("fl", FLAVOURS)
def test_glwe_trace_single_kernel_rounds(fl):
    """Trace at n = 1024: ten rounds of rsh + fused automorphism_add alternating between two buffers."""
    n, batch, log_n, k = 1024, 2, 10, 18 if fl == pb.NTT120 else 12
    g, o = pb.Module(n, fl), O.OracleModule(n, fl)
    rng = np.random.default_rng(710 + fl)
    for rank, skip in ((1, 0), (2, 3)):
        keys = [_key(g, o, rng, 3, rank, rank + 1, 4, k) for _ in range(log_n)]
        want = fill_uniform(rng, (batch, 3, rank + 1, n), k)
        res_g = g.vec_znx_from_numpy(want)
        l0 = g.launch_count
        g.glwe_trace_assign(res_g, k, skip, [x[0] for x in keys], k, 1)
        g.sync()
        assert g.launch_count - l0 <= (log_n - skip) * (6 + rank + 1) + 2
        for bi in range(batch):
            o.glwe_trace_assign(want[bi], k, skip, [x[1] for x in keys], k, 1)
        assert np.array_equal(g.vec_znx_to_numpy(res_g), want), (rank, skip)
