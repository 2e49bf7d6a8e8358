"""GPU parity of the bivariate convolution (HalImpl::cnv_*) and of the CKKS multiplication halves glwe_tensor_apply /
glwe_tensor_relinearize against the oracle.  Comparison points: after idft + big_normalize (normalised VecZnx,
bit-exact in both flavours) and the tensor / relinearised ciphertext columns themselves."""
import numpy as np
import pytest

import poulpy_b200 as pb
from oracle import pyoracle as O
from util import fill_uniform

pytestmark = pytest.mark.gpu
FLAVOURS = [pb.NTT120, pb.FFT64]


def _norm_g(g, res_size, k, big):
    out = g.vec_znx_alloc(1, res_size)
    g.vec_znx_big_normalize(out, k, 0, 0, big, k, 0)
    return g.vec_znx_to_numpy(out)


def _norm_o(o, res_size, k, big):
    out = np.zeros((res_size, 1, o.n), dtype=np.int64)
    o.vec_znx_big_normalize(out, k, 0, 0, big, k, 0)
    return out


@pytest.mark.parametrize("fl", FLAVOURS)
@pytest.mark.parametrize("n", [64, 1024])
def test_cnv_apply_pairwise_by_const(fl, n):
    """test_suite/convolution.rs:21-245 shapes scaled down: every cnv_offset, both operand columns, pairwise sums, masked prepare,
    results shorter and longer than a.size + b.size - 1."""
    g, o = pb.Module(n, fl), O.OracleModule(n, fl)
    rng = np.random.default_rng(71 + n + fl)
    k = 12 if fl == pb.FFT64 else 30
    a_size, b_size = 4, 3
    a, b = fill_uniform(rng, (a_size, 3, n), k), fill_uniform(rng, (b_size, 3, n), k)
    mask = -1 << 3
    apg, bpg = g.cnv_pvec_alloc(3, a_size), g.cnv_pvec_alloc(3, b_size)
    g.cnv_prepare_left(apg, g.vec_znx_from_numpy(a), mask)
    g.cnv_prepare_right(bpg, g.vec_znx_from_numpy(b), -1)
    apo, bpo = o.vec_znx_dft_alloc(3, a_size), o.vec_znx_dft_alloc(3, b_size)
    o.cnv_prepare(apo, a, mask)
    o.cnv_prepare(bpo, b, -1)
    for res_size in (2, a_size + b_size - 1, a_size + b_size + 1):
        for off in range(0, a_size + b_size + 1, 2):
            for (i, j) in ((0, 0), (2, 1), (0, 2)):
                rg, ro = g.vec_znx_dft_alloc(1, res_size), o.vec_znx_dft_alloc(1, res_size)
                rg.buf.upload(rng.integers(0, 255, rg.buf.nbytes, dtype=np.uint8))
                if i == j:
                    g.cnv_apply_dft(off, rg, 0, apg, i, bpg, j)
                    o.cnv_apply_dft(off, ro, 0, apo, i, bpo, j)
                else:
                    g.cnv_pairwise_apply_dft(off, rg, 0, apg, bpg, i, j)
                    o.cnv_pairwise_apply_dft(off, ro, 0, apo, bpo, i, j)
                got = _norm_g(g, res_size, k, g.vec_znx_idft_apply_consume(rg))
                want = _norm_o(o, res_size, k, o.vec_znx_idft_apply_consume(ro))
                assert np.array_equal(got, want), (res_size, off, i, j)
            bc = fill_uniform(rng, (b_size,), k)
            bg, bo = g.vec_znx_big_alloc(1, res_size), o.vec_znx_big_alloc(1, res_size)
            g.cnv_by_const_apply(off, bg, 0, g.vec_znx_from_numpy(a), 1, bc)
            o.cnv_by_const_apply(off, bo, 0, a, 1, bc)
            assert np.array_equal(_norm_g(g, res_size, k, bg), _norm_o(o, res_size, k, bo)), ("const", res_size, off)
    # prepare_self = prepare_left + prepare_right
    lg, rg2 = g.cnv_pvec_alloc(3, a_size), g.cnv_pvec_alloc(3, a_size)
    g.cnv_prepare_self(lg, rg2, g.vec_znx_from_numpy(a), mask)
    rd, ro = g.vec_znx_dft_alloc(1, 2 * a_size), o.vec_znx_dft_alloc(1, 2 * a_size)
    g.cnv_apply_dft(1, rd, 0, lg, 1, rg2, 2)
    o.cnv_apply_dft(1, ro, 0, apo, 1, apo, 2)
    assert np.array_equal(_norm_g(g, 2 * a_size, k, g.vec_znx_idft_apply_consume(rd)), _norm_o(o, 2 * a_size, k, o.vec_znx_idft_apply_consume(ro)))


@pytest.mark.parametrize("fl", FLAVOURS)
@pytest.mark.parametrize("rank,n", [(1, 256), (2, 64), (1, 2048)])
def test_glwe_tensor_apply_and_relinearize(fl, rank, n):
    """ckks_mul_into = glwe_tensor_apply + glwe_tensor_relinearize (poulpy-ckks/src/leveled/default/mul.rs:49-86), batched, for several
    cnv_offsets (below / above one limb), mixed res base2k, and the relinearisation with equal and different key base2k."""
    g, o = pb.Module(n, fl), O.OracleModule(n, fl)
    rng = np.random.default_rng(91 + rank + n + fl)
    ab = 12 if fl == pb.FFT64 else 26
    cols, pairs, size, batch = rank + 1, rank * (rank + 1) // 2, 3, 3
    tcols = cols * (cols + 1) // 2
    a, b = fill_uniform(rng, (batch, size, cols, n), ab), fill_uniform(rng, (batch, size, cols, n), ab)
    for cnv_offset, res_k in ((ab - 3, ab), (ab + 5, ab), (2 * ab, ab - 1)):
        want = fill_uniform(rng, (batch, size, tcols, n), ab)
        res_g = g.vec_znx_from_numpy(want)  # garbage pre-fill on both sides
        g.glwe_tensor_apply(cnv_offset, res_g, res_k, g.vec_znx_from_numpy(a), size * ab, g.vec_znx_from_numpy(b), size * ab - 2, ab)
        g.sync()
        for bi in range(batch):
            o.glwe_tensor_apply(cnv_offset, want[bi], res_k, a[bi], size * ab, b[bi], size * ab - 2, ab)
        assert np.array_equal(g.vec_znx_to_numpy(res_g), want), (cnv_offset, res_k)
    # relinearisation of the last tensor
    tensor = want
    for key_k, res_k2, t_k, dsize in ((ab, ab, ab, 1), (ab + 2, ab, ab, 1), (ab, ab, ab, 2)):
        key_size = size + 1
        dnum = -(-(-(-size * t_k // key_k)) // dsize)
        mat = fill_uniform(rng, (dnum, pairs, key_size, cols, n), key_k)
        pg, po = g.vmp_pmat_alloc(dnum, pairs, cols, key_size), o.vmp_pmat_alloc(dnum, pairs, cols, key_size)
        g.vmp_prepare(pg, g.mat_znx_from_numpy(mat))
        o.vmp_prepare(po, mat)
        want2 = fill_uniform(rng, (batch, size, cols, n), ab)
        rg = g.vec_znx_from_numpy(want2)
        g.glwe_tensor_relinearize(rg, res_k2, g.vec_znx_from_numpy(tensor), t_k, pg, key_k, dsize)
        g.sync()
        for bi in range(batch):
            o.glwe_tensor_relinearize(want2[bi], res_k2, tensor[bi], t_k, po, key_k, dsize)
        assert np.array_equal(g.vec_znx_to_numpy(rg), want2), (key_k, res_k2, dsize)

