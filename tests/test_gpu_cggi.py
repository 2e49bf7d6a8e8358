"""GPU parity of the batched CGGI blind rotation (block-binary) against the oracle, both flavours, bit-exact."""
import ctypes as C

import numpy as np
import pytest

import poulpy_b200 as pb
from oracle import pyoracle as O
from util import fill_uniform

pytestmark = pytest.mark.gpu


def _setup(n, fl, rank, dnum, brk_size, n_lwe, k, rng, trivial_secret=None):
    g, o = pb.Module(n, fl), O.OracleModule(n, fl)
    cols = rank + 1
    per = n * dnum * cols * cols * brk_size * g.prep_bytes
    gbuf = pb.DevBuf(per * n_lwe)
    obrk = []
    for i in range(n_lwe):
        if trivial_secret is None:
            mat = fill_uniform(rng, (dnum, cols, brk_size, cols, n), k)
        else:
            mat = np.zeros((dnum, cols, brk_size, cols, n), dtype=np.int64)
            for d in range(dnum):
                for c in range(cols):
                    mat[d, c, d, c, 0] = trivial_secret[i]
        pm = o.vmp_pmat_alloc(dnum, cols, cols, brk_size)
        o.vmp_prepare(pm, mat)
        obrk.append(pm)
        gp = pb.hal.VmpPMat(gbuf, n, dnum, cols, cols, brk_size, offset=i * per)
        g.vmp_prepare(gp, g.mat_znx_from_numpy(mat))
    gbrk = pb.hal.VmpPMat(gbuf, n, dnum, cols, cols, brk_size)
    return g, o, gbrk, obrk


@pytest.mark.parametrize("fl", [pb.NTT120, pb.FFT64])
@pytest.mark.parametrize("n", [128, 256, 512, 1024])
@pytest.mark.parametrize("rank,dnum,size,brk_size", [(1, 1, 1, 2), (2, 2, 2, 3), (3, 1, 1, 2), (1, 2, 2, 2), (1, 1, 3, 4)])
def test_blind_rotate_matches_oracle(fl, n, rank, dnum, size, brk_size):
    """n >= 256 with cols*dnum <= 8 and cols*brk_size <= 8 takes the single-kernel fused path in the FFT64 flavour; the other
    shapes (and NTT120) run the batched HAL sequence.  batch = 5 leaves a partially filled CTA in the fused kernel."""
    k, n_lwe, block, batch = 12, 12, 3, 5
    rng = np.random.default_rng(100 * rank + dnum + fl)
    g, o, gbrk, obrk = _setup(n, fl, rank, dnum, brk_size, n_lwe, k, rng)
    xg, xo = g.cggi_x_pow_a(), o.cggi_x_pow_a()
    lut = fill_uniform(rng, (size, 1, n), k)
    lwe = rng.integers(-n, n, size=(batch, n_lwe + 1), dtype=np.int64)
    want = np.zeros((batch, size, rank + 1, n), dtype=np.int64)
    for b in range(batch):
        o.cggi_blind_rotate_block_binary(want[b], lwe[b], lut, obrk, xo, block, k)
    res = g.vec_znx_from_numpy(fill_uniform(rng, want.shape, k))  # garbage pre-fill
    lwe_dev = pb.DevBuf(lwe.nbytes)
    lwe_dev.upload(lwe)
    g.cggi_blind_rotate(res, lwe_dev, n_lwe, g.vec_znx_from_numpy(lut), gbrk, xg, block, k)
    g.sync()
    assert np.array_equal(g.vec_znx_to_numpy(res), want)


@pytest.mark.parametrize("fl", [pb.NTT120, pb.FFT64])
def test_x_pow_a_table(fl):
    n = 64
    g, o = pb.Module(n, fl), O.OracleModule(n, fl)
    xg, xo = g.cggi_x_pow_a(), o.cggi_x_pow_a()
    if fl == pb.NTT120:
        got = xg.buf.download(np.uint32, (2 * n, 4, n))
        for k, q in enumerate(O.Q):
            # the oracle's SvpPPol is q120c: (r, r * 2^32 mod q) packed in one u64 (reference/ntt120/types.rs:216)
            assert np.array_equal(got[:, k, :].astype(np.uint64), xo[:, :, k] & np.uint64(0xFFFFFFFF))
    else:
        got = xg.buf.download(np.float64, (2 * n, n))
        assert np.max(np.abs(got - xo)) < 1e-12


@pytest.mark.parametrize("fl", [pb.NTT120, pb.FFT64])
def test_blind_rotate_bench_shape_semantics(fl):
    """BASELINE config 3 shape (n=512, rank=3, block=3, base2k=18, k_brk=36, dnum=1, k_glwe=18) with noiseless trivial keys and a
    shortened LWE dimension: the result must be X^(b + <a, s>) * LUT exactly (L4-style functional check)."""
    n, k, n_lwe, block, batch, rank = 512, 18, 24, 3, 7, 3
    rng = np.random.default_rng(77 + fl)
    s = np.zeros(n_lwe, dtype=np.int64)
    for b0 in range(0, n_lwe, block):
        s[b0 + rng.integers(0, block)] = rng.integers(0, 2)
    g, o, gbrk, obrk = _setup(n, fl, rank, 1, 2, n_lwe, k, rng, trivial_secret=s)
    xg = g.cggi_x_pow_a()
    lut = fill_uniform(rng, (1, 1, n), k - 1)
    lwe = rng.integers(-n, n, size=(batch, n_lwe + 1), dtype=np.int64)
    res = g.vec_znx_alloc(rank + 1, 1, batch)
    lwe_dev = pb.DevBuf(lwe.nbytes)
    lwe_dev.upload(lwe)
    g.cggi_blind_rotate(res, lwe_dev, n_lwe, g.vec_znx_from_numpy(lut), gbrk, xg, block, k)
    g.sync()
    got = g.vec_znx_to_numpy(res)
    for b in range(batch):
        shift = int(lwe[b, 0] + np.dot(lwe[b, 1:], s))
        want = np.zeros((1, rank + 1, n), dtype=np.int64)
        O.vec_znx_rotate(shift, want, 0, lut, 0)
        assert np.array_equal(got[b], want), b


@pytest.mark.parametrize("base2k,size,two_n", [(18, 1, 1024), (5, 3, 1024), (7, 2, 4096), (12, 2, 2048), (11, 1, 1024)])
@pytest.mark.parametrize("rot_left", [True, False])
def test_mod_switch_2n_device(base2k, size, two_n, rot_left):
    """mod_switch_2n (algorithms/mod.rs:136-181) on the device, both branches (rounding when base2k > log2(2N)+1, limb concatenation
    otherwise), bit-exact against the oracle for a batch."""
    n_lwe, batch = 37, 6
    rng = np.random.default_rng(base2k * 100 + size)
    g = pb.Module(64, pb.FFT64)
    lwe = fill_uniform(rng, (batch, size, 1, n_lwe + 1), base2k)
    dev = pb.DevBuf(lwe.nbytes)
    dev.upload(lwe)
    out = g.cggi_mod_switch_2n(dev, batch, n_lwe, size, base2k, two_n, rot_left)
    g.sync()
    got = out.download(np.int64, (batch, n_lwe + 1))
    for b in range(batch):
        assert np.array_equal(got[b], O.mod_switch_2n(two_n, lwe[b], base2k, rot_left)), b


@pytest.mark.parametrize("fl", [pb.NTT120, pb.FFT64])
@pytest.mark.parametrize("n,rank,dnum,size,brk_size", [(256, 1, 2, 2, 2), (512, 2, 1, 1, 2), (1024, 1, 3, 3, 3), (128, 1, 2, 2, 3)])
def test_blind_rotate_standard_matches_oracle(fl, n, rank, dnum, size, brk_size):
    """execute_standard (algorithm.rs:370-443): external product + (X^a - 1) + add per LWE coefficient, one normalize at the end;
    random (non-cryptographic) GGSWs, bit-exact against the oracle.  n = 1024 routes the external product through the
    single-kernel gadget path in the NTT120 flavour."""
    k, n_lwe, batch = 12, 7, 4
    rng = np.random.default_rng(300 * rank + dnum + fl + n)
    g, o, gbrk, obrk = _setup(n, fl, rank, dnum, brk_size, n_lwe, k, rng)
    lut = fill_uniform(rng, (size, 1, n), k)
    lwe = rng.integers(-n, n, size=(batch, n_lwe + 1), dtype=np.int64)
    want = np.zeros((batch, size, rank + 1, n), dtype=np.int64)
    for b in range(batch):
        o.cggi_blind_rotate_standard(want[b], k, lwe[b], lut, obrk, k)
    res = g.vec_znx_from_numpy(fill_uniform(rng, want.shape, k))  # garbage pre-fill
    lwe_dev = pb.DevBuf(lwe.nbytes)
    lwe_dev.upload(lwe)
    g.cggi_blind_rotate_standard(res, k, lwe_dev, n_lwe, g.vec_znx_from_numpy(lut), gbrk, k)
    g.sync()
    assert np.array_equal(g.vec_znx_to_numpy(res), want)


@pytest.mark.parametrize("fl", [pb.NTT120, pb.FFT64])
def test_blind_rotate_standard_semantics(fl):
    """Noiseless binary keys: execute_standard on mod-switched LWEs (device mod_switch_2n) gives X^(b + <a, s>) * LUT exactly."""
    n, k, n_lwe, batch, rank = 512, 18, 16, 5, 1
    rng = np.random.default_rng(177 + fl)
    s = rng.integers(0, 2, size=n_lwe, dtype=np.int64)
    g, o, gbrk, obrk = _setup(n, fl, rank, 2, 2, n_lwe, k, rng, trivial_secret=s)
    lut = fill_uniform(rng, (2, 1, n), k - 1)
    lwe_raw = fill_uniform(rng, (batch, 1, 1, n_lwe + 1), k)
    dev = pb.DevBuf(lwe_raw.nbytes)
    dev.upload(lwe_raw)
    lwe_dev = g.cggi_mod_switch_2n(dev, batch, n_lwe, 1, k, 2 * n, True)
    res = g.vec_znx_alloc(rank + 1, 2, batch)
    g.cggi_blind_rotate_standard(res, k, lwe_dev, n_lwe, g.vec_znx_from_numpy(lut), gbrk, k)
    g.sync()
    got = g.vec_znx_to_numpy(res)
    for b in range(batch):
        l2 = O.mod_switch_2n(2 * n, lwe_raw[b], k, True)
        shift = int(l2[0] + np.dot(l2[1:], s))
        want = np.zeros((2, rank + 1, n), dtype=np.int64)
        O.vec_znx_rotate(shift, want, 0, lut, 0)
        assert np.array_equal(got[b], want), b


@pytest.mark.parametrize("fl", [pb.NTT120, pb.FFT64])
@pytest.mark.parametrize("ext", [2, 4])
def test_blind_rotate_extended_matches_oracle(fl, ext):
    """execute_block_binary_extended (algorithm.rs:121-273): `ext` interleaved accumulator rings, source ring and X^a table entry chosen per
    ciphertext on the device; random keys and LWEs (including the positions where the reference skips the cross-ring update), batch 5."""
    n, n_lwe, block, rank, batch = 128, 9, 3, 1, 5
    k = 12 if fl == pb.FFT64 else 18
    dnum, brk_size, size = 1, 2, 2
    cols = rank + 1
    g, o = pb.Module(n, fl), O.OracleModule(n, fl)
    rng = np.random.default_rng(70 + ext + fl)
    per = n * cols * cols * brk_size * dnum * g.prep_bytes
    brk_buf = pb.DevBuf(per * n_lwe)
    brk_o = []
    for i in range(n_lwe):
        mt = fill_uniform(rng, (dnum, cols, brk_size, cols, n), k)
        g.vmp_prepare(pb.hal.VmpPMat(brk_buf, n, dnum, cols, cols, brk_size, offset=i * per), g.mat_znx_from_numpy(mt))
        pm = o.vmp_pmat_alloc(dnum, cols, cols, brk_size)
        o.vmp_prepare(pm, mt)
        brk_o.append(pm)
    luts = fill_uniform(rng, (ext, size, 1, n), k - 1)
    N = n * ext
    lwe_2n = rng.integers(-N, N, size=(batch, n_lwe + 1), dtype=np.int64)
    lwe_2n[1, 1:4] = [1, 2 * N - 1, ext]  # a_hi = 0 with a_lo != 0, a_hi = 2n - 1, and a pure in-ring rotation
    lwe_dev = pb.DevBuf(lwe_2n.nbytes)
    lwe_dev.upload(lwe_2n)
    want = fill_uniform(rng, (batch, size, cols, n), k)
    res_g = g.vec_znx_from_numpy(want)
    g.cggi_blind_rotate_extended(res_g, lwe_dev, n_lwe, g.vec_znx_from_numpy(luts), ext, pb.hal.VmpPMat(brk_buf, n, dnum, cols, cols, brk_size),
                                 g.cggi_x_pow_a(), block, k)
    g.sync()
    xpa = o.cggi_x_pow_a()
    for b in range(batch):
        o.cggi_blind_rotate_block_binary_extended(want[b], lwe_2n[b], [luts[j] for j in range(ext)], brk_o, xpa, block, k)
    assert np.array_equal(g.vec_znx_to_numpy(res_g), want)


@pytest.mark.parametrize("fl", [pb.NTT120, pb.FFT64])
@pytest.mark.parametrize("rank,dnum,size,brk_size", [(2, 3, 4, 4), (4, 2, 2, 2), (3, 3, 2, 3), (3, 4, 3, 4)])
def test_blind_rotate_tall_keys(fl, rank, dnum, size, brk_size):
    """cols * dnum = 9, 10, 12, 16 input polys: the row tiles of cggi_block_ntt120_kernel (RMAX = 9 / 10 / 12 / 16) and, in FFT64, the block
    kernel with two ciphertexts per thread + fft64_back_kernel (these shapes are outside the fully fused kernel); (2, 3, 4, 4) is the
    circuit-bootstrapping bench layout.  batch = 3 leaves a half-filled pair."""
    n, k, n_lwe, block, batch = 512, 12, 8, 4, 3  # n >= 512: the NTT120 fused back end (ntt120_fused_supported)
    rng = np.random.default_rng(4100 + 10 * rank + dnum + fl)
    g, o, gbrk, obrk = _setup(n, fl, rank, dnum, brk_size, n_lwe, k, rng)
    xg, xo = g.cggi_x_pow_a(), o.cggi_x_pow_a()
    lut = fill_uniform(rng, (size, 1, n), k)
    lwe = rng.integers(-n, n, size=(batch, n_lwe + 1), dtype=np.int64)
    want = np.zeros((batch, size, rank + 1, n), dtype=np.int64)
    for b in range(batch):
        o.cggi_blind_rotate_block_binary(want[b], lwe[b], lut, obrk, xo, block, k)
    res = g.vec_znx_from_numpy(fill_uniform(rng, want.shape, k))
    lwe_dev = pb.DevBuf(lwe.nbytes)
    lwe_dev.upload(lwe)
    l0 = g.launch_count
    g.cggi_blind_rotate(res, lwe_dev, n_lwe, g.vec_znx_from_numpy(lut), gbrk, xg, block, k)
    g.sync()
    if size >= dnum:  # (accumulators shorter than the key's row count transform column by column and zero-fill)
        assert g.launch_count - l0 <= 2 + 3 * (n_lwe // block)  # init + per block: forward, block kernel, fused back end
    assert np.array_equal(g.vec_znx_to_numpy(res), want)


@pytest.mark.parametrize("n,rank,size,brk_size,batch", [(512, 3, 1, 2, 9), (512, 3, 1, 2, 16), (256, 1, 1, 2, 23), (1024, 1, 1, 2, 5)])
def test_blind_rotate_cluster_multicast(n, rank, size, brk_size, batch):
    """OPT_CGGI_CLUSTER = 2: the FFT64 whole-rotation kernel runs in clusters of two CTAs whose key tiles are fetched once by rank 0 and
    multicast into both rings.  Same operations per value, so the results are those of the plain launch bit for bit and the oracle's;
    batches that leave a partially filled CTA, an odd number of CTAs (the padding CTA walks the ring with no live ciphertext) and a single
    cluster are covered."""
    k, n_lwe, block = 12, 24, 3
    rng = np.random.default_rng(7000 + n + batch)
    g, o, gbrk, obrk = _setup(n, pb.FFT64, rank, 1, brk_size, n_lwe, k, rng)
    xg, xo = g.cggi_x_pow_a(), o.cggi_x_pow_a()
    lut = fill_uniform(rng, (size, 1, n), k)
    lwe = rng.integers(-n, n, size=(batch, n_lwe + 1), dtype=np.int64)
    want = np.zeros((batch, size, rank + 1, n), dtype=np.int64)
    for b in range(batch):
        o.cggi_blind_rotate_block_binary(want[b], lwe[b], lut, obrk, xo, block, k)
    lwe_dev = pb.DevBuf(lwe.nbytes)
    lwe_dev.upload(lwe)
    got = []
    for cl in (2, 0):
        g.set_option(pb.hal.OPT_CGGI_CLUSTER, cl)
        res = g.vec_znx_from_numpy(fill_uniform(rng, want.shape, k))
        g.cggi_blind_rotate(res, lwe_dev, n_lwe, g.vec_znx_from_numpy(lut), gbrk, xg, block, k)
        g.sync()
        got.append(g.vec_znx_to_numpy(res).reshape(want.shape))
    assert np.array_equal(got[0], want) and np.array_equal(got[1], want)
