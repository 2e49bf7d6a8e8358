"""Bit-exact GPU-vs-oracle parity AT THE CONFIGURATIONS bench.py TIMES ("next round" item 1): every number the bench
prints is for a computation these tests compare with the oracle, with random non-trivial keys.

  * CGGI blind rotation, poulpy-bench/src/bench_suite/schemes/blind_rotation.rs:39-72: n=512, n_lwe=687, rank 3, block 3, base2k 18,
    k_brk=36 (2 limbs), dnum 1, k_glwe=18 (1 limb), batch 2368 (the bench batch), both flavours -- 229 blocks through the fused FFT64
    kernel's TMA ring (hundreds of mbarrier phase wraps) and through the NTT120 whole-rotation kernel.
  * circuit-bootstrapping blind rotation, poulpy-bench/src/bench_suite/schemes/circuit_bootstrapping.rs:47-129: n=1024, n_lwe=574,
    block 7, rank 2, base2k 13, BRK k=52 (4 limbs) dnum 3, batch 64, both flavours.
  * CKKS ct x ct multiplication, poulpy-bench/src/bench_suite/ckks.rs:31-37: N=2^15, base2k 52, K=728 (14 limbs), tensor key 15 limbs
    (NTT120 only, as the reference), batch 1: glwe_tensor_apply + glwe_tensor_relinearize.

The oracle computes a handful of distinct inputs (seconds); the GPU batch replicates them in shuffled order so every ciphertext slot of
every CTA / wave is checked."""
import numpy as np
import pytest

import poulpy_b200 as pb
from oracle import pyoracle as O
from util import fill_uniform

pytestmark = pytest.mark.gpu


def _brk(g, o, n, cols, dnum, size, n_lwe, k, rng):
    per = n * dnum * cols * cols * size * g.prep_bytes
    buf = pb.DevBuf(per * n_lwe)
    obrk = []
    for i in range(n_lwe):
        mat = fill_uniform(rng, (dnum, cols, size, cols, n), k)  # uniform k-bit digits in every entry: no structure to hide behind
        pm = o.vmp_pmat_alloc(dnum, cols, cols, size)
        o.vmp_prepare(pm, mat)
        obrk.append(pm)
        g.vmp_prepare(pb.hal.VmpPMat(buf, n, dnum, cols, cols, size, offset=i * per), g.mat_znx_from_numpy(mat))
    return pb.hal.VmpPMat(buf, n, dnum, cols, cols, size), obrk


def _blind_rotate_case(fl, n, n_lwe, rank, block, k, dnum, brk_size, acc_size, lut_size, distinct, batch, seed):
    rng = np.random.default_rng(seed + fl)
    g, o = pb.Module(n, fl), O.OracleModule(n, fl)
    cols = rank + 1
    gbrk, obrk = _brk(g, o, n, cols, dnum, brk_size, n_lwe, k, rng)
    xg, xo = g.cggi_x_pow_a(), o.cggi_x_pow_a()
    lut = fill_uniform(rng, (lut_size, 1, n), k)
    lwe = rng.integers(-n, n, size=(distinct, n_lwe + 1), dtype=np.int64)
    lwe[0, 1:9] = [0, 2 * n - 1, -(2 * n - 1), n, -n, 1, -1, 0]  # table corners: X^0 (identity), X^(2n-1), X^n = -1 (inputs outside
    # [-n, n) are legal: the kernels reduce them mod 2n like x_pow_a's index, algorithm.rs:349)
    want = np.zeros((distinct, acc_size, cols, n), dtype=np.int64)
    for b in range(distinct):
        o.cggi_blind_rotate_block_binary(want[b], lwe[b], lut, obrk, xo, block, k)
    idx = rng.permutation(np.arange(batch) % distinct)
    res = g.vec_znx_from_numpy(fill_uniform(rng, (batch, acc_size, cols, n), k))  # garbage pre-fill
    lwe_dev = pb.DevBuf(batch * (n_lwe + 1) * 8)
    lwe_dev.upload(lwe[idx])
    l0 = g.launch_count
    g.cggi_blind_rotate(res, lwe_dev, n_lwe, g.vec_znx_from_numpy(lut), gbrk, xg, block, k)
    g.sync()
    launches = g.launch_count - l0
    got = g.vec_znx_to_numpy(res)
    bad = [int(b) for b in range(batch) if not np.array_equal(got[b], want[idx[b]])]
    assert not bad, f"{len(bad)} of {batch} ciphertexts differ from the oracle (first: {bad[:8]})"
    assert np.abs(want).max() < (1 << (k - 1)) + 1 and np.abs(want).max() > (1 << (k - 3))  # the outputs are real digits, not zeros
    return launches


@pytest.mark.parametrize("fl", [pb.FFT64, pb.NTT120])
def test_cggi_blind_rotate_baseline_config(fl):
    """BASELINE.json config 4 exactly as bench.py times it (batch 2368 = 148 SMs x 4 ciphertexts x 4 waves), 16 distinct LWEs."""
    _blind_rotate_case(fl, n=512, n_lwe=687, rank=3, block=3, k=18, dnum=1, brk_size=2, acc_size=1, lut_size=1, distinct=16, batch=2368, seed=4000)


@pytest.mark.parametrize("fl", [pb.FFT64, pb.NTT120])
def test_cggi_blind_rotate_baseline_config_ragged_batch(fl):
    """Same configuration with a batch that leaves the last CTA / cluster partially filled and a shorter LWE dimension that is not a
    multiple of the block size (chunks_exact drops the tail, algorithm.rs:338)."""
    _blind_rotate_case(fl, n=512, n_lwe=62, rank=3, block=3, k=18, dnum=1, brk_size=2, acc_size=1, lut_size=1, distinct=5, batch=149 * 4 + 3, seed=4010)


@pytest.mark.parametrize("fl", [pb.FFT64, pb.NTT120])
def test_circuit_bootstrap_blind_rotate_bench_shape(fl):
    """The blind rotation inside the circuit-bootstrapping bench: R = 9 input polys, C = 12 output polys, 82 blocks of 7, accumulator in
    the BRK layout (4 limbs), 2-limb LUT."""
    _blind_rotate_case(fl, n=1024, n_lwe=574, rank=2, block=7, k=13, dnum=3, brk_size=4, acc_size=4, lut_size=2, distinct=4, batch=64, seed=4020)


def test_ckks_mul_bench_shape():
    """ckks_mul_into at the reference's bench parameters, one ciphertext pair: tensor (3 columns x 14 limbs) and relinearised result
    compared bit for bit.  Digits are uniform 52-bit values, so the 14-term limb convolutions exceed Q/2 and wrap modulo Q exactly as in
    the reference (i128 arithmetic modulo Q, centred): the wrap is part of what is compared."""
    n, k, size = 1 << 15, 52, 14
    rng = np.random.default_rng(4030)
    g, o = pb.Module(n, pb.NTT120), O.OracleModule(n, O.NTT120)
    a, b = fill_uniform(rng, (1, size, 2, n), k), fill_uniform(rng, (1, size, 2, n), k)
    want_t = np.zeros((1, size, 3, n), dtype=np.int64)
    o.glwe_tensor_apply(size * k, want_t[0], k, a[0], size * k, b[0], size * k, k)
    tg = g.vec_znx_from_numpy(fill_uniform(rng, want_t.shape, k))
    g.glwe_tensor_apply(size * k, tg, k, g.vec_znx_from_numpy(a), size * k, g.vec_znx_from_numpy(b), size * k, k)
    g.sync()
    assert np.array_equal(g.vec_znx_to_numpy(tg), want_t[0])  # (a batch of one downloads without the batch axis)
    mat = fill_uniform(rng, (size, 1, size + 1, 2, n), k)
    pg, po = g.vmp_pmat_alloc(size, 1, 2, size + 1), o.vmp_pmat_alloc(size, 1, 2, size + 1)
    g.vmp_prepare(pg, g.mat_znx_from_numpy(mat))
    o.vmp_prepare(po, mat)
    want_r = np.zeros((1, size, 2, n), dtype=np.int64)
    o.glwe_tensor_relinearize(want_r[0], k, want_t[0], k, po, k, 1)
    rg = g.vec_znx_from_numpy(fill_uniform(rng, want_r.shape, k))
    g.glwe_tensor_relinearize(rg, k, tg, k, pg, k, 1)
    g.sync()
    assert np.array_equal(g.vec_znx_to_numpy(rg), want_r[0])


@pytest.mark.parametrize("fl", [pb.FFT64, pb.NTT120])
@pytest.mark.parametrize("ext", [False, True])
def test_keyswitch_and_external_product_bench_shapes_base2k18(fl, ext):
    """C1 (n=4096, key 3 x 4 limbs) and C3 (n=2048, GGSW 3 rows x 2 cols x 3 limbs) at the bench base2k = 18 in BOTH flavours: the FFT64
    instances at base2k 18 are the regime where a contracted (FMA) transform and the reference's non-contracted one could round
    differently if the pre-rounding error approached 1/2 -- it does not (|values| < 2^50, error ~ 2^-5), so the rounded bigs are the exact
    integers on both sides."""
    n, k = (2048, 18) if ext else (4096, 18)
    cols_in, key_size, batch = (2, 3, 37) if ext else (1, 4, 37)
    rng = np.random.default_rng(4040 + fl + 2 * ext)
    g, o = pb.Module(n, fl), O.OracleModule(n, fl)
    mat = fill_uniform(rng, (3, cols_in, key_size, 2, n), k)
    pg, po = g.vmp_pmat_alloc(3, cols_in, 2, key_size), o.vmp_pmat_alloc(3, cols_in, 2, key_size)
    g.vmp_prepare(pg, g.mat_znx_from_numpy(mat))
    o.vmp_prepare(po, mat)
    a = fill_uniform(rng, (batch, 3, 2, n), k)
    want = np.zeros_like(a)
    (o.glwe_external_product_batch if ext else o.glwe_keyswitch_batch)(want, k, a, k, po, k, 1)
    res = g.vec_znx_from_numpy(fill_uniform(rng, a.shape, k))
    (g.glwe_external_product if ext else g.glwe_keyswitch)(res, k, g.vec_znx_from_numpy(a), k, pg, k, 1)
    g.sync()
    assert np.array_equal(g.vec_znx_to_numpy(res), want)


@pytest.mark.parametrize("fl", [pb.FFT64, pb.NTT120])
def test_cggi_blind_rotate_host_front_end(fl):
    """pgb_cggi_blind_rotate_host (bench.py's CGGI e2e leg): host LWEs (base2k digits, not yet mod-switched) in, host GLWEs out, at the
    bench shape with a shortened LWE dimension; batch 1500 spans several whole-wave chunks plus a ragged tail.  Expected values: the oracle's
    mod_switch_2n + execute_block_binary per distinct input."""
    n, n_lwe, rank, block, k, distinct, batch = 512, 30, 3, 3, 18, 6, 1500
    rng = np.random.default_rng(4050 + fl)
    g, o = pb.Module(n, fl), O.OracleModule(n, fl)
    cols = rank + 1
    gbrk, obrk = _brk(g, o, n, cols, 1, 2, n_lwe, k, rng)
    xg, xo = g.cggi_x_pow_a(), o.cggi_x_pow_a()
    lut = fill_uniform(rng, (1, 1, n), k)
    lwe = fill_uniform(rng, (distinct, 1, 1, n_lwe + 1), k)
    want = np.zeros((distinct, 1, cols, n), dtype=np.int64)
    for b in range(distinct):
        o.cggi_blind_rotate_block_binary(want[b], O.mod_switch_2n(2 * n, lwe[b], k, True), lut, obrk, xo, block, k)
    idx = rng.permutation(np.arange(batch) % distinct)
    for pinned in (True, False):
        lwe_h = pb.pinned_empty((batch, 1, 1, n_lwe + 1)) if pinned else np.empty((batch, 1, 1, n_lwe + 1), dtype=np.int64)
        lwe_h[:] = lwe[idx]
        res_h = pb.pinned_empty((batch, 1, cols, n)) if pinned else np.empty((batch, 1, cols, n), dtype=np.int64)
        res_h[:] = -7
        g.cggi_blind_rotate_host(res_h, lwe_h, k, g.vec_znx_from_numpy(lut), gbrk, xg, block, k)
        assert np.array_equal(res_h, want[idx]), pinned


@pytest.mark.parametrize("fl", [pb.FFT64, pb.NTT120])
def test_circuit_bootstrap_bench_shape_whole_circuit(fl):
    """The WHOLE constant-mode circuit bootstrap at the reference's bench shape (poulpy-bench circuit_bootstrapping.rs:47-129 as bench.py
    times it: n=1024, n_lwe=574, block 7, rank 2, base2k 13, BRK / ATK / TSK with 3 rows x 4 limbs, result 2 rows x 2 limbs, 1-bit
    domain) with uniform random key digits: the device orchestration against the independent restatement of circuit_bootstrap_core over
    the oracle (tests/semantics_circuit.py), bit for bit, for three LWEs."""
    import semantics_circuit as SC
    from poulpy_b200 import circuit
    n, log_n, n_lwe, block, rank, k, batch = 1024, 10, 574, 7, 2, 13, 3
    cols, ksz, kd, res_size, dnum_res, log_domain = rank + 1, 4, 3, 2, 2, 1
    rng = np.random.default_rng(4060 + fl)
    g, o = pb.Module(n, fl), O.OracleModule(n, fl)
    gbrk, obrk = _brk(g, o, n, cols, kd, ksz, n_lwe, k, rng)

    def keys(count):
        out_g, out_o = [], []
        for _ in range(count):
            mt = fill_uniform(rng, (kd, rank, ksz, cols, n), k)
            pg, po = g.vmp_pmat_alloc(kd, rank, cols, ksz), o.vmp_pmat_alloc(kd, rank, cols, ksz)
            g.vmp_prepare(pg, g.mat_znx_from_numpy(mt))
            o.vmp_prepare(po, mt)
            out_g.append(pg)
            out_o.append(po)
        return out_g, out_o

    atk_g, atk_o = keys(log_n)
    tsk_g, tsk_o = keys(rank)
    lwe = fill_uniform(rng, (batch, 1, 1, n_lwe + 1), k)
    lwe_dev = pb.DevBuf(lwe.nbytes)
    lwe_dev.upload(lwe)
    ggsw = circuit.circuit_bootstrap_to_constant(g, lwe_dev, batch, n_lwe, 1, k, gbrk, g.cggi_x_pow_a(), block, atk_g, tsk_g, k, rank, dnum_res,
                                                 res_size, log_domain)
    got = ggsw.download(np.int64, (batch, dnum_res, cols, res_size, cols, n))
    xpa = o.cggi_x_pow_a()
    for b in range(batch):
        want = SC.circuit_bootstrap_to_constant_ref(o, lwe[b], k, obrk, xpa, block, atk_o, tsk_o, rank, dnum_res, res_size, log_domain, ksz)
        assert np.array_equal(got[b], want), b
