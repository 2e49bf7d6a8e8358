"""GPU parity tests of the HAL primitives, through the C ABI, against the CPU oracle.

Procedures follow the reference's cross-backend suites (poulpy-hal/src/test_suite/{vec_znx_dft,svp,vmp,vec_znx_big}.rs
instantiated by poulpy-cpu-ref/src/tests.rs:47-141): same call sequence on both "backends", bit-exact comparison of
the backend-independent results (normalised VecZnx), plus stricter intermediate checks:
  L0  NTT120 DFT-domain residues equal the oracle's (mod Q[k], same frequency order)
  L1  VecZnxBig bit-exact (i128 / i64)
  L2  normalised VecZnx bit-exact
  L3  FFT64 DFT domain within 2^-(53 - log_m - 1) * max|x| * growth   (tolerance stated per test)
"""
import numpy as np
import pytest

import poulpy_b200 as pb
from oracle import pyoracle as O
from util import fill_uniform

pytestmark = pytest.mark.gpu

FLAVOURS = [pb.NTT120, pb.FFT64]
_mods = {}


def mods(n, fl):
    key = (n, fl)
    if key not in _mods:
        _mods[key] = (pb.Module(n, fl), O.OracleModule(n, fl))
    return _mods[key]


def base2k_for(fl):
    return 50 if fl == pb.NTT120 else 12  # poulpy-cpu-avx/src/ntt120/tests.rs (50), poulpy-cpu-ref/src/tests.rs (12)


def dft_equal(fl, g_dft_np, o_dft_np, scale=1.0, log_m=8):
    """g: GPU container dump, o: oracle container. NTT120 exact mod Q, FFT64 within tolerance."""
    if fl == pb.NTT120:
        for k, q in enumerate(O.Q):
            assert np.array_equal(g_dft_np[:, :, k, :].astype(np.uint64), o_dft_np[:, :, :, k] % np.uint64(q))
    else:
        tol = 2.0 ** -(53 - log_m - 1) * max(scale, 1.0) * 8
        assert np.max(np.abs(g_dft_np - o_dft_np)) <= tol, (np.max(np.abs(g_dft_np - o_dft_np)), tol)


def big_equal(fl, g_big_np, o_big_np):
    assert np.array_equal(g_big_np, o_big_np)


@pytest.mark.parametrize("fl", FLAVOURS)
@pytest.mark.parametrize("n", [16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536])
def test_dft_idft_roundtrip_all_sizes(fl, n):
    g, o = mods(n, fl)
    rng = np.random.default_rng(n + fl)
    k = base2k_for(fl)
    a = fill_uniform(rng, (2, 2, n), k)
    a_g = g.vec_znx_from_numpy(a)
    dg, do = g.vec_znx_dft_alloc(2, 2), o.vec_znx_dft_alloc(2, 2)
    for c in range(2):
        g.vec_znx_dft_apply(1, 0, dg, c, a_g, c)
        o.vec_znx_dft_apply(1, 0, do, c, a, c)
    dft_equal(fl, g.vec_znx_dft_to_numpy(dg), do, scale=float(n) * 2.0 ** k, log_m=max(n.bit_length() - 2, 1))
    bg, bo = g.vec_znx_big_alloc(2, 2), o.vec_znx_big_alloc(2, 2)
    for c in range(2):
        g.vec_znx_idft_apply(bg, c, dg, c)
        o.vec_znx_idft_apply(bo, c, do, c)
    big_equal(fl, g.vec_znx_big_to_numpy(bg), bo)
    # the round trip is the identity on small inputs
    if fl == pb.NTT120:
        big = g.vec_znx_big_to_numpy(bg)
        assert np.array_equal(big[..., 0].view(np.int64), a) and np.array_equal(big[..., 1].view(np.int64), a >> 63)
    else:
        assert np.array_equal(g.vec_znx_big_to_numpy(bg), a)
    # consume (in place) gives the same big values
    cons = g.vec_znx_idft_apply_consume(dg)
    big_equal(fl, g.vec_znx_big_to_numpy(cons), bo)


@pytest.mark.parametrize("n", [64, 1024])
def test_ntt120_full_i64_range(n):
    """b_from_znx64 is exact for the full i64 range (arithmetic.rs:39-60); so is the GPU reduction."""
    g, o = mods(n, pb.NTT120)
    rng = np.random.default_rng(3)
    a = fill_uniform(rng, (1, 1, n), 64)
    a[0, 0, :4] = [np.iinfo(np.int64).min, np.iinfo(np.int64).max, -1, 0]
    dg, do = g.vec_znx_dft_alloc(1, 1), o.vec_znx_dft_alloc(1, 1)
    g.vec_znx_dft_apply(1, 0, dg, 0, g.vec_znx_from_numpy(a), 0)
    o.vec_znx_dft_apply(1, 0, do, 0, a, 0)
    dft_equal(pb.NTT120, g.vec_znx_dft_to_numpy(dg), do)
    bg, bo = g.vec_znx_big_alloc(1, 1), o.vec_znx_big_alloc(1, 1)
    g.vec_znx_idft_apply(bg, 0, dg, 0)
    o.vec_znx_idft_apply(bo, 0, do, 0)
    big_equal(pb.NTT120, g.vec_znx_big_to_numpy(bg), bo)


@pytest.mark.parametrize("fl", FLAVOURS)
def test_vec_znx_dft_apply_step_offset(fl):
    """poulpy-hal/src/test_suite/vec_znx_dft.rs:360-460: a_size, res_size in 1..4 and (step, offset) grid."""
    n = 256
    g, o = mods(n, fl)
    rng = np.random.default_rng(11)
    k = base2k_for(fl)
    for a_size in range(1, 5):
        a = fill_uniform(rng, (a_size, 2, n), k)
        a_g = g.vec_znx_from_numpy(a)
        for res_size in range(1, 5):
            for step, offset in [(1, 0), (1, 1), (1, 2), (2, 2), (2, 0), (3, 1)]:
                # garbage pre-fill (test_suite/vmp.rs:215-216 style) -- same bytes on both sides are not required for
                # NTT120 (everything is overwritten); FFT64 leaves some limbs untouched, so start from zeros there.
                dg, do = g.vec_znx_dft_alloc(2, res_size), o.vec_znx_dft_alloc(2, res_size)
                for c in range(2):
                    g.vec_znx_dft_apply(step, offset, dg, c, a_g, c)
                    o.vec_znx_dft_apply(step, offset, do, c, a, c)
                dft_equal(fl, g.vec_znx_dft_to_numpy(dg), do, scale=n * 2.0 ** k)
                bg, bo = g.vec_znx_big_alloc(2, res_size), o.vec_znx_big_alloc(2, res_size)
                for c in range(2):
                    g.vec_znx_idft_apply(bg, c, dg, c)
                    o.vec_znx_idft_apply(bo, c, do, c)
                big_equal(fl, g.vec_znx_big_to_numpy(bg), bo)


@pytest.mark.parametrize("fl", FLAVOURS)
def test_idft_variants_and_sizes(fl):
    """idft_apply / _tmpa / _consume agree with the oracle for every (a_size, res_size) in 1..4 (vec_znx_dft.rs:462-645)."""
    n = 256
    g, o = mods(n, fl)
    rng = np.random.default_rng(5)
    k = base2k_for(fl)
    for a_size in range(1, 5):
        a = fill_uniform(rng, (a_size, 2, n), k)
        a_g = g.vec_znx_from_numpy(a)
        dg, do = g.vec_znx_dft_alloc(2, a_size), o.vec_znx_dft_alloc(2, a_size)
        for c in range(2):
            g.vec_znx_dft_apply(1, 0, dg, c, a_g, c)
            o.vec_znx_dft_apply(1, 0, do, c, a, c)
        for res_size in range(1, 5):
            bg, bo = g.vec_znx_big_alloc(2, res_size), o.vec_znx_big_alloc(2, res_size)
            bg2 = g.vec_znx_big_alloc(2, res_size)
            for c in range(2):
                g.vec_znx_idft_apply(bg, c, dg, c)
                g.vec_znx_idft_apply_tmpa(bg2, c, dg, c)
                o.vec_znx_idft_apply(bo, c, do, c)
            big_equal(fl, g.vec_znx_big_to_numpy(bg), bo)
            big_equal(fl, g.vec_znx_big_to_numpy(bg2), bo)
        cons_o = o.vec_znx_idft_apply_consume(do.copy())
        cons_g = g.vec_znx_idft_apply_consume(dg)
        big_equal(fl, g.vec_znx_big_to_numpy(cons_g), cons_o)


@pytest.mark.parametrize("fl", FLAVOURS)
def test_svp_apply_dft_to_dft(fl):
    """poulpy-hal/src/test_suite/svp.rs:108-: scalar x vector products, every size combination, normalised compare."""
    n = 256
    g, o = mods(n, fl)
    rng = np.random.default_rng(7)
    k = base2k_for(fl)
    s = fill_uniform(rng, (2, n), k)
    pg, po = g.svp_ppol_alloc(2), o.svp_ppol_alloc(2)
    s_g = g.scalar_znx_from_numpy(s)
    for c in range(2):
        g.svp_prepare(pg, c, s_g, c)
        o.svp_prepare(po, c, s, c)
    for b_size in range(1, 5):
        b = fill_uniform(rng, (b_size, 2, n), k)
        b_g = g.vec_znx_from_numpy(b)
        bdg, bdo = g.vec_znx_dft_alloc(2, b_size), o.vec_znx_dft_alloc(2, b_size)
        for c in range(2):
            g.vec_znx_dft_apply(1, 0, bdg, c, b_g, c)
            o.vec_znx_dft_apply(1, 0, bdo, c, b, c)
        for res_size in range(1, 5):
            rg, ro = g.vec_znx_dft_alloc(2, res_size), o.vec_znx_dft_alloc(2, res_size)
            for c in range(2):
                g.svp_apply_dft_to_dft(rg, c, pg, c, bdg, c)
                o.svp_apply_dft_to_dft(ro, c, po, c, bdo, c)
            dft_equal(fl, g.vec_znx_dft_to_numpy(rg), ro, scale=(n * 2.0 ** k) ** 2)
            # assign variant on a copy of b's DFT
            if res_size == b_size:
                ag, ao = g.vec_znx_dft_alloc(2, b_size), bdo.copy()
                for c in range(2):
                    g.vec_znx_dft_copy(1, 0, ag, c, bdg, c)
                    g.svp_apply_dft_to_dft_assign(ag, c, pg, c)
                    o.svp_apply_dft_to_dft_assign(ao, c, po, c)
                dft_equal(fl, g.vec_znx_dft_to_numpy(ag), ao, scale=(n * 2.0 ** k) ** 2)
            big_g, big_o = g.vec_znx_idft_apply_consume(rg), o.vec_znx_idft_apply_consume(ro)
            big_equal(fl, g.vec_znx_big_to_numpy(big_g), big_o)
            out_g, out_o = g.vec_znx_alloc(2, res_size), o.vec_znx_alloc(2, res_size)
            for c in range(2):
                g.vec_znx_big_normalize(out_g, k, 0, c, big_g, k, c)
                o.vec_znx_big_normalize(out_o, k, 0, c, big_o, k, c)
            assert np.array_equal(g.vec_znx_to_numpy(out_g), out_o)


@pytest.mark.parametrize("fl", FLAVOURS)
def test_vmp_apply_dft_to_dft_grid(fl):
    """poulpy-hal/src/test_suite/vmp.rs:131-313: cols_in/out in {1,2}, size_in/out in 1..4, every limb_offset."""
    n = 256
    g, o = mods(n, fl)
    rng = np.random.default_rng(9)
    k = base2k_for(fl)
    for cols_in in (1, 2):
        for cols_out in (1, 2):
            for size_in in range(1, 5):
                for size_out in range(1, 5):
                    rows = size_in
                    a = fill_uniform(rng, (size_in, cols_in, n), k)
                    a_g = g.vec_znx_from_numpy(a)
                    adg, ado = g.vec_znx_dft_alloc(cols_in, size_in), o.vec_znx_dft_alloc(cols_in, size_in)
                    for c in range(cols_in):
                        g.vec_znx_dft_apply(1, 0, adg, c, a_g, c)
                        o.vec_znx_dft_apply(1, 0, ado, c, a, c)
                    mat = fill_uniform(rng, (rows, cols_in, size_out, cols_out, n), k)
                    pmg, pmo = g.vmp_pmat_alloc(rows, cols_in, cols_out, size_out), o.vmp_pmat_alloc(rows, cols_in, cols_out, size_out)
                    g.vmp_prepare(pmg, g.mat_znx_from_numpy(mat))
                    o.vmp_prepare(pmo, mat)
                    for limb_offset in range(0, size_out):
                        rg, ro = g.vec_znx_dft_alloc(cols_out, size_out), o.vec_znx_dft_alloc(cols_out, size_out)
                        if limb_offset == 0 and fl == pb.NTT120:
                            # output buffers pre-filled with garbage (vmp.rs:215-216); FFT64 with limb_offset keeps stale polys
                            rg.buf.upload(rng.integers(0, 255, rg.buf.nbytes, dtype=np.uint8))
                        g.vmp_apply_dft_to_dft(rg, adg, pmg, limb_offset)
                        o.vmp_apply_dft_to_dft(ro, ado, pmo, limb_offset)
                        dft_equal(fl, g.vec_znx_dft_to_numpy(rg), ro, scale=rows * cols_in * (n * 2.0 ** k) ** 2)
                        big_g, big_o = g.vec_znx_idft_apply_consume(rg), o.vec_znx_idft_apply_consume(ro)
                        big_equal(fl, g.vec_znx_big_to_numpy(big_g), big_o)
                        out_g, out_o = g.vec_znx_alloc(cols_out, size_out), o.vec_znx_alloc(cols_out, size_out)
                        for c in range(cols_out):
                            g.vec_znx_big_normalize(out_g, k, 0, c, big_g, k, c)
                            o.vec_znx_big_normalize(out_o, k, 0, c, big_o, k, c)
                        assert np.array_equal(g.vec_znx_to_numpy(out_g), out_o)
                    # inputs are not mutated (digest checks of vmp.rs:198,210)
                    assert np.array_equal(g.vec_znx_to_numpy(a_g), a)


@pytest.mark.parametrize("fl", FLAVOURS)
def test_vmp_many_rows(fl):
    """More rows than one 16-row accumulation chunk (bench sweep shape [.., 31, 1, 2, ..] of poulpy-bench/src/params.rs:75-81)."""
    n, rows, cols_in, cols_out, size = 64, 31, 1, 2, 3
    g, o = mods(n, fl)
    rng = np.random.default_rng(19)
    k = base2k_for(fl)
    a = fill_uniform(rng, (rows, cols_in, n), k)
    adg, ado = g.vec_znx_dft_alloc(cols_in, rows), o.vec_znx_dft_alloc(cols_in, rows)
    g.vec_znx_dft_apply(1, 0, adg, 0, g.vec_znx_from_numpy(a), 0)
    o.vec_znx_dft_apply(1, 0, ado, 0, a, 0)
    mat = fill_uniform(rng, (rows, cols_in, size, cols_out, n), k)
    pmg, pmo = g.vmp_pmat_alloc(rows, cols_in, cols_out, size), o.vmp_pmat_alloc(rows, cols_in, cols_out, size)
    g.vmp_prepare(pmg, g.mat_znx_from_numpy(mat))
    o.vmp_prepare(pmo, mat)
    rg, ro = g.vec_znx_dft_alloc(cols_out, size), o.vec_znx_dft_alloc(cols_out, size)
    g.vmp_apply_dft_to_dft(rg, adg, pmg, 0)
    o.vmp_apply_dft_to_dft(ro, ado, pmo, 0)
    big_equal(fl, g.vec_znx_big_to_numpy(g.vec_znx_idft_apply_consume(rg)), o.vec_znx_idft_apply_consume(ro))


@pytest.mark.parametrize("fl", FLAVOURS)
def test_dft_add_sub_copy(fl):
    """DFT-domain add/sub/copy/zero size rules (ntt120/vec_znx_dft.rs:418-652, fft64/vec_znx_dft.rs)."""
    n = 64
    g, o = mods(n, fl)
    rng = np.random.default_rng(13)
    k = base2k_for(fl)

    def mk(size):
        x = fill_uniform(rng, (size, 1, n), k)
        dg, do = g.vec_znx_dft_alloc(1, size), o.vec_znx_dft_alloc(1, size)
        g.vec_znx_dft_apply(1, 0, dg, 0, g.vec_znx_from_numpy(x), 0)
        o.vec_znx_dft_apply(1, 0, do, 0, x, 0)
        return dg, do

    def check(dg, do):
        big_g = g.vec_znx_big_alloc(1, dg.size)
        big_o = o.vec_znx_big_alloc(1, dg.size)
        g.vec_znx_idft_apply(big_g, 0, dg, 0)
        o.vec_znx_idft_apply(big_o, 0, do, 0)
        big_equal(fl, g.vec_znx_big_to_numpy(big_g), big_o)

    for a_size in range(1, 4):
        for b_size in range(1, 4):
            ag, ao = mk(a_size)
            bg, bo = mk(b_size)
            for res_size in range(1, 5):
                for name in ("vec_znx_dft_add_into", "vec_znx_dft_sub"):
                    rg, ro = mk(res_size)
                    getattr(g, name)(rg, 0, ag, 0, bg, 0)
                    getattr(o, name)(ro, 0, ao, 0, bo, 0)
                    check(rg, ro)
            for name in ("vec_znx_dft_add_assign", "vec_znx_dft_sub_assign", "vec_znx_dft_sub_negate_assign"):
                rg, ro = mk(a_size)
                getattr(g, name)(rg, 0, bg, 0)
                getattr(o, name)(ro, 0, bo, 0)
                check(rg, ro)
            for step, offset in [(1, 0), (1, 1), (2, 1), (2, 0)]:
                rg, ro = mk(a_size)
                g.vec_znx_dft_copy(step, offset, rg, 0, bg, 0)
                o.vec_znx_dft_copy(step, offset, ro, 0, bo, 0)
                check(rg, ro)
    rg, ro = mk(3)
    g.vec_znx_dft_zero(rg, 0)
    o.vec_znx_dft_zero(ro, 0)
    check(rg, ro)


def _big_from_ints(g, fl, vals):
    """Upload a VecZnxBig from an int64 array (sign-extended), shape (size, cols, n)."""
    size, cols, n = vals.shape
    big = g.vec_znx_big_alloc(cols, size)
    if fl == pb.NTT120:
        big.buf.upload(O.int_to_i128(vals.astype(object)))
    else:
        big.buf.upload(vals)
    return big


@pytest.mark.parametrize("fl", FLAVOURS)
def test_vec_znx_big_normalize_grid(fl):
    """poulpy-hal/src/test_suite/vec_znx_big.rs:785-872: a_size, res_size in 1..4, res_offset in [-base2k, base2k],
    inputs = sign-extended 63-bit uniform; plus the fused +-assign variants (:874-1028, NTT120 only)."""
    n = 64
    g, o = mods(n, fl)
    rng = np.random.default_rng(17)
    k = base2k_for(fl)
    for a_size in range(1, 5):
        vals = fill_uniform(rng, (a_size, 2, n), 63)
        big_g = _big_from_ints(g, fl, vals)
        big_o = O.int_to_i128(vals.astype(object)) if fl == pb.NTT120 else vals.copy()
        for res_size in range(1, 5):
            for res_offset in list(range(-k, k + 1, 5)) + [-1, 1, k - 1, -(k - 1)]:
                init = fill_uniform(rng, (res_size, 2, n), k)
                out_g, out_o = g.vec_znx_from_numpy(init), init.copy()
                for c in range(2):
                    g.vec_znx_big_normalize(out_g, k, res_offset, c, big_g, k, c)
                    o.vec_znx_big_normalize(out_o, k, res_offset, c, big_o, k, c)
                assert np.array_equal(g.vec_znx_to_numpy(out_g), out_o), (a_size, res_size, res_offset)
                if fl == pb.NTT120:
                    for op in (1, -1):
                        out_g, out_o = g.vec_znx_from_numpy(init), init.copy()
                        for c in range(2):
                            g.vec_znx_big_normalize(out_g, k, res_offset, c, big_g, k, c, op=op)
                            o.vec_znx_big_normalize(out_o, k, res_offset, c, big_o, k, c, op=op)
                        assert np.array_equal(g.vec_znx_to_numpy(out_g), out_o), (a_size, res_size, res_offset, op)


@pytest.mark.parametrize("fl", FLAVOURS)
@pytest.mark.parametrize("a_k,res_k", [(12, 17), (17, 12), (18, 19), (19, 18), (52, 50), (50, 52), (7, 51)])
def test_vec_znx_big_normalize_cross_base2k(fl, a_k, res_k):
    n = 32
    g, o = mods(n, fl)
    rng = np.random.default_rng(a_k * 100 + res_k)
    for a_size in (1, 2, 4):
        vals = fill_uniform(rng, (a_size, 1, n), 60)
        big_g = _big_from_ints(g, fl, vals)
        big_o = O.int_to_i128(vals.astype(object)) if fl == pb.NTT120 else vals.copy()
        for res_size in (1, 3, 5):
            for res_offset in (-(a_k + 1), -a_k, -3, 0, 2, a_k - 1, a_k, a_k + 1, 2 * a_k + 3):
                init = fill_uniform(rng, (res_size, 1, n), res_k)
                ops = (0, 1, -1) if fl == pb.NTT120 else (0,)
                for op in ops:
                    out_g, out_o = g.vec_znx_from_numpy(init), init.copy()
                    g.vec_znx_big_normalize(out_g, res_k, res_offset, 0, big_g, a_k, 0, op=op)
                    o.vec_znx_big_normalize(out_o, res_k, res_offset, 0, big_o, a_k, 0, op=op)
                    assert np.array_equal(g.vec_znx_to_numpy(out_g), out_o), (a_size, res_size, res_offset, op)


@pytest.mark.parametrize("fl", FLAVOURS)
def test_big_add_small_and_rotate(fl):
    n = 128
    g, o = mods(n, fl)
    rng = np.random.default_rng(23)
    vals = fill_uniform(rng, (3, 2, n), 63)
    small = fill_uniform(rng, (2, 2, n), 64)
    big_g = _big_from_ints(g, fl, vals)
    big_o = O.int_to_i128(vals.astype(object)) if fl == pb.NTT120 else vals.copy()
    for c in range(2):
        g.vec_znx_big_add_small_assign(big_g, c, g.vec_znx_from_numpy(small), c)
        o.vec_znx_big_add_small_assign(big_o, c, small, c)
    big_equal(fl, g.vec_znx_big_to_numpy(big_g), big_o)
    for p in (0, 1, -1, 5, n - 1, n, n + 3, 2 * n - 1, -n - 7, 12345):
        rg, ro = g.vec_znx_alloc(2, 3), np.zeros((3, 2, n), dtype=np.int64)
        for c in range(2):
            g.vec_znx_rotate(p, rg, c, g.vec_znx_from_numpy(small), c)
            O.vec_znx_rotate(p, ro, c, small, c)
        assert np.array_equal(g.vec_znx_to_numpy(rg), ro), p


def test_managed_buffers_are_host_dereferenceable():
    """The reference requires DataRef: AsRef<[u8]> (poulpy-hal/src/layouts/mod.rs:56): pgb_alloc_bytes memory can be read and
    written from the host between (synchronous) HAL calls, e.g. `res_dft.zero()` on the host (keyswitching/glwe.rs:90)."""
    n = 256
    g = pb.Module(n, pb.NTT120, managed=True)
    o = O.OracleModule(n, O.NTT120)
    rng = np.random.default_rng(29)
    a = fill_uniform(rng, (2, 1, n), 18)
    a_g = g.vec_znx_alloc(1, 2)
    a_g.buf.host_view(np.int64, (2, 1, n))[:] = a  # host write
    dg, do = g.vec_znx_dft_alloc(1, 2), o.vec_znx_dft_alloc(1, 2)
    g.vec_znx_dft_apply(1, 0, dg, 0, a_g, 0)
    o.vec_znx_dft_apply(1, 0, do, 0, a, 0)
    host = dg.buf.host_view(np.uint32, (2, 1, 4, n))  # host read of the backend-owned layout
    for k, q in enumerate(O.Q):
        assert np.array_equal(host[:, :, k, :].astype(np.uint64), do[:, :, :, k] % np.uint64(q))
    host[1] = 0  # host-side zero of one limb, as the reference's core code does
    big_g = g.vec_znx_idft_apply_consume(dg)
    vals = O.i128_to_int(big_g.buf.host_view(np.uint64, (2, 1, n, 2)))
    assert np.array_equal(vals[0].astype(np.int64), a[0]) and not np.any(vals[1])


def test_error_model():
    """Shape violations return an error (reference: assert!/panic!) instead of corrupting memory."""
    g, _ = mods(64, pb.NTT120)
    a = g.vec_znx_alloc(1, 2)
    d = g.vec_znx_dft_alloc(1, 2)
    with pytest.raises(pb.PoulpyError):
        g.vec_znx_dft_apply(1, 0, d, 3, a, 0)  # column out of range
    with pytest.raises(pb.PoulpyError):
        g.vec_znx_dft_apply(0, 0, d, 0, a, 0)  # step == 0
    other = pb.Module(128, pb.NTT120)
    with pytest.raises(pb.PoulpyError):
        other.vec_znx_dft_apply(1, 0, d, 0, a, 0)  # ring degree mismatch
    with pytest.raises(pb.PoulpyError):
        pb.Module(48, pb.NTT120)  # not a power of two


@pytest.mark.parametrize("fl", FLAVOURS)
def test_svp_apply_dft_and_vmp_apply_dft_from_coefficients(fl):
    """HalImpl::svp_apply_dft (hal_impl.rs:600) and vmp_apply_dft (:636): the coefficient-domain front ends.  The oracle side spells
    out the reference bodies: fft64/svp.rs:21-55 / hal_defaults/svp_ppol.rs:93-107 (dft_apply, then svp_apply_dft_to_dft) and
    hal_impl/family_common.rs:17-58 (last min(a.cols, cols_in) columns transformed into a zero-padded VecZnxDft, then
    vmp_apply_dft_to_dft); outputs compared after idft + normalize, bit for bit."""
    n = 128
    g, o = mods(n, fl)
    rng = np.random.default_rng(19)
    k = base2k_for(fl)
    s = fill_uniform(rng, (1, n), k)
    pg, po = g.svp_ppol_alloc(1), o.svp_ppol_alloc(1)
    g.svp_prepare(pg, 0, g.scalar_znx_from_numpy(s), 0)
    o.svp_prepare(po, 0, s, 0)
    for b_size, res_size in ((1, 1), (3, 2), (2, 4)):
        b = fill_uniform(rng, (b_size, 2, n), k)
        rg, ro = g.vec_znx_dft_alloc(1, res_size), o.vec_znx_dft_alloc(1, res_size)
        rg.buf.upload(rng.integers(0, 255, rg.buf.nbytes, dtype=np.uint8))
        g.svp_apply_dft(rg, 0, pg, 0, g.vec_znx_from_numpy(b), 1)
        bdo = o.vec_znx_dft_alloc(1, b_size)
        o.vec_znx_dft_apply(1, 0, bdo, 0, b, 1)
        o.svp_apply_dft_to_dft(ro, 0, po, 0, bdo, 0)
        big_g, big_o = g.vec_znx_idft_apply_consume(rg), o.vec_znx_idft_apply_consume(ro)
        out_g, out_o = g.vec_znx_alloc(1, res_size), o.vec_znx_alloc(1, res_size)
        g.vec_znx_big_normalize(out_g, k, 0, 0, big_g, k, 0)
        o.vec_znx_big_normalize(out_o, k, 0, 0, big_o, k, 0)
        assert np.array_equal(g.vec_znx_to_numpy(out_g), out_o), (b_size, res_size)
    for a_cols, cols_in, cols_out, a_size, rows, size_out in ((1, 1, 2, 3, 3, 4), (2, 1, 1, 2, 4, 2), (1, 2, 2, 4, 2, 3), (3, 2, 1, 2, 2, 2)):
        a = fill_uniform(rng, (a_size, a_cols, n), k)
        mat = fill_uniform(rng, (rows, cols_in, size_out, cols_out, n), k)
        pmg, pmo = g.vmp_pmat_alloc(rows, cols_in, cols_out, size_out), o.vmp_pmat_alloc(rows, cols_in, cols_out, size_out)
        g.vmp_prepare(pmg, g.mat_znx_from_numpy(mat))
        o.vmp_prepare(pmo, mat)
        rg, ro = g.vec_znx_dft_alloc(cols_out, size_out), o.vec_znx_dft_alloc(cols_out, size_out)
        g.vmp_apply_dft(rg, g.vec_znx_from_numpy(a), pmg)
        copy = min(a_cols, cols_in)
        a_dft_size = min(a_size, rows)
        ado = o.vec_znx_dft_alloc(cols_in, a_dft_size)  # allocated zeroed: the leading `offset` columns stay zero
        for j in range(copy):
            o.vec_znx_dft_apply(1, 0, ado, cols_in - copy + j, a, a_cols - copy + j)
        o.vmp_apply_dft_to_dft(ro, ado, pmo, 0)
        big_g, big_o = g.vec_znx_idft_apply_consume(rg), o.vec_znx_idft_apply_consume(ro)
        out_g, out_o = g.vec_znx_alloc(cols_out, size_out), o.vec_znx_alloc(cols_out, size_out)
        for c in range(cols_out):
            g.vec_znx_big_normalize(out_g, k, 0, c, big_g, k, c)
            o.vec_znx_big_normalize(out_o, k, 0, c, big_o, k, c)
        assert np.array_equal(g.vec_znx_to_numpy(out_g), out_o), (a_cols, cols_in, cols_out, a_size, rows, size_out)


def test_coefficient_domain_helpers():
    """vec_znx_add_assign / sub_assign / mul_xp_minus_one / normalize_assign (the helpers of execute_standard) against
    the oracle / numpy, including size mismatches and rotations across the negacyclic wrap."""
    n = 64
    g = pb.Module(n, pb.FFT64)
    rng = np.random.default_rng(23)
    for res_size, a_size in ((3, 3), (2, 4), (4, 2)):
        r0 = rng.integers(-(1 << 40), 1 << 40, size=(res_size, 2, n), dtype=np.int64)
        a = rng.integers(-(1 << 40), 1 << 40, size=(a_size, 2, n), dtype=np.int64)
        mn = min(res_size, a_size)
        rg, ag = g.vec_znx_from_numpy(r0), g.vec_znx_from_numpy(a)
        g.vec_znx_add_assign(rg, 1, ag, 0)
        want = r0.copy()
        want[:mn, 1] += a[:mn, 0]
        assert np.array_equal(g.vec_znx_to_numpy(rg), want)
        g.vec_znx_sub_assign(rg, 1, ag, 0)
        assert np.array_equal(g.vec_znx_to_numpy(rg), r0)
        for p in (0, 1, n - 1, n, n + 7, -5, 3 * n + 1):
            rg = g.vec_znx_from_numpy(r0)
            g.vec_znx_mul_xp_minus_one(p, rg, 0, ag, 1)
            rot = np.zeros((res_size, 1, n), dtype=np.int64)
            O.vec_znx_rotate(p, rot, 0, np.ascontiguousarray(a[:, 1:2]), 0)
            want = r0.copy()
            want[:, 0] = rot[:, 0]
            want[:mn, 0] -= a[:mn, 1]
            assert np.array_equal(g.vec_znx_to_numpy(rg), want), (res_size, a_size, p)
    for size in (1, 2, 4):
        x = rng.integers(-(1 << 50), 1 << 50, size=(size, 2, n), dtype=np.int64)
        xg = g.vec_znx_from_numpy(x)
        want = x.copy()
        for c in range(2):
            g.vec_znx_normalize_assign(13, xg, c)
            O.vec_znx_normalize_assign(13, want, c)
        assert np.array_equal(g.vec_znx_to_numpy(xg), want), size


@pytest.mark.parametrize("batch", [6, 37])
def test_vmp_batch_tiled_large_key(batch):
    """Matrices beyond 48 MB take the batch-tiled vmp kernel (one matrix read per four ciphertexts).  batch = 6 leaves a partial tile of two,
    batch = 37 one of one; the result must equal the per-item kernel bit for bit, and item 0 must match the oracle after idft + normalize."""
    n, rows, cols_out, size, k = 16384, 13, 2, 8, 30
    g, o = pb.Module(n, pb.NTT120), O.OracleModule(n, pb.NTT120)
    rng = np.random.default_rng(29)
    mat = fill_uniform(rng, (rows, 1, size, cols_out, n), k)
    pmg, pmo = g.vmp_pmat_alloc(rows, 1, cols_out, size), o.vmp_pmat_alloc(rows, 1, cols_out, size)
    g.vmp_prepare(pmg, g.mat_znx_from_numpy(mat))
    a = fill_uniform(rng, (batch, rows, 1, n), k)
    a_g = g.vec_znx_from_numpy(a)
    adg = g.vec_znx_dft_alloc(1, rows, batch)
    g.vec_znx_dft_apply(1, 0, adg, 0, a_g, 0)
    outs = []
    for no_bt in (False, True):
        g.set_option(pb.hal.OPT_VMP_NO_BT, int(no_bt))
        try:
            rg = g.vec_znx_dft_alloc(cols_out, size, batch)
            rg.buf.upload(rng.integers(0, 255, rg.buf.nbytes, dtype=np.uint8))
            g.vmp_apply_dft_to_dft(rg, adg, pmg, 0)
            g.sync()
            outs.append(rg.buf.download(np.uint32, (rg.buf.nbytes // 4,)).copy())
        finally:
            g.set_option(pb.hal.OPT_VMP_NO_BT, 0)
    assert np.array_equal(outs[0], outs[1])
    o.vmp_prepare(pmo, mat)
    ado, ro = o.vec_znx_dft_alloc(1, rows), o.vec_znx_dft_alloc(cols_out, size)
    o.vec_znx_dft_apply(1, 0, ado, 0, a[0], 0)
    o.vmp_apply_dft_to_dft(ro, ado, pmo, 0)
    big_o = o.vec_znx_idft_apply_consume(ro)
    r0 = g.vec_znx_dft_alloc(cols_out, size)
    g.vmp_apply_dft_to_dft(r0, _first_item(g, adg), pmg, 0)
    big_g = g.vec_znx_idft_apply_consume(r0)
    out_g, out_o = g.vec_znx_alloc(cols_out, size), o.vec_znx_alloc(cols_out, size)
    for c in range(cols_out):
        g.vec_znx_big_normalize(out_g, k, 0, c, big_g, k, c)
        o.vec_znx_big_normalize(out_o, k, 0, c, big_o, k, c)
    assert np.array_equal(g.vec_znx_to_numpy(out_g), out_o)


def _first_item(g, v):
    """View of batch item 0 of a batched device container."""
    return type(v)(v.buf, v.n, v.cols, v.size, v.offset, 1, v.batch_stride)


def test_device_bytes_view_for_collectives():
    """sharding._DeviceBytes: torch sees a raw device allocation through __cuda_array_interface__ without a copy (what broadcast_prepared hands
    to NCCL); without a process group broadcast_prepared is a no-op."""
    import torch

    from poulpy_b200.sharding import _DeviceBytes, broadcast_prepared
    g = pb.Module(256, pb.NTT120)
    rng = np.random.default_rng(77)
    pm = g.vmp_pmat_alloc(2, 1, 2, 2)
    g.vmp_prepare(pm, g.mat_znx_from_numpy(rng.integers(-100, 100, size=(2, 1, 2, 2, 256), dtype=np.int64)))
    g.sync()
    want = pm.buf.download(np.uint8, (pm.buf.nbytes,))
    t = torch.as_tensor(_DeviceBytes(pm.buf.ptr, pm.buf.nbytes), device="cuda")
    assert t.data_ptr() == pm.buf.ptr and t.numel() == pm.buf.nbytes
    assert np.array_equal(t.cpu().numpy(), want)
    t.zero_()  # writes through to the allocation
    torch.cuda.synchronize()
    assert not pm.buf.download(np.uint8, (pm.buf.nbytes,)).any()
    broadcast_prepared(g, pm.buf)


def test_ntt120_vmp_odd_col_max_matches_the_reference_quirk():
    """reference/ntt120/vmp.rs:262-273: when col_max = min(ncols, res polys + offset) is ODD and smaller than the matrix, the reference
    computes the last output poly from the paired-column block read with the single-column stride (row i of the product takes matrix row
    i >> 1, column last + (i & 1)).  Round 1 computed the mathematically defined product there and documented the deviation; "identical to
    the reference on the same inputs" now includes this shape: the GPU reproduces what the oracle (a restatement of that code) returns."""
    n, k = 256, 18
    g, o = mods(n, pb.NTT120)
    rng = np.random.default_rng(2100)
    for rows, cols_in, cols_out, size, res_cols, res_size, off in ((3, 1, 2, 4, 1, 3, 0), (4, 2, 2, 3, 1, 5, 0), (3, 1, 2, 4, 1, 1, 1), (5, 1, 3, 3, 1, 3, 0),
                                                                   (3, 1, 2, 4, 1, 2, 1)):
        a = fill_uniform(rng, (rows, cols_in, n), k)
        mat = fill_uniform(rng, (rows, cols_in, size, cols_out, n), k)
        adg, ado = g.vec_znx_dft_alloc(cols_in, rows), o.vec_znx_dft_alloc(cols_in, rows)
        a_g = g.vec_znx_from_numpy(a)
        for c in range(cols_in):
            g.vec_znx_dft_apply(1, 0, adg, c, a_g, c)
            o.vec_znx_dft_apply(1, 0, ado, c, a, c)
        pmg, pmo = g.vmp_pmat_alloc(rows, cols_in, cols_out, size), o.vmp_pmat_alloc(rows, cols_in, cols_out, size)
        g.vmp_prepare(pmg, g.mat_znx_from_numpy(mat))
        o.vmp_prepare(pmo, mat)
        rg, ro = g.vec_znx_dft_alloc(res_cols, res_size), o.vec_znx_dft_alloc(res_cols, res_size)
        rg.buf.upload(rng.integers(0, 255, rg.buf.nbytes, dtype=np.uint8))
        g.vmp_apply_dft_to_dft(rg, adg, pmg, off)
        o.vmp_apply_dft_to_dft(ro, ado, pmo, off)
        col_max = min(cols_out * size, res_cols * res_size + off * cols_out)
        dft_equal(pb.NTT120, g.vec_znx_dft_to_numpy(rg), ro, scale=1.0)
        big_g, big_o = g.vec_znx_idft_apply_consume(rg), o.vec_znx_idft_apply_consume(ro)
        big_equal(pb.NTT120, g.vec_znx_big_to_numpy(big_g), big_o)
        assert (col_max % 2 == 1 and col_max < cols_out * size) or (rows, off) == (3, 1), (rows, cols_out, size, res_size, off, col_max)
