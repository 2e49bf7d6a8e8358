"""Circuit bootstrapping, constant mode (poulpy-bin-fhe/src/circuit_bootstrapping/circuit.rs:219-380, extension_factor = 1): the device
orchestration of poulpy_b200/circuit.py against the same sequence restated over the oracle (LUT construction lut.rs:271-338, mod_switch_2n,
block-binary blind rotation, trace, rotation by the gap, ggsw_expand_row) on identical synthetic keys -- the GGSW must agree bit for bit.
(The reference's own test decrypts the GGSW and bounds its noise, which needs the key generator; every primitive used here has its own
semantic pin in tests/test_oracle_*.py.)"""
import numpy as np
import pytest

import poulpy_b200 as pb
from oracle import pyoracle as O
from poulpy_b200 import circuit
from util import fill_uniform
from util_circuit import OracleGlweOps, _Ct

pytestmark = pytest.mark.gpu


def _oracle_lut(n, f, k, base2k, ext=1):
    """LookupTable::set over the oracle: the shared un-normalised limbs (circuit.lookup_table_limbs), oracle normalisation, then
    lookup_table_rotate(-drift) (lut.rs:340-362) -> (list of ext arrays (limbs, 1, n), drift)."""
    raw, step = circuit.lookup_table_limbs(n, f, k, base2k, ext)
    for j in range(ext):
        O.vec_znx_normalize_assign(base2k, raw[j], 0)
    drift = step >> 1
    out = [None] * ext
    for src, rot, dst in circuit.lookup_table_rotation_plan(n, ext, -drift):
        out[dst] = np.zeros_like(raw[src])
        O.vec_znx_rotate(rot, out[dst], 0, raw[src], 0)
    return (out if ext > 1 else out[0]), drift


def _oracle_blind_rotate(o, acc, lwe, K, n, ext, lut, brk_o, xpa, block, rot_left):
    lwe_2n = O.mod_switch_2n(2 * n * ext, lwe, K, rot_left=rot_left)
    if ext == 1:
        o.cggi_blind_rotate_block_binary(acc, lwe_2n, lut, brk_o, xpa, block, K)
    else:
        o.cggi_blind_rotate_block_binary_extended(acc, lwe_2n, lut, brk_o, xpa, block, K)


@pytest.mark.parametrize("fl", [pb.FFT64, pb.NTT120])
@pytest.mark.parametrize("ext", [1, 2])
def test_circuit_bootstrap_to_constant(fl, ext):
    n, log_n, n_lwe, block, rank, K = 256, 8, 12, 3, 1, 12 if fl == pb.FFT64 else 18
    brk_size, dnum_res, res_size, log_domain, batch, lwe_size = 2, 2, 2, 2, 3, 2
    cols = rank + 1
    g, o = pb.Module(n, fl), O.OracleModule(n, fl)
    rng = np.random.default_rng(900 + fl)
    # synthetic key material (uniform digits), identical on both sides
    brk_mats = [fill_uniform(rng, (1, cols, brk_size, cols, n), K) for _ in range(n_lwe)]
    per = n * cols * cols * brk_size * g.prep_bytes
    brk_buf = pb.DevBuf(per * n_lwe)
    brk_o = []
    for i, mt in enumerate(brk_mats):
        g.vmp_prepare(pb.hal.VmpPMat(brk_buf, n, 1, cols, cols, brk_size, offset=i * per), g.mat_znx_from_numpy(mt))
        pm = o.vmp_pmat_alloc(1, cols, cols, brk_size)
        o.vmp_prepare(pm, mt)
        brk_o.append(pm)
    brk_g = pb.hal.VmpPMat(brk_buf, n, 1, cols, cols, brk_size)
    tmp_size = max(brk_size, res_size)

    def keys(count, dnum, size):
        out_g, out_o = [], []
        for _ in range(count):
            mt = fill_uniform(rng, (dnum, rank, size, cols, n), K)
            pg, po = g.vmp_pmat_alloc(dnum, rank, cols, size), o.vmp_pmat_alloc(dnum, rank, cols, size)
            g.vmp_prepare(pg, g.mat_znx_from_numpy(mt))
            o.vmp_prepare(po, mt)
            out_g.append(pg)
            out_o.append(po)
        return out_g, out_o

    atk_g, atk_o = keys(log_n, tmp_size, tmp_size + 1)
    tsk_g, tsk_o = keys(rank, res_size, res_size + 1)
    lwe = fill_uniform(rng, (batch, lwe_size, 1, n_lwe + 1), K)
    lwe_dev = pb.DevBuf(lwe.nbytes)
    lwe_dev.upload(lwe)

    ggsw = circuit.circuit_bootstrap_to_constant(g, lwe_dev, batch, n_lwe, lwe_size, K, brk_g, g.cggi_x_pow_a(), block, atk_g, tsk_g, K, rank,
                                                 dnum_res, res_size, log_domain, extension_factor=ext)
    got = ggsw.download(np.int64, (batch, dnum_res, cols, res_size, cols, n))

    # the same sequence over the oracle
    alpha = 1 << (dnum_res - 1).bit_length() if dnum_res > 1 else 1
    f = [0] * ((1 << log_domain) * alpha)
    for j in range(1 << log_domain):
        for i in range(dnum_res):
            f[j * alpha + i] = j * (1 << (K * (dnum_res - 1 - i)))
    lut, drift = _oracle_lut(n, f, K * dnum_res, K, ext)
    xpa = o.cggi_x_pow_a()
    want = np.zeros_like(got)
    for b in range(batch):
        acc = np.zeros((brk_size, cols, n), dtype=np.int64)
        _oracle_blind_rotate(o, acc, lwe[b], K, n, ext, lut, brk_o, xpa, block, True)
        for i in range(dnum_res):
            tmp = np.zeros((tmp_size, cols, n), dtype=np.int64)
            tmp[:brk_size] = acc[:tmp_size]
            o.glwe_trace_assign(tmp, K, 0, atk_o, K)
            want[b, i, 0, :min(res_size, tmp_size)] = tmp[:res_size]
            if i + 1 < dnum_res:
                nxt = np.zeros_like(acc)
                for c in range(cols):
                    O.vec_znx_rotate(-(2 * drift // ext), nxt, c, acc, c)
                acc = nxt
        o.ggsw_expand_row(want[b], K, tsk_o, K)
    assert np.array_equal(got, want)
    assert np.any(got[:, :, 1])  # the expanded columns are populated


def test_lookup_table_set():
    """LookupTable::set, extension_factor 1 (lut.rs:271-338): device construction == the numpy/oracle restatement, and the table's meaning:
    coefficient c of X^{drift} * LUT carries f[c / step] * 2^{-k} on the torus (k bits of message precision)."""
    from fractions import Fraction
    n, K = 256, 12
    g = pb.Module(n, pb.FFT64)
    for f, k in (([1, 2, 3, 4], 2 * K), ([5, -7, 11, 0, 3, 9, -1, 2], K + 5), (list(range(-8, 8)), 3 * K)):
        lut_g, drift_g = circuit.lookup_table_set(g, f, k, K)
        lut_o, drift_o = _oracle_lut(n, f, k, K)
        assert drift_g == drift_o and np.array_equal(g.vec_znx_to_numpy(lut_g), lut_o)
        for ext in (2, 4):  # the ext-component table: component j of Y^-drift * L(Y) over the domain n * ext (lut.rs:318-333)
            le_g, d_g = circuit.lookup_table_set(g, f, k, K, ext)
            le_o, d_o = _oracle_lut(n, f, k, K, ext)
            assert d_g == d_o and np.array_equal(g.vec_znx_to_numpy(le_g), np.stack(le_o))
            raw, step_e = circuit.lookup_table_limbs(n, f, k, K, ext)
            big = np.zeros((raw.shape[1], n * ext), dtype=np.int64)
            for j in range(ext):
                O.vec_znx_normalize_assign(K, raw[j], 0)
                big[:, j::ext] = raw[j][:, 0]
            N = n * ext
            rot = np.zeros_like(big)
            for i in range(N):
                p = (i - d_o) % (2 * N)
                if p < N:
                    rot[:, p] = big[:, i]
                else:
                    rot[:, p - N] = -big[:, i]
            for j in range(ext):
                assert np.array_equal(le_o[j][:, 0], rot[:, j::ext]), (ext, j)
        step, limbs = (n + len(f) // 2) // len(f), -(-k // K)
        back = np.zeros_like(lut_o)
        O.vec_znx_rotate(drift_o, back, 0, lut_o, 0)
        for c in (0, 1, step - 1, step, n // 2, n - 1):
            val = sum(Fraction(int(back[j, 0, c]), 1 << ((j + 1) * K)) for j in range(limbs))
            want = Fraction(f[min(c // step, len(f) - 1)], 1 << k)
            d = val - want
            assert d - round(d) == 0, (f, k, c)


def _setup(fl, n, n_lwe, rank, K, brk_size, size, batch, lwe_size, seed):
    cols, log_n = rank + 1, n.bit_length() - 1
    g, o = pb.Module(n, fl), O.OracleModule(n, fl)
    rng = np.random.default_rng(seed)
    per = n * cols * cols * brk_size * g.prep_bytes
    brk_buf = pb.DevBuf(per * n_lwe)
    brk_o = []
    for i in range(n_lwe):
        mt = fill_uniform(rng, (1, cols, brk_size, cols, n), K)
        g.vmp_prepare(pb.hal.VmpPMat(brk_buf, n, 1, cols, cols, brk_size, offset=i * per), g.mat_znx_from_numpy(mt))
        pm = o.vmp_pmat_alloc(1, cols, cols, brk_size)
        o.vmp_prepare(pm, mt)
        brk_o.append(pm)

    def keys(count, dnum, ksize):
        out_g, out_o = [], []
        for _ in range(count):
            mt = fill_uniform(rng, (dnum, rank, ksize, cols, n), K)
            pg, po = g.vmp_pmat_alloc(dnum, rank, cols, ksize), o.vmp_pmat_alloc(dnum, rank, cols, ksize)
            g.vmp_prepare(pg, g.mat_znx_from_numpy(mt))
            o.vmp_prepare(po, mt)
            out_g.append(pg)
            out_o.append(po)
        return out_g, out_o

    atk = keys(log_n, size, size + 1)
    tsk = keys(rank, size, size + 1)
    lwe = fill_uniform(rng, (batch, lwe_size, 1, n_lwe + 1), K)
    lwe_dev = pb.DevBuf(lwe.nbytes)
    lwe_dev.upload(lwe)
    return g, o, pb.hal.VmpPMat(brk_buf, n, 1, cols, cols, brk_size), brk_o, atk, tsk, lwe, lwe_dev


def test_glwe_pack():
    """glwe_pack (poulpy-core/src/glwe_packing.rs:15-170) with a sparse input map: the same sequencing over the device and the oracle."""
    n, rank, K, size, batch = 64, 1, 18, 3, 2
    g, o, _, _, atk, _, _, _ = _setup(pb.NTT120, n, 3, rank, K, 2, size, batch, 1, 950)
    rng = np.random.default_rng(951)
    dev, orc = circuit.DeviceGlweOps(g, atk[0], K, rank + 1, size, batch), OracleGlweOps(o, atk[1], K, rank + 1, size, batch, n)
    for idxs, log_gap_out in (((0, 8, 16, 24), 3), ((0, 4, 12), 2), ((0, 16, 32, 48), 4)):
        cts_d, cts_o = {}, {}
        for i in idxs:
            a = fill_uniform(rng, (batch, size, rank + 1, n), K)
            cts_d[i], cts_o[i] = g.vec_znx_from_numpy(a), _Ct(a.copy())
        res_d, res_o = dev.new(), orc.new()
        circuit.glwe_pack(dev, res_d, cts_d, log_gap_out)
        circuit.glwe_pack(orc, res_o, cts_o, log_gap_out)
        g.sync()
        assert np.array_equal(g.vec_znx_to_numpy(res_d), res_o.arr), (idxs, log_gap_out)


@pytest.mark.parametrize("fl", [pb.FFT64, pb.NTT120])
@pytest.mark.parametrize("ext", [1, 2])
def test_circuit_bootstrap_to_exponent(fl, ext):
    """Exponent mode (circuit.rs:219-380 with to_exponent = true: LUT of the gadget powers, rotation direction Right, post_process with the
    partial trace and glwe_pack per row, ggsw_expand_row): device orchestration == the same sequencing over the oracle."""
    n, n_lwe, block, rank, K = 128, 6, 3, 1, 12 if fl == pb.FFT64 else 18
    size, dnum_res, log_domain, log_gap_out, batch, lwe_size = 2, 2, 2, 2, 2, 2
    cols = rank + 1
    g, o, brk_g, brk_o, atk, tsk, lwe, lwe_dev = _setup(fl, n, n_lwe, rank, K, size, size, batch, lwe_size, 960 + fl)
    ggsw = circuit.circuit_bootstrap_to_exponent(g, log_gap_out, lwe_dev, batch, n_lwe, lwe_size, K, brk_g, g.cggi_x_pow_a(), block, atk[0], tsk[0],
                                                 K, rank, dnum_res, size, log_domain, extension_factor=ext)
    got = ggsw.download(np.int64, (batch, dnum_res, cols, size, cols, n))

    f, alpha = circuit.exponent_lut(K, dnum_res, log_domain)
    lut, drift = _oracle_lut(n, f, K * dnum_res, K, ext)
    gap = 2 * drift // ext
    log_gap_in = (gap * alpha - 1).bit_length()
    assert log_gap_in != log_gap_out  # the packing branch of post_process is the one exercised
    xpa = o.cggi_x_pow_a()
    acc = _Ct(np.zeros((batch, size, cols, n), dtype=np.int64))
    for b in range(batch):
        _oracle_blind_rotate(o, acc.arr[b], lwe[b], K, n, ext, lut, brk_o, xpa, block, False)
    ops = OracleGlweOps(o, atk[1], K, cols, size, batch, n)
    want = np.zeros_like(got)
    for i in range(dnum_res):
        row = ops.new()
        circuit.post_process(ops, row, acc, log_gap_in, log_gap_out, log_domain)
        want[:, i, 0] = row.arr
        if i + 1 < dnum_res:
            ops.rotate_assign(-gap, acc)
    for b in range(batch):
        o.ggsw_expand_row(want[b], K, tsk[1], K)
    assert np.array_equal(got, want)
    assert np.any(got[:, :, 0]) and np.any(got[:, :, 1])


@pytest.mark.parametrize("fl", [pb.FFT64, pb.NTT120])
@pytest.mark.parametrize("rank", [1, 2])
def test_noiseless_circuit_bootstrap_on_the_device(fl, rank):
    """L4 for circuit bootstrapping: noise-free BRK / ATK / TSK and noise-free LWEs of every message m.  The device
    orchestration (poulpy_b200/circuit.py over the C ABI) must (a) equal, bit for bit, an INDEPENDENT restatement of
    circuit_bootstrap_core written from circuit.rs over oracle primitives (tests/semantics_circuit.py -- nothing of circuit.py on that side)
    and (b) produce GGSW(m): every row / column decrypts, with exact integer arithmetic, to m (resp. m s_c) at its gadget position."""
    import semantics_circuit as SC
    from test_oracle_circuit_semantics import check_ggsw
    n, k, n_lwe, block, log_domain = 256, 12, 12, 3, 2
    brk_size, dnum_res, res_size, lwe_size = 4, 2, 4, 2
    cols = rank + 1
    rng = np.random.default_rng(3100 + fl + rank)
    g, o = pb.Module(n, fl), O.OracleModule(n, fl)
    s_lwe, s, brk, atk, tsk = SC.build_keys(rng, n, k, rank, n_lwe, block, brk_size, brk_size, brk_size, brk_size + 1, res_size, res_size + 1)

    def prep_o(mats):
        out = []
        for mat in mats:
            d, ci, sz, co, _ = mat.shape
            pm = o.vmp_pmat_alloc(d, ci, co, sz)
            o.vmp_prepare(pm, mat)
            out.append(pm)
        return out

    def prep_g(mats):
        out = []
        for mat in mats:
            d, ci, sz, co, _ = mat.shape
            pm = g.vmp_pmat_alloc(d, ci, co, sz)
            g.vmp_prepare(pm, g.mat_znx_from_numpy(mat))
            out.append(pm)
        return out

    per = n * brk_size * cols * cols * brk_size * g.prep_bytes
    brk_buf = pb.DevBuf(per * n_lwe)
    for i, mat in enumerate(brk):
        g.vmp_prepare(pb.hal.VmpPMat(brk_buf, n, brk_size, cols, cols, brk_size, offset=i * per), g.mat_znx_from_numpy(mat))
    brk_g = pb.hal.VmpPMat(brk_buf, n, brk_size, cols, cols, brk_size)
    brk_o, atk_o, tsk_o = prep_o(brk), prep_o(atk), prep_o(tsk)
    atk_g, tsk_g = prep_g(atk), prep_g(tsk)
    msgs = list(range(1 << log_domain))
    lwe = np.stack([SC.noiseless_lwe(rng, m, log_domain, s_lwe, k, lwe_size) for m in msgs])
    lwe_dev = pb.DevBuf(lwe.nbytes)
    lwe_dev.upload(lwe)
    ggsw = circuit.circuit_bootstrap_to_constant(g, lwe_dev, len(msgs), n_lwe, lwe_size, k, brk_g, g.cggi_x_pow_a(), block, atk_g, tsk_g, k, rank,
                                                 dnum_res, res_size, log_domain)
    got = ggsw.download(np.int64, (len(msgs), dnum_res, cols, res_size, cols, n))
    xpa = o.cggi_x_pow_a()
    for b, m in enumerate(msgs):
        want = SC.circuit_bootstrap_to_constant_ref(o, lwe[b], k, brk_o, xpa, block, atk_o, tsk_o, rank, dnum_res, res_size, log_domain, brk_size)
        assert np.array_equal(got[b], want), m
        check_ggsw(got[b], m, s, k, res_size)
