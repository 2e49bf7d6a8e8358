"""Wire format of the coefficient-domain containers (poulpy-hal/src/layouts/{vec_znx,mat_znx,scalar_znx}.rs ReaderFrom / WriterTo):
header words, payload order, round trips and the reference's error cases.  The key-import path (serialised MatZnx -> vmp_prepare on the
device) is checked against the oracle in the GPU test."""
import io
import struct

import numpy as np
import pytest

from poulpy_b200 import serialize as S


def test_vec_znx_wire_format():
    a = np.arange(2 * 3 * 4, dtype=np.int64).reshape(2, 3, 4) - 7  # (size, cols, n)
    raw = S.dumps("vec_znx", a)
    assert struct.unpack("<5Q", raw[:40]) == (4, 3, 2, 2, 2 * 3 * 4 * 8)  # n, cols, size, max_size, byte length (vec_znx.rs:379-394)
    # limb-major, column-minor: limb j of column i at scalar offset n * (j * cols + i) (znx_base.rs:74)
    flat = np.frombuffer(raw[40:], dtype="<i8")
    assert np.array_equal(flat[4 * (1 * 3 + 2):4 * (1 * 3 + 2) + 4], a[1, 2])
    back, max_size = S.read_vec_znx(io.BytesIO(raw))
    assert max_size == 2 and np.array_equal(back, a)


def test_mat_and_scalar_round_trip():
    rng = np.random.default_rng(1)
    m = rng.integers(-(1 << 62), 1 << 62, size=(3, 2, 4, 2, 8), dtype=np.int64)  # rows, cols_in, size, cols_out, n
    raw = S.dumps("mat_znx", m)
    assert struct.unpack("<6Q", raw[:48]) == (8, 4, 3, 2, 2, m.nbytes)  # n, size, rows, cols_in, cols_out, len (mat_znx.rs:327-345)
    assert np.array_equal(S.read_mat_znx(io.BytesIO(raw)), m)
    s = rng.integers(-1, 2, size=(2, 16), dtype=np.int64)
    raw = S.dumps("scalar_znx", s)
    assert struct.unpack("<3Q", raw[:24]) == (16, 2, s.nbytes)
    assert np.array_equal(S.read_scalar_znx(io.BytesIO(raw)), s)


def test_reader_errors():
    a = np.zeros((2, 1, 8), dtype=np.int64)
    raw = bytearray(S.dumps("vec_znx", a))
    with pytest.raises(EOFError):
        S.read_vec_znx(io.BytesIO(bytes(raw[:-1])))  # truncated payload
    with pytest.raises(EOFError):
        S.read_vec_znx(io.BytesIO(bytes(raw[:30])))  # truncated header
    with pytest.raises(ValueError, match="buffer too small"):
        S.read_vec_znx(io.BytesIO(bytes(raw)), capacity_bytes=a.nbytes - 8)
    raw[16:24] = struct.pack("<Q", 3)  # size no longer matches the byte length (vec_znx.rs:348-356)
    with pytest.raises(ValueError, match="metadata inconsistent"):
        S.read_vec_znx(io.BytesIO(bytes(raw)))
    m = bytearray(S.dumps("mat_znx", np.zeros((1, 1, 1, 1, 8), dtype=np.int64)))
    m[8:16] = struct.pack("<Q", 2)
    with pytest.raises(ValueError, match="MatZnx metadata inconsistent"):
        S.read_mat_znx(io.BytesIO(bytes(m)))


def test_untrusted_headers_are_rejected_before_the_payload_is_read():
    """A self-consistent but implausible header must fail with ValueError without touching the payload (no allocation of `len` bytes)."""

    class Tripwire(io.BytesIO):
        def read(self, size=-1):
            assert size <= 64, "payload read attempted"
            return super().read(size)

    huge = struct.pack("<5Q", 1 << 16, 1 << 12, 1 << 12, 1 << 12, (1 << 16) * (1 << 24) * 8)      # 8 TiB, self-consistent
    with pytest.raises(ValueError, match="buffer too small"):
        S.read_vec_znx(Tripwire(huge))
    for hdr in (struct.pack("<5Q", 24, 1, 1, 1, 24 * 8),                 # n not a power of two
                struct.pack("<5Q", 0, 1, 1, 1, 0),                       # n = 0
                struct.pack("<5Q", 1 << 20, 1, 1, 1, (1 << 20) * 8),     # n beyond any parameter set
                struct.pack("<5Q", 8, 0, 1, 1, 0),                       # zero columns
                struct.pack("<5Q", 8, 1 << 40, 1, 1, (1 << 40) * 64),    # absurd column count
                struct.pack("<5Q", 8, 1 << 61, 8, 8, 0)):                # a product that wraps a usize to 0: checked in Python integers
        with pytest.raises(ValueError):
            S.read_vec_znx(Tripwire(hdr))
    with pytest.raises(ValueError):
        S.read_mat_znx(Tripwire(struct.pack("<6Q", 8, 1, 0, 1, 1, 0)))   # zero rows
    with pytest.raises(ValueError):
        S.read_scalar_znx(Tripwire(struct.pack("<3Q", 12, 1, 96)))


@pytest.mark.gpu
def test_import_key_rejects_an_unexpected_layout():
    import poulpy_b200 as pb

    g = pb.Module(1024, pb.NTT120)
    mat = np.zeros((3, 1, 4, 2, 1024), dtype=np.int64)
    with pytest.raises(ValueError, match="buffer too small|key layout"):
        S.import_key(g, io.BytesIO(S.dumps("mat_znx", mat)), expect=(2, 1, 2, 4))
    with pytest.raises(ValueError, match="key layout"):
        S.import_key(g, io.BytesIO(S.dumps("mat_znx", mat)), expect=(4, 1, 2, 3))
    with pytest.raises(ValueError, match="ring degree"):
        S.import_key(pb.Module(2048, pb.NTT120), io.BytesIO(S.dumps("mat_znx", mat)))
    S.import_key(g, io.BytesIO(S.dumps("mat_znx", mat)), expect=(3, 1, 2, 4))


@pytest.mark.gpu
def test_import_key_matches_oracle_prepare():
    import poulpy_b200 as pb
    from oracle import pyoracle as O

    n = 1024
    rng = np.random.default_rng(2)
    mat = rng.integers(-(1 << 17), 1 << 17, size=(3, 1, 4, 2, n), dtype=np.int64)
    g, o = pb.Module(n, pb.NTT120), O.OracleModule(n, pb.NTT120)
    pg = S.import_key(g, io.BytesIO(S.dumps("mat_znx", mat)))
    po = o.vmp_pmat_alloc(3, 1, 2, 4)
    o.vmp_prepare(po, mat)
    a = rng.integers(-(1 << 17), 1 << 17, size=(2, 3, 2, n), dtype=np.int64)
    want = np.zeros((2, 3, 2, n), dtype=np.int64)
    o.glwe_keyswitch_batch(want, 18, a, 18, po, 18)
    res = g.vec_znx_alloc(2, 3, 2)
    g.glwe_keyswitch(res, 18, g.vec_znx_from_numpy(a), 18, pg, 18)
    g.sync()
    assert np.array_equal(g.vec_znx_to_numpy(res), want)
