"""Front-end serialisation: the reference's little-endian wire format of the coefficient-domain containers, for key and
ciphertext import / export.  Mirrors `WriterTo` / `ReaderFrom` of
  VecZnx    poulpy-hal/src/layouts/vec_znx.rs:339-398    u64 n, cols, size, max_size, byte_len | n*cols*size i64 (limb-major, column-minor)
  MatZnx    poulpy-hal/src/layouts/mat_znx.rs:288-349    u64 n, size, rows, cols_in, cols_out, byte_len | rows*cols_in*size*cols_out*n i64
  ScalarZnx poulpy-hal/src/layouts/scalar_znx.rs:289-337 u64 n, cols, byte_len | cols*n i64
on numpy arrays in the shapes the host mirror uses ((size, cols, n), (rows, cols_in, size, cols_out, n), (cols, n)).  Prepared layouts
(VmpPMat, SvpPPol, VecZnxDft) are backend-private and never serialised by the reference either: keys travel as MatZnx and are
prepared on the device (`Module.vmp_prepare`).  Errors follow the reference: inconsistent metadata and short buffers raise ValueError
(`std::io::ErrorKind::InvalidData`), a truncated stream raises EOFError (`UnexpectedEof`).  The stream is untrusted: before any payload
byte is read the header must describe a plausible container -- n a power of two <= 2^MAX_LOG_N, every dimension non-zero and <= MAX_DIM, the
payload <= MAX_BYTES (or the caller's `capacity_bytes`, the size of the destination buffer the reference reads into) -- otherwise ValueError;
the products are checked in Python integers, so no header can wrap them the way a usize product would."""
import io
import struct

import numpy as np


MAX_LOG_N = 16          # the largest ring degree of the reference's parameter sets (poulpy-bench/src/params.rs: log_n <= 16)
MAX_DIM = 1 << 12       # limbs / rows / columns of one container
MAX_BYTES = 1 << 33     # 8 GiB of payload without an explicit capacity


def _check_header(kind, n, dims, length, capacity_bytes):
    if n == 0 or n & (n - 1) or n > (1 << MAX_LOG_N):
        raise ValueError(f"{kind} header: n={n} is not a power of two <= 2^{MAX_LOG_N}")
    for name, v in dims:
        if v == 0 or v > MAX_DIM:
            raise ValueError(f"{kind} header: {name}={v} outside [1, {MAX_DIM}]")
    limit = MAX_BYTES if capacity_bytes is None else capacity_bytes
    if length > limit:
        raise ValueError(f"{kind} buffer too small: self.data.len()={limit} < read len={length}")


def _w64(w, *vals):
    w.write(struct.pack("<%dQ" % len(vals), *[int(v) for v in vals]))


def _r64(r, count):
    raw = r.read(8 * count)
    if len(raw) != 8 * count:
        raise EOFError("unexpected end of stream in header")
    return struct.unpack("<%dQ" % count, raw)


def _payload(r, length):
    raw = r.read(length)
    if len(raw) != length:
        raise EOFError("unexpected end of stream in payload")
    return np.frombuffer(raw, dtype="<i8").astype(np.int64)


def write_vec_znx(w, a: np.ndarray, max_size=None):
    """a: int64 (size, cols, n)."""
    a = np.ascontiguousarray(a, dtype=np.int64)
    size, cols, n = a.shape
    _w64(w, n, cols, size, size if max_size is None else max_size, a.nbytes)
    w.write(a.astype("<i8").tobytes())


def read_vec_znx(r, capacity_bytes=None):
    """-> (int64 (size, cols, n), max_size).  capacity_bytes = size of the destination buffer (the reference reads in place)."""
    n, cols, size, max_size, length = _r64(r, 5)
    if n * cols * size * 8 != length:
        raise ValueError(f"VecZnx metadata inconsistent: n={n} * cols={cols} * size={size} * 8 = {n * cols * size * 8} != data len={length}")
    _check_header("VecZnx", n, (("cols", cols), ("size", size)), length, capacity_bytes)
    return _payload(r, length).reshape(size, cols, n), max_size


def write_mat_znx(w, a: np.ndarray):
    """a: int64 (rows, cols_in, size, cols_out, n) -- poulpy-hal/src/layouts/mat_znx.rs:161-176."""
    a = np.ascontiguousarray(a, dtype=np.int64)
    rows, cols_in, size, cols_out, n = a.shape
    _w64(w, n, size, rows, cols_in, cols_out, a.nbytes)
    w.write(a.astype("<i8").tobytes())


def read_mat_znx(r, capacity_bytes=None):
    n, size, rows, cols_in, cols_out, length = _r64(r, 6)
    expected = rows * cols_in * n * cols_out * size * 8
    if expected != length:
        raise ValueError(f"MatZnx metadata inconsistent: rows={rows} * cols_in={cols_in} * n={n} * cols_out={cols_out} * size={size} * 8 = "
                         f"{expected} != data len={length}")
    _check_header("MatZnx", n, (("size", size), ("rows", rows), ("cols_in", cols_in), ("cols_out", cols_out)), length, capacity_bytes)
    return _payload(r, length).reshape(rows, cols_in, size, cols_out, n)


def write_scalar_znx(w, a: np.ndarray):
    """a: int64 (cols, n)."""
    a = np.ascontiguousarray(a, dtype=np.int64)
    cols, n = a.shape
    _w64(w, n, cols, a.nbytes)
    w.write(a.astype("<i8").tobytes())


def read_scalar_znx(r, capacity_bytes=None):
    n, cols, length = _r64(r, 3)
    if n * cols * 8 != length:
        raise ValueError(f"ScalarZnx metadata inconsistent: n={n} * cols={cols} * 8 = {n * cols * 8} != data len={length}")
    _check_header("ScalarZnx", n, (("cols", cols),), length, capacity_bytes)
    return _payload(r, length).reshape(cols, n)


def dumps(kind: str, a: np.ndarray) -> bytes:
    w = io.BytesIO()
    {"vec_znx": write_vec_znx, "mat_znx": write_mat_znx, "scalar_znx": write_scalar_znx}[kind](w, a)
    return w.getvalue()


def import_key(module, stream, expect=None):
    """Key import: a serialised MatZnx (the coefficient-domain GGLWE / GGSW key) -> prepared VmpPMat on the module's device.
    `expect` = (rows, cols_in, cols_out, size) of the key the caller is about to use; a stream that describes another layout is rejected."""
    cap = None if expect is None else 8 * module.n * expect[0] * expect[1] * expect[2] * expect[3]
    mat = read_mat_znx(stream, capacity_bytes=cap)
    rows, cols_in, size, cols_out, n = mat.shape
    if n != module.n:
        raise ValueError(f"key ring degree {n} != module ring degree {module.n}")
    if expect is not None and (rows, cols_in, cols_out, size) != tuple(expect):
        raise ValueError(f"key layout (rows, cols_in, cols_out, size) = {(rows, cols_in, cols_out, size)} != expected {tuple(expect)}")
    pmat = module.vmp_pmat_alloc(rows, cols_in, cols_out, size)
    module.vmp_prepare(pmat, module.mat_znx_from_numpy(mat))
    return pmat
