"""Golden fixtures (tests/golden/*.npz): seeded inputs and the oracle's outputs for one small instance of every hot-path composition, in
both flavours.  The reference holds no stored byte vectors for this path (its tests are KATs and property tests,
restated in test_oracle_kat.py) and cannot be built in this image, so these files are ORACLE-generated (scripts/make_golden.py): they pin
the oracle against drift (test_golden.py, CPU) and give the CUDA path committed bytes to reproduce (test_golden.py, -m gpu).

Each case: inputs(rng) -> dict of int64 arrays; oracle(flavour, inp) -> int64 array; gpu(flavour, inp) -> int64 array."""
import numpy as np

from util import fill_uniform

N = 256


def _k(fl):
    return 12 if fl == 1 else 18  # the first FFT64 fixtures use base2k 12; the *_k18 cases below take the bench base2k in FFT64 too


_K_OVERRIDE = [None]  # set by the *_k18 cases (FFT64 at the bench base2k = 18: |values| < 2^45 at N = 256, exact after rounding)


def _kk(fl):
    return _K_OVERRIDE[0] if _K_OVERRIDE[0] is not None else _k(fl)


def _with_k(k, fn):
    def run(*a):
        _K_OVERRIDE[0] = k
        try:
            return fn(*a)
        finally:
            _K_OVERRIDE[0] = None
    return run


# ---- GLWE key-switch (rank 1 -> 1, 3 limbs, key of 3 rows x 4 limbs) and GGSW x GLWE external product ------------------------------
def ks_inputs(rng, fl, cols_in):
    k = _kk(fl)
    return {"key": fill_uniform(rng, (3, cols_in, 4, 2, N), k), "a": fill_uniform(rng, (2, 3, 2, N), k)}


def ks_oracle(fl, inp, ext):
    from oracle import pyoracle as O
    o, k = O.OracleModule(N, fl), _kk(fl)
    cols_in = inp["key"].shape[1]
    pm = o.vmp_pmat_alloc(3, cols_in, 2, 4)
    o.vmp_prepare(pm, inp["key"])
    res = np.zeros((2, 3, 2, N), dtype=np.int64)
    (o.glwe_external_product_batch if ext else o.glwe_keyswitch_batch)(res, k, inp["a"], k, pm, k, 1)
    return res


def ks_gpu(fl, inp, ext):
    import poulpy_b200 as pb
    g, k = pb.Module(N, fl), _kk(fl)
    cols_in = inp["key"].shape[1]
    pm = g.vmp_pmat_alloc(3, cols_in, 2, 4)
    g.vmp_prepare(pm, g.mat_znx_from_numpy(inp["key"]))
    res = g.vec_znx_alloc(2, 3, 2)
    (g.glwe_external_product if ext else g.glwe_keyswitch)(res, k, g.vec_znx_from_numpy(inp["a"]), k, pm, k, 1)
    g.sync()
    return g.vec_znx_to_numpy(res)


# ---- CGGI block-binary blind rotation (rank 1, one-limb accumulator, 6 LWE coefficients in blocks of 3) ------------------------------
def br_inputs(rng, fl):
    k = _kk(fl)
    return {"brk": fill_uniform(rng, (6, 1, 2, 2, 2, N), k), "lut": fill_uniform(rng, (1, 1, N), k),
            "lwe": rng.integers(-N, N, size=(2, 7), dtype=np.int64)}


def br_oracle(fl, inp):
    from oracle import pyoracle as O
    o, k = O.OracleModule(N, fl), _kk(fl)
    brk = []
    for mat in inp["brk"]:
        pm = o.vmp_pmat_alloc(1, 2, 2, 2)
        o.vmp_prepare(pm, mat)
        brk.append(pm)
    res = np.zeros((2, 1, 2, N), dtype=np.int64)
    for b in range(2):
        o.cggi_blind_rotate_block_binary(res[b], inp["lwe"][b], inp["lut"], brk, o.cggi_x_pow_a(), 3, k)
    return res


def br_gpu(fl, inp):
    import poulpy_b200 as pb
    g, k = pb.Module(N, fl), _kk(fl)
    per = N * 2 * 2 * 2 * g.prep_bytes
    buf = pb.DevBuf(per * 6)
    for i, mat in enumerate(inp["brk"]):
        g.vmp_prepare(pb.hal.VmpPMat(buf, N, 1, 2, 2, 2, offset=i * per), g.mat_znx_from_numpy(mat))
    lwe = pb.DevBuf(inp["lwe"].nbytes)
    lwe.upload(inp["lwe"])
    res = g.vec_znx_alloc(2, 1, 2)
    g.cggi_blind_rotate(res, lwe, 6, g.vec_znx_from_numpy(inp["lut"]), pb.hal.VmpPMat(buf, N, 1, 2, 2, 2), g.cggi_x_pow_a(), 3, k)
    g.sync()
    return g.vec_znx_to_numpy(res)


# ---- GLWE trace (8 automorphism keys of 2 rows x 3 limbs) -----------------------------------------------------------------------------
def tr_inputs(rng, fl):
    k = _k(fl)
    return {"keys": fill_uniform(rng, (8, 2, 1, 3, 2, N), k), "a": fill_uniform(rng, (2, 2, 2, N), k)}


def tr_oracle(fl, inp):
    from oracle import pyoracle as O
    o, k = O.OracleModule(N, fl), _k(fl)
    keys = []
    for mat in inp["keys"]:
        pm = o.vmp_pmat_alloc(2, 1, 2, 3)
        o.vmp_prepare(pm, mat)
        keys.append(pm)
    res = inp["a"].copy()
    for b in range(res.shape[0]):
        o.glwe_trace_assign(res[b], k, 0, keys, k, 1)
    return res


def tr_gpu(fl, inp):
    import poulpy_b200 as pb
    g, k = pb.Module(N, fl), _k(fl)
    keys = []
    for mat in inp["keys"]:
        pm = g.vmp_pmat_alloc(2, 1, 2, 3)
        g.vmp_prepare(pm, g.mat_znx_from_numpy(mat))
        keys.append(pm)
    res = g.vec_znx_from_numpy(inp["a"])
    g.glwe_trace_assign(res, k, 0, keys, k, 1)
    g.sync()
    return g.vec_znx_to_numpy(res)


CASES = {}
for _fl, _nm in ((0, "ntt120"), (1, "fft64")):
    CASES[f"glwe_keyswitch_{_nm}"] = (lambda rng, fl=_fl: ks_inputs(rng, fl, 1), lambda inp, fl=_fl: ks_oracle(fl, inp, False),
                                      lambda inp, fl=_fl: ks_gpu(fl, inp, False))
    CASES[f"glwe_external_product_{_nm}"] = (lambda rng, fl=_fl: ks_inputs(rng, fl, 2), lambda inp, fl=_fl: ks_oracle(fl, inp, True),
                                             lambda inp, fl=_fl: ks_gpu(fl, inp, True))
    CASES[f"cggi_blind_rotate_{_nm}"] = (lambda rng, fl=_fl: br_inputs(rng, fl), lambda inp, fl=_fl: br_oracle(fl, inp),
                                         lambda inp, fl=_fl: br_gpu(fl, inp))
    CASES[f"glwe_trace_{_nm}"] = (lambda rng, fl=_fl: tr_inputs(rng, fl), lambda inp, fl=_fl: tr_oracle(fl, inp),
                                  lambda inp, fl=_fl: tr_gpu(fl, inp))

# FFT64 at the bench base2k ("add an FFT64 golden case at base2k 18")
CASES["glwe_keyswitch_fft64_k18"] = (_with_k(18, lambda rng: ks_inputs(rng, 1, 1)), _with_k(18, lambda inp: ks_oracle(1, inp, False)),
                                     _with_k(18, lambda inp: ks_gpu(1, inp, False)))
CASES["glwe_external_product_fft64_k18"] = (_with_k(18, lambda rng: ks_inputs(rng, 1, 2)), _with_k(18, lambda inp: ks_oracle(1, inp, True)),
                                            _with_k(18, lambda inp: ks_gpu(1, inp, True)))
CASES["cggi_blind_rotate_fft64_k18"] = (_with_k(18, lambda rng: br_inputs(rng, 1)), _with_k(18, lambda inp: br_oracle(1, inp)),
                                        _with_k(18, lambda inp: br_gpu(1, inp)))
