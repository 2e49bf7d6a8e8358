"""Shared helpers for the parity tests (seeded inputs in the reference's distributions)."""
import numpy as np


def fill_uniform(rng, shape, log_bound):
    """poulpy-hal/src/layouts/vec_znx.rs:283-295: uniform in [-2^(log_bound-1), 2^(log_bound-1))."""
    if log_bound == 64:
        return rng.integers(-(1 << 63), (1 << 63) - 1, size=shape, dtype=np.int64, endpoint=True)
    return rng.integers(-(1 << (log_bound - 1)), 1 << (log_bound - 1), size=shape, dtype=np.int64)


def bitrev(j, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (j & 1)
        j >>= 1
    return r


def negacyclic_mul(a, b):
    """Schoolbook product in Z[X]/(X^n+1) on Python ints."""
    n = len(a)
    res = [0] * n
    for i, ai in enumerate(a):
        ai = int(ai)
        if ai == 0:
            continue
        for j, bj in enumerate(b):
            k = i + j
            if k < n:
                res[k] += ai * int(bj)
            else:
                res[k - n] -= ai * int(bj)
    return res
