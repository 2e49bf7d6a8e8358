"""Noise-free key material for circuit bootstrapping and an INDEPENDENT restatement of circuit_bootstrap_core(to_exponent = false)
(poulpy-bin-fhe/src/circuit_bootstrapping/circuit.rs:219-380) over oracle primitives -- written from the Rust source, importing nothing
from poulpy_b200.circuit (the earlier exponent / constant tests drove the product's own circuit.py on both sides).

Key formats (zero noise, uniform masks):
  BRK_i = GGSW(s_lwe[i])   poulpy-core/src/encryption/ggsw.rs:62-120: row d, column 0 has phase m 2^-(d+1)K, column c >= 1 phase m s_{c-1} 2^-(d+1)K
  ATK_p = GGLWE(s -> sigma_{p^-1}(s))   encryption/glwe_automorphism_key.rs: row d, input column ci encrypts s_ci 2^-(d+1)K under s(X^(p^-1))
  TSK_i = GGLWE(s_i s_j)   encryption/gglwe_to_ggsw_key.rs:60-106: input column j of key i encrypts s_i s_j under s"""
import numpy as np

from oracle import pyoracle as O
from util import fill_uniform


def negacyclic_np(a, b):
    n = len(a)
    full = np.convolve(np.asarray(a, dtype=np.int64), np.asarray(b, dtype=np.int64))
    res = full[:n].copy()
    res[: n - 1] -= full[n:]
    return res


def automorphism_np(p, a):
    """a(X) -> a(X^p) in Z[X]/(X^n + 1), p odd."""
    n = len(a)
    out = np.zeros(n, dtype=np.int64)
    for j in range(n):
        e = (j * p) % (2 * n)
        if e < n:
            out[e] += a[j]
        else:
            out[e - n] -= a[j]
    return out


def noiseless_row(rng, msg_limbs, secrets, k, size):
    """(size, 1 + rank, n) int64: uniform masks, body = msg - sum mask_c (*) s_c as balanced base-2^k digits (exact torus arithmetic on
    size * k bits).  msg_limbs: dict limb -> integer poly (small coefficients)."""
    n = len(secrets[0])
    masks = [fill_uniform(rng, (size, n), k) for _ in secrets]
    acc = [0] * n
    for j in range(size):
        w = 1 << ((size - 1 - j) * k)
        v = np.zeros(n, dtype=np.int64)
        if j in msg_limbs:
            v = v + np.asarray(msg_limbs[j], dtype=np.int64)
        for mask, s in zip(masks, secrets):
            v = v - negacyclic_np(mask[j], s)
        for i in range(n):
            acc[i] += int(v[i]) * w
    mod = 1 << (size * k)
    out = np.zeros((size, 1 + len(secrets), n), dtype=np.int64)
    for i in range(n):
        v = acc[i] % mod
        for j in range(size - 1, -1, -1):
            d = v & ((1 << k) - 1)
            if d >= 1 << (k - 1):
                d -= 1 << k
            out[j, 0, i] = d
            v = (v - d) >> k
    for c, mask in enumerate(masks):
        out[:, 1 + c, :] = mask
    return out


def phase_scaled(ct, secrets, k):
    """body + sum mask_c (*) s_c as exact integers scaled by 2^(size k), reduced to the centred residue modulo 2^(size k)."""
    size, _, n = ct.shape
    out = [0] * n
    for j in range(size):
        v = ct[j, 0].astype(np.int64).copy()
        for c, s in enumerate(secrets):
            v = v + negacyclic_np(ct[j, 1 + c], s)
        w = 1 << ((size - 1 - j) * k)
        out = [o + int(x) * w for o, x in zip(out, v)]
    mod = 1 << (size * k)
    half = mod >> 1
    return [((o + half) % mod) - half for o in out]


def build_keys(rng, n, k, rank, n_lwe, block, brk_dnum, brk_size, atk_dnum, atk_size, tsk_dnum, tsk_size):
    log_n = n.bit_length() - 1
    cols = rank + 1
    s_lwe = np.zeros(n_lwe, dtype=np.int64)
    for b0 in range(0, n_lwe - block + 1, block):  # binary block: at most one 1 per block (fill_binary_block)
        if rng.integers(0, 4) != 0:
            s_lwe[b0 + rng.integers(0, block)] = 1
    s = [rng.integers(-1, 2, size=n).astype(np.int64) for _ in range(rank)]
    brk = []
    for i in range(n_lwe):
        mat = np.zeros((brk_dnum, cols, brk_size, cols, n), dtype=np.int64)
        one = np.zeros(n, dtype=np.int64)
        one[0] = s_lwe[i]
        for d in range(brk_dnum):
            for c in range(cols):
                pt = one if c == 0 else s[c - 1] * s_lwe[i]
                mat[d, c] = noiseless_row(rng, {d: pt}, s, k, brk_size)
        brk.append(mat)
    atk = []
    for i in range(log_n):
        p = O.trace_galois_element(i, n)
        pinv = pow(p % (2 * n), -1, 2 * n)
        s_out = [automorphism_np(pinv, sc) for sc in s]
        mat = np.zeros((atk_dnum, rank, atk_size, cols, n), dtype=np.int64)
        for d in range(atk_dnum):
            for ci in range(rank):
                mat[d, ci] = noiseless_row(rng, {d: s[ci]}, s_out, k, atk_size)
        atk.append(mat)
    tsk = []
    for i in range(rank):
        mat = np.zeros((tsk_dnum, rank, tsk_size, cols, n), dtype=np.int64)
        for d in range(tsk_dnum):
            for j in range(rank):
                mat[d, j] = noiseless_row(rng, {d: negacyclic_np(s[i], s[j])}, s, k, tsk_size)
        tsk.append(mat)
    return s_lwe, s, brk, atk, tsk


def noiseless_lwe(rng, m, log_domain, s_lwe, k, lwe_size):
    """LWE of m encoded on log_domain + 1 bits (circuit_bootstrapping/tests/circuit_bootstrapping.rs:131-133: one padding bit), zero noise:
    (lwe_size, 1, n_lwe + 1) digits of (b, a_0, ..) with b + <a, s> = m / 2^(log_domain + 1) on lwe_size * k bits."""
    n_lwe = len(s_lwe)
    a = fill_uniform(rng, (lwe_size, n_lwe), k)
    tot = lwe_size * k
    acc = m << (tot - log_domain - 1)
    for j in range(lwe_size):
        acc -= int(np.dot(a[j], s_lwe)) << ((lwe_size - 1 - j) * k)
    acc %= 1 << tot
    out = np.zeros((lwe_size, 1, n_lwe + 1), dtype=np.int64)
    for j in range(lwe_size - 1, -1, -1):
        d = acc & ((1 << k) - 1)
        if d >= 1 << (k - 1):
            d -= 1 << k
        out[j, 0, 0] = d
        acc = (acc - d) >> k
    out[:, 0, 1:] = a
    return out


def circuit_bootstrap_to_constant_ref(o, lwe, k, brk_o, xpa, block, atk_o, tsk_o, rank, dnum_res, res_size, log_domain, brk_size):
    """circuit_bootstrap_core(to_exponent = false, extension_factor = 1), one LWE, every layout in base 2^k.  Line references: circuit.rs."""
    n, cols = o.n, rank + 1
    alpha = 1 << (dnum_res - 1).bit_length() if dnum_res > 1 else 1                       # :259 next_power_of_two
    f = [0] * ((1 << log_domain) * alpha)                                                  # :276
    for j in range(1 << log_domain):                                                       # :283-287
        for i in range(dnum_res):
            f[j * alpha + i] = j * (1 << (k * (dnum_res - 1 - i)))
    # LookupTable::set (lut.rs:271-338), k_lut = k * dnum_res (a multiple of base2k: scale = 1), extension factor 1
    limbs = dnum_res
    step = (n + len(f) // 2) // len(f)                                                     # div_round
    lut_full = np.zeros((limbs, 1, n), dtype=np.int64)
    for i, fi in enumerate(f):
        lut_full[limbs - 1, 0, i * step:(i + 1) * step] = fi
    O.vec_znx_normalize_assign(k, lut_full, 0)
    drift = step >> 1
    lut = np.zeros_like(lut_full)
    O.vec_znx_rotate((-drift) % (2 * n), lut, 0, lut_full, 0)                              # res.rotate(-drift)
    # blind rotation over the BRK layout (:316-327); rotation direction Left (constant mode)
    lwe_2n = O.mod_switch_2n(2 * n, lwe, k, True)
    acc = np.zeros((brk_size, cols, n), dtype=np.int64)
    o.cggi_blind_rotate_block_binary(acc, lwe_2n, lut, brk_o, xpa, block, k)
    gap = 2 * drift                                                                        # :329
    assert gap > 0
    tmp_size = max(brk_size, res_size)
    ggsw = np.zeros((dnum_res, cols, res_size, cols, n), dtype=np.int64)
    for i in range(dnum_res):                                                              # :340-364
        tmp = np.zeros((tmp_size, cols, n), dtype=np.int64)                                # glwe_trace: copy, trace in place, copy out
        tmp[:brk_size] = acc
        o.glwe_trace_assign(tmp, k, 0, atk_o, k, 1)
        ggsw[i, 0] = tmp[:res_size]
        if i + 1 < dnum_res:                                                               # glwe_rotate_assign(-gap)
            rot = np.zeros_like(acc)
            for c in range(cols):
                O.vec_znx_rotate((-gap) % (2 * n), rot, c, acc, c)
            acc = rot
    o.ggsw_expand_row(ggsw, k, tsk_o, k, 1)                                                # :367
    return ggsw
