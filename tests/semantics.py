"""Exact (Python big-integer / Fraction) models of what the compositions MEAN, written from the reference's call sequences without
going through oracle/core.c -- the independent pin VERDICT r1 asked for ("oracle pin stops at the leaves").

Torus convention (poulpy-hal/src/layouts/vec_znx.rs, base2k representation): a column with limbs d_0 .. d_{s-1} of base 2^K stands for the
torus polynomial sum_j d_j 2^{-(j+1) K} mod 1 (coefficient-wise)."""
from fractions import Fraction

from util import negacyclic_mul


def torus(limbs, k):
    """limbs: (size, n) integers -> list of n Fractions (NOT reduced modulo 1)."""
    size, n = len(limbs), len(limbs[0])
    return [sum(Fraction(int(limbs[j][i]), 1 << ((j + 1) * k)) for j in range(size)) for i in range(n)]


def centred_mod1(x):
    """x mod 1 in [-1/2, 1/2)."""
    y = x - (x.numerator // x.denominator)
    return y - 1 if y >= Fraction(1, 2) else y


def gadget_product_big(a, key, row_cols, first_col, dsize, bound_by_dnum):
    """The integer polynomials res_big[c][t] of gglwe_product_dft (poulpy-core/src/keyswitching/glwe.rs:298-380) / the dsize loop of
    glwe_external_product_internal (external_product/glwe.rs:197-271), from the DEFINITION of vmp_apply_dft_to_dft with a limb offset
    (reference/ntt120/vmp.rs:301-341: res poly (t, c) = sum over rows of a_row (*) pmat[row][limb t + off][c]):

        digit group di in 0..dsize takes the input limbs  l = dsize-1-di + j * dsize  (j = 0, 1, ..), its j-th limb multiplies key row j,
        and the product lands `di` limbs higher (limb_offset = di), i.e. key limb t + di contributes to result limb t.

    a: (a_size, cols_a, n) ints; key: (dnum, cols_in, key_size, cols_out, n); the columns first_col .. first_col+row_cols-1 of `a` are the
    rows.  -> big[c][t] = list of n ints."""
    a_size, n = len(a), len(a[0][0])
    dnum, cols_in, key_size, cols_out = len(key), len(key[0]), len(key[0][0]), len(key[0][0][0])
    assert cols_in == row_cols
    big = [[[0] * n for _ in range(key_size)] for _ in range(cols_out)]
    for di in range(dsize):
        group = (a_size + di) // dsize
        if bound_by_dnum:
            group = min(group, dnum)
        group = min(group, dnum)  # vmp's row_max = min(rows * cols_in, a.cols * a.size): rows beyond the matrix do not exist
        size_di = key_size - max(dsize - di - 2, 0)  # "small optimisation for dsize > 2" (:361)
        for j in range(group):
            l = dsize - 1 - di + j * dsize
            if l >= a_size:
                continue
            for ci in range(row_cols):
                for t in range(size_di):
                    if t + di >= key_size:
                        continue
                    for c in range(cols_out):
                        prod = negacyclic_mul(a[l][first_col + ci], key[j][ci][t + di][c])
                        big[c][t] = [x + y for x, y in zip(big[c][t], prod)]
    return big


def keyswitch_torus(ain, key, key_k, dsize):
    """Torus value (unreduced Fractions) of every output column of glwe_keyswitch BEFORE the final normalisation
    (keyswitching/glwe.rs:207-239): gadget product of the mask columns with the key + the body limbs on column 0
    (vec_znx_big_add_small_assign, limb for limb)."""
    big = gadget_product_big(ain, key, len(ain[0]) - 1, 1, dsize, True)
    key_size = len(key[0][0])
    for l in range(min(len(ain), key_size)):
        big[0][l] = [x + int(y) for x, y in zip(big[0][l], ain[l][0])]
    return [torus(big[c], key_k) for c in range(len(big))]


def external_product_torus(ain, ggsw, ggsw_k, dsize):
    """Same for glwe_external_product (external_product/glwe.rs:197-271): every column of `a` is a row block, no body term, and the
    digit groups are not bounded by dnum (:233)."""
    big = gadget_product_big(ain, ggsw, len(ain[0]), 0, dsize, False)
    return [torus(big[c], ggsw_k) for c in range(len(big))]


def assert_normalised_equals(out, out_k, want_torus, prec_bits, what="", balanced=True):
    """out: (size, n) digits of base 2^out_k.  The torus value equals `want_torus` modulo 1 up to the rounding of a normalisation to
    `prec_bits` bits (|diff| <= 2^-prec_bits; exact when prec_bits is None).  balanced: every digit lies in [-2^(k-1), 2^(k-1)] -- what the
    same-base2k carry chain guarantees; the cross-base2k path of the reference (reference/vec_znx/normalize.rs:151-) re-packs bits and can
    leave a digit a few units outside (its own property tests only bound the torus value), so there only |digit| < 2^k is asserted."""
    half = 1 << (out_k - 1) if balanced else (1 << out_k) - 1
    for j in range(len(out)):
        for v in out[j]:
            assert -half <= int(v) <= half, (what, "digit out of range", j, int(v))
    got = torus(out, out_k)
    tol = Fraction(0) if prec_bits is None else Fraction(1, 1 << prec_bits)
    for i, (g, w) in enumerate(zip(got, want_torus)):
        d = centred_mod1(g - w)
        assert abs(d) <= tol, (what, "coefficient", i, float(d), float(tol))


# ---- noiseless key material (encryption with zero error: poulpy-core/src/encryption/{gglwe.rs, ggsw.rs}) -----------------------------
def digits_of_torus_int(vals, k, size):
    """Balanced base-2^k digits (size limbs, most significant first) of the torus polynomial vals / 2^(size k) mod 1, vals = list of ints."""
    mod = 1 << (size * k)
    out = [[0] * len(vals) for _ in range(size)]
    for i, v in enumerate(vals):
        v %= mod
        for j in range(size - 1, -1, -1):
            d = v & ((1 << k) - 1)
            if d >= 1 << (k - 1):
                d -= 1 << k
            out[j][i] = d
            v = (v - d) >> k
    return out


def noiseless_glwe_row(msg_limbs, masks, secrets, k, size):
    """One noiseless GLWE encryption in limb form: body = msg - sum_c mask_c (*) s_c (torus arithmetic, exact), columns (body, mask_1 ..).
    msg_limbs: (size, n) digits of the message; masks: list (per secret) of (size, n) digit arrays; secrets: list of small int polys.
    -> (size, 1 + len(secrets), n) ints, the body normalised to balanced digits."""
    n = len(secrets[0])
    acc = [0] * n  # message - <mask, s> scaled by 2^(size k)
    for j in range(size):
        w = 1 << ((size - 1 - j) * k)
        for i in range(n):
            acc[i] += int(msg_limbs[j][i]) * w
        for mask, s in zip(masks, secrets):
            prod = negacyclic_mul(mask[j], s)
            for i in range(n):
                acc[i] -= prod[i] * w
    body = digits_of_torus_int(acc, k, size)
    return [[body[j]] + [list(map(int, mask[j])) for mask in masks] for j in range(size)]


def phase(ct, secrets, k):
    """Decryption without rounding: body + sum_c mask_c (*) s_c as a torus polynomial (list of Fractions, unreduced)."""
    size, n = len(ct), len(ct[0][0])
    out = [Fraction(0)] * n
    for j in range(size):
        w = 1 << ((j + 1) * k)
        vals = [int(x) for x in ct[j][0]]
        for c, s in enumerate(secrets):
            prod = negacyclic_mul(ct[j][1 + c], s)
            vals = [x + y for x, y in zip(vals, prod)]
        out = [o + Fraction(v, w) for o, v in zip(out, vals)]
    return out
