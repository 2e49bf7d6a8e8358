"""Circuit bootstrapping with NOISE-FREE keys (model: circuit_bootstrapping/tests/circuit_bootstrapping.rs:196-228, which
decrypts the GGSW and bounds its noise): an independent restatement of circuit_bootstrap_core over the oracle must turn a noiseless LWE of
m into GGSW(m): row i, column 0 decrypts to m 2^-(i+1)K in the constant coefficient (every other coefficient 0), column c >= 1 to
m s_{c-1} 2^-(i+1)K -- up to the rounding of the trace's right shifts and the final truncation, orders of magnitude below one digit.
CPU only; the device orchestration is compared with this restatement in tests/test_gpu_circuit.py."""
import numpy as np
import pytest

import semantics_circuit as SC
from oracle import pyoracle as O


def check_ggsw(ggsw, m, s, k, res_size):
    """phase(row i, col c) = m * (1 or s_{c-1}) * 2^-(i+1)k up to rounding, for every coefficient.  With noise-free keys the only error left
    is the rounding of ciphertext components at the last kept bit (the right shifts of the log n trace rounds, the final truncations),
    which decryption multiplies by a ternary secret: at most (1 + rank n) / 2 units of 2^-(res_size k) per rounding stage on column 0,
    and column c >= 1 inherits the column-0 error multiplied by s_{c-1} (expand-row).  The bounds below (16 (1 + rank n) units, times
    n / 8 for c >= 1) sit at least five bits under the lowest message bit of the last row."""
    dnum, cols = ggsw.shape[0], ggsw.shape[1]
    n = ggsw.shape[-1]
    tot = res_size * k
    for i in range(dnum):
        for c in range(cols):
            ph = SC.phase_scaled(ggsw[i, c], s, k)
            pt = np.zeros(n, dtype=np.int64)
            if c == 0:
                pt[0] = m
            else:
                pt = s[c - 1] * m
            want = [int(x) << (tot - (i + 1) * k) for x in pt]
            err = max(abs(a - b) for a, b in zip(ph, want))
            tol = 16 * (1 + len(s) * n) * (1 if c == 0 else n // 8)
            assert tol < 1 << (tot - dnum * k - 5)  # the bound itself is meaningful: 1/32 of the lowest message bit
            assert err <= tol, (i, c, err, tol, tot)


@pytest.mark.parametrize("fl", [O.NTT120, O.FFT64])
@pytest.mark.parametrize("rank", [1, 2])
def test_noiseless_circuit_bootstrap_gives_ggsw_of_the_message(fl, rank):
    n, k, n_lwe, block, log_domain = 256, 12, 12, 3, 2
    brk_size, dnum_res, res_size, lwe_size = 4, 2, 4, 2
    rng = np.random.default_rng(3000 + fl + rank)
    o = O.OracleModule(n, fl)
    s_lwe, s, brk, atk, tsk = SC.build_keys(rng, n, k, rank, n_lwe, block, brk_size, brk_size, brk_size, brk_size + 1, res_size, res_size + 1)
    cols = rank + 1

    def prep(mats):
        out = []
        for mat in mats:
            d, ci, sz, co, _ = mat.shape
            pm = o.vmp_pmat_alloc(d, ci, co, sz)
            o.vmp_prepare(pm, mat)
            out.append(pm)
        return out

    brk_o, atk_o, tsk_o = prep(brk), prep(atk), prep(tsk)
    xpa = o.cggi_x_pow_a()
    assert s_lwe.sum() >= 1
    for m in range(1 << log_domain):
        lwe = SC.noiseless_lwe(rng, m, log_domain, s_lwe, k, lwe_size)
        ggsw = SC.circuit_bootstrap_to_constant_ref(o, lwe, k, brk_o, xpa, block, atk_o, tsk_o, rank, dnum_res, res_size, log_domain, brk_size)
        # rows carry m 2^-12 and m 2^-24; the result keeps 48 bits
        check_ggsw(ggsw, m, s, k, res_size)
