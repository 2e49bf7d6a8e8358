/*
 * binfhe.c -- CGGI blind rotation (block-binary key distribution) restated over the oracle's HAL
 * functions.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates: poulpy-bin-fhe/src/blind_rotation/algorithms/mod.rs:136-181 (mod_switch_2n, div_round_by_pow2),
 *   poulpy-bin-fhe/src/blind_rotation/utils.rs:6-41 (set_xai_plus_y with y = 0),
 *   poulpy-bin-fhe/src/blind_rotation/algorithms/cggi/key_prepared.rs:66-75 (x_pow_a table),
 *   poulpy-bin-fhe/src/blind_rotation/algorithms/cggi/algorithm.rs:275-368 (execute_block_binary),
 *   :121-273 (execute_block_binary_extended).
 */
#include "poulpy_oracle.h"

#include <assert.h>
#include <stdlib.h>
#include <string.h>

/* algorithms/mod.rs:136-181 */
void orc_mod_switch_2n(size_t n, int64_t *res, const orc_vec_znx *lwe, size_t base2k, int rot_left) {
    size_t len = lwe->n;
    size_t log2n = 0;
    while (((size_t)1 << log2n) < n) log2n++; /* usize::BITS - (n-1).leading_zeros() */
    log2n += 1;
    memcpy(res, lwe->data, 8 * len); /* at(0, 0) */
    if (rot_left)
        for (size_t i = 0; i < len; i++) res[i] = -res[i];
    if (base2k > log2n) {
        size_t diff = base2k - (log2n - 1);
        for (size_t i = 0; i < len; i++) res[i] = (res[i] + ((int64_t)1 << (diff - 1))) >> diff;
    } else {
        size_t rem = base2k - (log2n % base2k);
        size_t size = (log2n + base2k - 1) / base2k;
        for (size_t i = 1; i < size; i++) {
            const int64_t *x = lwe->data + lwe->n * (i * lwe->cols);
            if (i == size - 1 && rem != base2k) {
                size_t k_rem = base2k - rem;
                for (size_t j = 0; j < len; j++) res[j] = (int64_t)((uint64_t)res[j] << k_rem) + (x[j] >> rem);
            } else {
                for (size_t j = 0; j < len; j++) res[j] = (int64_t)((uint64_t)res[j] << base2k) + x[j];
            }
        }
    }
}

static size_t prep_bytes(int flavour) { return flavour == 0 ? 32u : 8u; }

/* key_prepared.rs:66-75 + utils.rs:6-41 */
void orc_cggi_x_pow_a(int flavour, const void *mod, orc_svp_ppol *res) {
    size_t n = res->n;
    assert(res->cols == 2 * n);
    int64_t *buf = (int64_t *)calloc(n, 8);
    orc_scalar_znx sz = {buf, n, 1};
    for (size_t ai = 0; ai < 2 * n; ai++) {
        if (ai < n) buf[ai] = 1;
        else buf[(ai - n) & (n - 1)] = -1;
        if (flavour == 0) orc_ntt120_svp_prepare((const orc_ntt120_module *)mod, res, ai, &sz, 0);
        else orc_fft64_svp_prepare((const orc_fft64_module *)mod, res, ai, &sz, 0);
        if (ai < n) buf[ai] = 0;
        else buf[(ai - n) & (n - 1)] = 0;
        buf[0] = 0;
    }
    free(buf);
}

/* algorithm.rs:275-368 */
void orc_cggi_blind_rotate_block_binary(int flavour, const void *mod, orc_vec_znx *res, const int64_t *lwe_2n,
                                        size_t n_lwe, const orc_vec_znx *lut, const orc_vmp_pmat *brk,
                                        const orc_svp_ppol *x_pow_a, size_t block_size, size_t base2k) {
    size_t n = res->n, cols = res->cols, two_n = 2 * n;
    size_t dnum = brk[0].rows, bsize = brk[0].size;
    size_t pb = prep_bytes(flavour), bb = flavour == 0 ? 16u : 8u;
    const int64_t *a = lwe_2n + 1;
    int64_t b = lwe_2n[0];

    memset(res->data, 0, 8 * n * cols * res->size);
    orc_vec_znx_rotate(b, res, 0, lut, 0);

    orc_vec_znx_dft acc_dft = {calloc(n * cols * dnum, pb), n, cols, dnum};
    orc_vec_znx_dft vmp_res = {calloc(n * cols * bsize, pb), n, cols, bsize};
    orc_vec_znx_dft acc_add = {calloc(n * cols * bsize, pb), n, cols, bsize};
    orc_vec_znx_dft vmp_xai = {calloc(n * bsize, pb), n, 1, bsize};
    orc_vec_znx_big acc_big = {calloc(n * bsize, bb), n, 1, bsize};

    for (size_t blk = 0; blk + block_size <= n_lwe; blk += block_size) { /* chunks_exact */
        for (size_t j = 0; j < cols; j++) {
            if (flavour == 0) {
                orc_ntt120_vec_znx_dft_apply((const orc_ntt120_module *)mod, 1, 0, &acc_dft, j, res, j);
                orc_ntt120_vec_znx_dft_zero(&acc_add, j);
            } else {
                orc_fft64_vec_znx_dft_apply((const orc_fft64_module *)mod, 1, 0, &acc_dft, j, res, j);
                orc_fft64_vec_znx_dft_zero(&acc_add, j);
            }
        }
        for (size_t t = 0; t < block_size; t++) {
            int64_t aii = a[blk + t];
            size_t ai_pos = (size_t)((aii + (int64_t)two_n) & (int64_t)(two_n - 1));
            const orc_vmp_pmat *sk = &brk[blk + t];
            if (flavour == 0) {
                const orc_ntt120_module *m = (const orc_ntt120_module *)mod;
                orc_ntt120_vmp_apply_dft_to_dft(m, &vmp_res, &acc_dft, sk, 0);
                for (size_t i = 0; i < cols; i++) {
                    orc_ntt120_svp_apply_dft_to_dft(m, &vmp_xai, 0, x_pow_a, ai_pos, &vmp_res, i);
                    orc_ntt120_vec_znx_dft_add_assign(&acc_add, i, &vmp_xai, 0);
                    orc_ntt120_vec_znx_dft_sub_assign(&acc_add, i, &vmp_res, i);
                }
            } else {
                const orc_fft64_module *m = (const orc_fft64_module *)mod;
                orc_fft64_vmp_apply_dft_to_dft(m, &vmp_res, &acc_dft, sk, 0);
                for (size_t i = 0; i < cols; i++) {
                    orc_fft64_svp_apply_dft_to_dft(m, &vmp_xai, 0, x_pow_a, ai_pos, &vmp_res, i);
                    orc_fft64_vec_znx_dft_add_assign(&acc_add, i, &vmp_xai, 0);
                    orc_fft64_vec_znx_dft_sub_assign(&acc_add, i, &vmp_res, i);
                }
            }
        }
        for (size_t i = 0; i < cols; i++) {
            if (flavour == 0) {
                orc_ntt120_vec_znx_idft_apply((const orc_ntt120_module *)mod, &acc_big, 0, &acc_add, i);
                orc_ntt120_vec_znx_big_add_small_assign(&acc_big, 0, res, i);
                orc_ntt120_vec_znx_big_normalize(res, base2k, 0, i, &acc_big, base2k, 0, 0);
            } else {
                orc_fft64_vec_znx_idft_apply((const orc_fft64_module *)mod, &acc_big, 0, &acc_add, i);
                orc_fft64_vec_znx_big_add_small_assign(&acc_big, 0, res, i);
                orc_fft64_vec_znx_big_normalize(res, base2k, 0, i, &acc_big, base2k, 0, 0);
            }
        }
    }
    free(acc_dft.data);
    free(vmp_res.data);
    free(acc_add.data);
    free(vmp_xai.data);
    free(acc_big.data);
}

/* algorithm.rs:370-443 (execute_standard, block_size == 1): acc = X^b LUT; for every LWE coefficient
 *   acc_tmp = acc (x) BRK_i (glwe_external_product); acc_tmp *= X^{a_i} - 1; acc += acc_tmp;   then one glwe_normalize_assign.
 * acc and acc_tmp share layout and base2k with `res` (take_glwe(&out_mut), :423). */
void orc_cggi_blind_rotate_standard(int flavour, const void *mod, orc_vec_znx *res, size_t res_base2k, const int64_t *lwe_2n,
                                    size_t n_lwe, const orc_vec_znx *lut, const orc_vmp_pmat *brk, size_t brk_base2k) {
    size_t n = res->n, cols = res->cols;
    const int64_t *a = lwe_2n + 1;
    memset(res->data, 0, 8 * n * cols * res->size);
    orc_vec_znx_rotate(lwe_2n[0], res, 0, lut, 0);
    orc_vec_znx tmp = {(int64_t *)calloc(n * cols * res->size, 8), n, cols, res->size};
    for (size_t i = 0; i < n_lwe; i++) {
        orc_glwe_external_product(flavour, mod, &tmp, res_base2k, res, res_base2k, &brk[i], brk_base2k, 1);
        for (size_t c = 0; c < cols; c++) orc_vec_znx_mul_xp_minus_one_assign(a[i], &tmp, c); /* operations/glwe.rs:1051-1062 */
        for (size_t c = 0; c < cols; c++) orc_vec_znx_add_assign(res, c, &tmp, c);             /* api/operations.rs:273-289 */
    }
    for (size_t c = 0; c < cols; c++) orc_vec_znx_normalize_assign(res_base2k, res, c);         /* operations/glwe.rs:1312-1327 */
    free(tmp.data);
}

/* ---- flavour dispatch for the extended variant ------------------------------------------------------------------------------------- */
#define FL(f0, f1, ...) do { if (flavour == 0) f0(__VA_ARGS__); else f1(__VA_ARGS__); } while (0)

/* algorithm.rs:121-273 (execute_block_binary_extended): the accumulator lives in `ext` interleaved rings of degree n (domain size n * ext);
 * lwe_2n is mod-switched to 2 * n * ext; lut: `ext` VecZnx(1 col); res receives acc[0] (res.size limbs). */
void orc_cggi_blind_rotate_block_binary_extended(int flavour, const void *mod, orc_vec_znx *res, const int64_t *lwe_2n, size_t n_lwe,
                                                 const orc_vec_znx *lut, size_t ext, const orc_vmp_pmat *brk, const orc_svp_ppol *x_pow_a,
                                                 size_t block_size, size_t base2k) {
    size_t n = res->n, cols = res->cols, two_n = 2 * n, two_n_ext = 2 * n * ext;
    size_t dnum = brk[0].rows, bsize = brk[0].size;
    size_t pb = prep_bytes(flavour), bb = flavour == 0 ? 16u : 8u;
    const orc_ntt120_module *m0 = (const orc_ntt120_module *)mod;
    const orc_fft64_module *m1 = (const orc_fft64_module *)mod;
    orc_vec_znx *acc = (orc_vec_znx *)calloc(ext, sizeof *acc);
    orc_vec_znx_dft *acc_dft = (orc_vec_znx_dft *)calloc(ext, sizeof *acc_dft), *vmp_res = (orc_vec_znx_dft *)calloc(ext, sizeof *vmp_res),
                    *acc_add = (orc_vec_znx_dft *)calloc(ext, sizeof *acc_add);
    for (size_t i = 0; i < ext; i++) {
        acc[i] = (orc_vec_znx){(int64_t *)calloc(n * cols * res->size, 8), n, cols, res->size};
        acc_dft[i] = (orc_vec_znx_dft){calloc(n * cols * dnum, pb), n, cols, dnum};
        vmp_res[i] = (orc_vec_znx_dft){calloc(n * cols * bsize, pb), n, cols, bsize};
        acc_add[i] = (orc_vec_znx_dft){calloc(n * cols * bsize, pb), n, cols, bsize};
    }
    orc_vec_znx_dft vmp_xai = {calloc(n * bsize, pb), n, 1, bsize};
    orc_vec_znx_big acc_big = {calloc(n * bsize, bb), n, 1, bsize};

    const int64_t *a = lwe_2n + 1;
    size_t b_pos = (size_t)((lwe_2n[0] + (int64_t)two_n_ext) & (int64_t)(two_n_ext - 1));
    size_t b_hi = b_pos / ext, b_lo = b_pos & (ext - 1);
    for (size_t i = 0; i < b_lo; i++) orc_vec_znx_rotate((int64_t)b_hi + 1, &acc[i], 0, &lut[ext - b_lo + i], 0);   /* :185-187 */
    for (size_t i = b_lo; i < ext; i++) orc_vec_znx_rotate((int64_t)b_hi, &acc[i], 0, &lut[i - b_lo], 0);            /* :188-190 */

    for (size_t blk = 0; blk + block_size <= n_lwe; blk += block_size) {
        for (size_t i = 0; i < ext; i++)
            for (size_t j = 0; j < cols; j++) {
                if (flavour == 0) { orc_ntt120_vec_znx_dft_apply(m0, 1, 0, &acc_dft[i], j, &acc[i], j); orc_ntt120_vec_znx_dft_zero(&acc_add[i], j); }
                else { orc_fft64_vec_znx_dft_apply(m1, 1, 0, &acc_dft[i], j, &acc[i], j); orc_fft64_vec_znx_dft_zero(&acc_add[i], j); }
            }
        for (size_t t = 0; t < block_size; t++) {
            int64_t aii = a[blk + t];
            size_t ai_pos = (size_t)((aii + (int64_t)two_n_ext) & (int64_t)(two_n_ext - 1));
            size_t ai_hi = ai_pos / ext, ai_lo = ai_pos & (ext - 1);
            const orc_vmp_pmat *sk = &brk[blk + t];
            for (size_t i = 0; i < ext; i++) {
                if (flavour == 0) orc_ntt120_vmp_apply_dft_to_dft(m0, &vmp_res[i], &acc_dft[i], sk, 0);
                else orc_fft64_vmp_apply_dft_to_dft(m1, &vmp_res[i], &acc_dft[i], sk, 0);
            }
/* acc_add[I][k] += x_pow_a[IDX] * vmp_res[J][k] - vmp_res[I][k] */
#define XAI(I, J, IDX)                                                                                                       \
    for (size_t k = 0; k < cols; k++) {                                                                                      \
        if (flavour == 0) {                                                                                                  \
            orc_ntt120_svp_apply_dft_to_dft(m0, &vmp_xai, 0, x_pow_a, (IDX), &vmp_res[(J)], k);                              \
            orc_ntt120_vec_znx_dft_add_assign(&acc_add[(I)], k, &vmp_xai, 0);                                                \
            orc_ntt120_vec_znx_dft_sub_assign(&acc_add[(I)], k, &vmp_res[(I)], k);                                           \
        } else {                                                                                                             \
            orc_fft64_svp_apply_dft_to_dft(m1, &vmp_xai, 0, x_pow_a, (IDX), &vmp_res[(J)], k);                               \
            orc_fft64_vec_znx_dft_add_assign(&acc_add[(I)], k, &vmp_xai, 0);                                                 \
            orc_fft64_vec_znx_dft_sub_assign(&acc_add[(I)], k, &vmp_res[(I)], k);                                            \
        }                                                                                                                    \
    }
            if (ai_lo == 0) { /* :216-228 */
                if (ai_hi != 0)
                    for (size_t j = 0; j < ext; j++) XAI(j, j, ai_hi)
            } else {          /* :235-258 */
                if (((ai_hi + 1) & (two_n - 1)) != 0)
                    for (size_t i = 0; i < ai_lo; i++) XAI(i, ext - ai_lo + i, ai_hi + 1)
                if (ai_hi != 0)
                    for (size_t i = ai_lo; i < ext; i++) XAI(i, i - ai_lo, ai_hi)
            }
#undef XAI
        }
        for (size_t j = 0; j < ext; j++)
            for (size_t i = 0; i < cols; i++) {
                if (flavour == 0) {
                    orc_ntt120_vec_znx_idft_apply(m0, &acc_big, 0, &acc_add[j], i);
                    orc_ntt120_vec_znx_big_add_small_assign(&acc_big, 0, &acc[j], i);
                    orc_ntt120_vec_znx_big_normalize(&acc[j], base2k, 0, i, &acc_big, base2k, 0, 0);
                } else {
                    orc_fft64_vec_znx_idft_apply(m1, &acc_big, 0, &acc_add[j], i);
                    orc_fft64_vec_znx_big_add_small_assign(&acc_big, 0, &acc[j], i);
                    orc_fft64_vec_znx_big_normalize(&acc[j], base2k, 0, i, &acc_big, base2k, 0, 0);
                }
            }
    }
    memcpy(res->data, acc[0].data, 8 * n * cols * res->size); /* vec_znx_copy of every column (:268-270) */
    for (size_t i = 0; i < ext; i++) {
        free(acc[i].data); free(acc_dft[i].data); free(vmp_res[i].data); free(acc_add[i].data);
    }
    free(acc); free(acc_dft); free(vmp_res); free(acc_add); free(vmp_xai.data); free(acc_big.data);
}
