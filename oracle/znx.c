/*
 * znx.c -- coefficient-domain (i64) helpers of poulpy-cpu-ref restated in C.
 * TEST INFRASTRUCTURE ONLY (see poulpy_oracle.h).
 *
 * Restates: reference/vec_znx/normalize.rs:18-426 (through normalize_impl.inc
 * instantiated for int64_t, steps of reference/znx/normalization.rs:4-323),
 * reference/znx/rotate.rs:3-26, reference/vec_znx/rotate.rs:9-38.
 */
#include "poulpy_oracle.h"

#include <stdlib.h>
#include <string.h>

static inline int64_t *znx_at(const orc_vec_znx *v, size_t col, size_t limb) {
    return v->data + v->n * (limb * v->cols + col);
}

#define NT int64_t
#define NUT uint64_t
#define NBITS 64
#define NBIG orc_vec_znx
#define NLIMB(v, c, l) znx_at(v, c, l)
#define NF(name) name##_i64
#include "normalize_impl.inc"
#undef NT
#undef NUT
#undef NBITS
#undef NBIG
#undef NLIMB
#undef NF

/* reference/vec_znx/normalize.rs:18-50.  The reference only has op = 0 for i64. */
void orc_vec_znx_normalize(orc_vec_znx *res, size_t res_base2k, int64_t res_offset, size_t res_col,
                           const orc_vec_znx *a, size_t a_base2k, size_t a_col, int op) {
    int64_t *scratch = (int64_t *)malloc(3 * res->n * sizeof(int64_t));
    if (res_base2k == a_base2k)
        normalize_inter_i64(res_base2k, res, res_offset, res_col, a, a_col, scratch, op);
    else
        normalize_cross_i64(res, res_base2k, res_offset, res_col, a, a_base2k, a_col, scratch, op);
    free(scratch);
}

/* reference/znx/rotate.rs:3-26 */
void orc_znx_rotate(int64_t p, int64_t *res, const int64_t *src, size_t n) {
    size_t mp_2n = (size_t)(p & (int64_t)(2 * n - 1));
    size_t mp_1n = mp_2n & (n - 1);
    size_t mp_1n_neg = n - mp_1n;
    int neg_first = mp_2n < n;
    /* dst1 = res[..mp_1n], dst2 = res[mp_1n..]; src1 = src[..mp_1n_neg], src2 = src[mp_1n_neg..] */
    for (size_t i = 0; i < mp_1n; i++) {
        int64_t v = src[mp_1n_neg + i];
        res[i] = neg_first ? (int64_t)(0 - (uint64_t)v) : v;
    }
    for (size_t i = 0; i < mp_1n_neg; i++) {
        int64_t v = src[i];
        res[mp_1n + i] = neg_first ? v : (int64_t)(0 - (uint64_t)v);
    }
}

/* reference/vec_znx/rotate.rs:9-38 */
void orc_vec_znx_rotate(int64_t p, orc_vec_znx *res, size_t res_col, const orc_vec_znx *a, size_t a_col) {
    size_t mn = res->size < a->size ? res->size : a->size;
    for (size_t j = 0; j < mn; j++) orc_znx_rotate(p, znx_at(res, res_col, j), znx_at(a, a_col, j), res->n);
    for (size_t j = mn; j < res->size; j++) memset(znx_at(res, res_col, j), 0, 8 * res->n);
}

/* reference/vec_znx/add.rs:60-82 */
void orc_vec_znx_add_assign(orc_vec_znx *res, size_t res_col, const orc_vec_znx *a, size_t a_col) {
    size_t mn = res->size < a->size ? res->size : a->size;
    for (size_t j = 0; j < mn; j++) {
        int64_t *r = znx_at(res, res_col, j);
        const int64_t *x = znx_at(a, a_col, j);
        for (size_t i = 0; i < res->n; i++) r[i] = (int64_t)((uint64_t)r[i] + (uint64_t)x[i]);
    }
}

/* reference/vec_znx/mul_xp_minus_one.rs:24-38: per limb tmp = rotate(p, x); x = tmp - x (znx/sub.rs:26-36) */
void orc_vec_znx_mul_xp_minus_one_assign(int64_t p, orc_vec_znx *res, size_t res_col) {
    int64_t *tmp = (int64_t *)malloc(8 * res->n);
    for (size_t j = 0; j < res->size; j++) {
        int64_t *x = znx_at(res, res_col, j);
        orc_znx_rotate(p, tmp, x, res->n);
        for (size_t i = 0; i < res->n; i++) x[i] = (int64_t)((uint64_t)tmp[i] - (uint64_t)x[i]);
    }
    free(tmp);
}

/* reference/znx/normalization.rs:4-11 */
static inline int64_t get_digit64(size_t k, int64_t x) { return (int64_t)((uint64_t)x << (64 - k)) >> (64 - k); }
static inline int64_t get_carry64(size_t k, int64_t x, int64_t d) { return (int64_t)((uint64_t)x - (uint64_t)d) >> k; }

/* reference/vec_znx/normalize.rs:403-425 with the lsh = 0 steps of reference/znx/normalization.rs:44-56, :132-146, :254-264 */
void orc_vec_znx_normalize_assign(size_t base2k, orc_vec_znx *res, size_t res_col) {
    size_t n = res->n, size = res->size;
    int64_t *carry = (int64_t *)calloc(n, 8);
    for (size_t jj = 0; jj < size; jj++) {
        size_t j = size - 1 - jj;
        int64_t *x = znx_at(res, res_col, j);
        if (j == size - 1) {
            for (size_t i = 0; i < n; i++) {
                int64_t d = get_digit64(base2k, x[i]);
                carry[i] = get_carry64(base2k, x[i], d);
                x[i] = d;
            }
        } else if (j == 0) {
            for (size_t i = 0; i < n; i++) x[i] = get_digit64(base2k, (int64_t)((uint64_t)get_digit64(base2k, x[i]) + (uint64_t)carry[i]));
        } else {
            for (size_t i = 0; i < n; i++) {
                int64_t d = get_digit64(base2k, x[i]);
                int64_t c = get_carry64(base2k, x[i], d);
                int64_t dc = (int64_t)((uint64_t)d + (uint64_t)carry[i]);
                x[i] = get_digit64(base2k, dc);
                carry[i] = (int64_t)((uint64_t)c + (uint64_t)get_carry64(base2k, dc, x[i]));
            }
        }
    }
    free(carry);
}

/* reference/znx/automorphism.rs:1-17: a(X) -> a(X^p); coefficient i lands on i*p mod 2n, negated in the upper half */
void orc_znx_automorphism(int64_t p, int64_t *res, const int64_t *a, size_t n) {
    size_t k = 0, mask = 2 * n - 1, p_2n = (size_t)(p & (int64_t)mask);
    res[0] = a[0];
    for (size_t i = 1; i < n; i++) {
        k = (k + p_2n) & mask;
        if (k < n) res[k] = a[i];
        else res[k - n] = (int64_t)(0 - (uint64_t)a[i]);
    }
}
/* reference/vec_znx/automorphism.rs:9-38 */
void orc_vec_znx_automorphism(int64_t p, orc_vec_znx *res, size_t res_col, const orc_vec_znx *a, size_t a_col) {
    size_t mn = res->size < a->size ? res->size : a->size;
    for (size_t j = 0; j < mn; j++) orc_znx_automorphism(p, znx_at(res, res_col, j), znx_at(a, a_col, j), res->n);
    for (size_t j = mn; j < res->size; j++) memset(znx_at(res, res_col, j), 0, 8 * res->n);
}

/* reference/vec_znx/shift.rs:186-243 (vec_znx_rsh_assign), restated limb step for limb step with the i64 slice kernels of
 * reference/znx/normalization.rs (first/middle carry-only, middle_step_assign, final_step_assign); `carry` and `tmp` are the 2n words of
 * scratch the reference splits.  The third loop is kept verbatim (zero limb j, then step on limb steps-1-j). */
void orc_vec_znx_rsh_assign(size_t base2k, size_t k, orc_vec_znx *res, size_t res_col) {
    size_t n = res->n, size = res->size;
    size_t steps = k / base2k, k_rem = k % base2k;
    if (k_rem != 0) steps += 1;
    size_t lsh = (base2k - k_rem) % base2k;
    if (steps > size) abort(); /* the reference indexes limb size - j - 1: it panics here */
    int64_t *carry = (int64_t *)calloc(2 * n, sizeof(int64_t)), *tmp = carry + n;
    for (size_t j = 0; j < steps; j++) {
        if (j == 0) nfc_first_carry_only_i64(n, base2k, lsh, znx_at(res, res_col, size - j - 1), carry);
        else nfc_middle_carry_only_i64(n, base2k, lsh, znx_at(res, res_col, size - j - 1), carry);
    }
    for (size_t j = 0; j + steps < size; j++) {
        memcpy(tmp, znx_at(res, res_col, size - steps - j - 1), 8 * n);
        nfc_middle_step_assign_i64(n, base2k, lsh, tmp, carry);
        memcpy(znx_at(res, res_col, size - j - 1), tmp, 8 * n);
    }
    for (size_t j = 0; j < steps; j++) {
        memset(znx_at(res, res_col, j), 0, 8 * n);
        if (j == 0) nfc_final_step_assign_i64(n, base2k, lsh, znx_at(res, res_col, steps - j - 1), carry);
        else nfc_middle_step_assign_i64(n, base2k, lsh, znx_at(res, res_col, steps - j - 1), carry);
    }
    free(carry);
}
