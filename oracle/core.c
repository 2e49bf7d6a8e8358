/*
 * core.c -- the poulpy-core compositions that sit directly on the HAL hot path,
 * restated over the oracle's HAL functions.  TEST INFRASTRUCTURE ONLY.
 *
 * Restates: poulpy-core/src/keyswitching/glwe.rs:53-109 (glwe_keyswitch_default),
 *   :207-239 (glwe_keyswitch_internal), :298-380 (gglwe_product_dft, dsize == 1 and > 1),
 *   poulpy-core/src/external_product/glwe.rs:99-141, :197-271,
 *   poulpy-core/src/operations/glwe.rs:1286-1310 (glwe_normalize).
 */
#include "poulpy_oracle.h"

#include <assert.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int flavour;
    const void *mod;
    size_t prep_bytes; /* sizeof(ScalarPrep): 32 (Q120bScalar) or 8 (f64) */
} be_t;

static void be_dft_apply(const be_t *b, size_t step, size_t off, orc_vec_znx_dft *r, size_t rc, const orc_vec_znx *a, size_t ac) {
    if (b->flavour == 0) orc_ntt120_vec_znx_dft_apply((const orc_ntt120_module *)b->mod, step, off, r, rc, a, ac);
    else orc_fft64_vec_znx_dft_apply((const orc_fft64_module *)b->mod, step, off, r, rc, a, ac);
}
static void be_vmp(const be_t *b, orc_vec_znx_dft *r, const orc_vec_znx_dft *a, const orc_vmp_pmat *p, size_t lo) {
    if (b->flavour == 0) orc_ntt120_vmp_apply_dft_to_dft((const orc_ntt120_module *)b->mod, r, a, p, lo);
    else orc_fft64_vmp_apply_dft_to_dft((const orc_fft64_module *)b->mod, r, a, p, lo);
}
static void be_dft_copy(const be_t *b, size_t step, size_t off, orc_vec_znx_dft *r, size_t rc, const orc_vec_znx_dft *a, size_t ac) {
    if (b->flavour == 0) orc_ntt120_vec_znx_dft_copy(step, off, r, rc, a, ac);
    else orc_fft64_vec_znx_dft_copy(step, off, r, rc, a, ac);
}
static void be_dft_add_assign(const be_t *b, orc_vec_znx_dft *r, size_t rc, const orc_vec_znx_dft *a, size_t ac) {
    if (b->flavour == 0) orc_ntt120_vec_znx_dft_add_assign(r, rc, a, ac);
    else orc_fft64_vec_znx_dft_add_assign(r, rc, a, ac);
}
static void be_idft_consume(const be_t *b, orc_vec_znx_dft *a) {
    if (b->flavour == 0) orc_ntt120_vec_znx_idft_apply_consume((const orc_ntt120_module *)b->mod, a);
    else orc_fft64_vec_znx_idft_apply_consume((const orc_fft64_module *)b->mod, a);
}
static void be_big_add_small_assign(const be_t *b, orc_vec_znx_big *r, size_t rc, const orc_vec_znx *a, size_t ac) {
    if (b->flavour == 0) orc_ntt120_vec_znx_big_add_small_assign(r, rc, a, ac);
    else orc_fft64_vec_znx_big_add_small_assign(r, rc, a, ac);
}
static void be_big_normalize(const be_t *b, orc_vec_znx *r, size_t rk, int64_t off, size_t rc, const orc_vec_znx_big *a, size_t ak, size_t ac) {
    if (b->flavour == 0) orc_ntt120_vec_znx_big_normalize(r, rk, off, rc, a, ak, ac, 0);
    else orc_fft64_vec_znx_big_normalize(r, rk, off, rc, a, ak, ac, 0);
}
static orc_vec_znx_dft dft_alloc(const be_t *b, size_t n, size_t cols, size_t size) {
    orc_vec_znx_dft v = {calloc(n * cols * (size ? size : 1), b->prep_bytes), n, cols, size};
    return v;
}
static size_t zmin(size_t a, size_t b) { return a < b ? a : b; }
static size_t div_ceil(size_t a, size_t b) { return (a + b - 1) / b; }

/* keyswitching/glwe.rs:298-380 */
static void gglwe_product_dft(const be_t *b, orc_vec_znx_dft *res, const orc_vec_znx_dft *a, const orc_vmp_pmat *pmat,
                              size_t dsize) {
    if (dsize == 1) {
        be_vmp(b, res, a, pmat, 0);
        return;
    }
    size_t n = res->n, cols = a->cols, a_size = a->size, dnum = pmat->rows, cols_out = res->cols;
    size_t res_max = res->size;
    size_t ai_max = zmin(div_ceil(a_size, dsize), dnum);
    orc_vec_znx_dft ai = dft_alloc(b, n, cols, ai_max);
    orc_vec_znx_dft tmp = dft_alloc(b, n, cols_out, pmat->size);
    for (size_t di = 0; di < dsize; di++) {
        ai.size = zmin((a_size + di) / dsize, dnum);
        long cut = (long)(dsize - di) - 2;
        res->size = pmat->size - (size_t)(cut > 0 ? cut : 0);
        for (size_t j = 0; j < cols; j++) be_dft_copy(b, dsize, dsize - di - 1, &ai, j, a, j);
        if (di == 0) {
            be_vmp(b, res, &ai, pmat, 0);
        } else {
            tmp.size = res->size;
            be_vmp(b, &tmp, &ai, pmat, di);
            for (size_t c = 0; c < cols_out; c++) be_dft_add_assign(b, res, c, &tmp, c);
        }
    }
    res->size = res_max;
    free(ai.data);
    free(tmp.data);
}

/* operations/glwe.rs:1286-1310 over a freshly taken GLWE of size ceil(a.size*a_base2k / base2k) */
static orc_vec_znx conv_base2k(const orc_vec_znx *a, size_t a_base2k, size_t base2k) {
    size_t size = div_ceil(a->size * a_base2k, base2k);
    orc_vec_znx c = {(int64_t *)calloc(a->n * a->cols * size, 8), a->n, a->cols, size};
    for (size_t i = 0; i < a->cols; i++) orc_vec_znx_normalize(&c, base2k, 0, i, a, a_base2k, i, 0);
    return c;
}

static be_t make_be(int flavour, const void *mod) {
    be_t b = {flavour, mod, flavour == 0 ? 32u : 8u};
    return b;
}

/* keyswitching/glwe.rs:53-109 + :207-239 */
void orc_glwe_keyswitch(int flavour, const void *mod, orc_vec_znx *res, size_t res_base2k, const orc_vec_znx *a,
                        size_t a_base2k, const orc_vmp_pmat *key, size_t key_base2k, size_t dsize) {
    be_t b = make_be(flavour, mod);
    size_t n = res->n;
    assert(a->cols - 1 == key->cols_in && res->cols == key->cols_out);
    orc_vec_znx_dft res_dft = dft_alloc(&b, n, res->cols, key->size); /* zeroed */
    orc_vec_znx a_conv = {0};
    const orc_vec_znx *ain = a;
    if (a_base2k != key_base2k) {
        a_conv = conv_base2k(a, a_base2k, key_base2k);
        ain = &a_conv;
    }
    /* glwe_keyswitch_internal */
    size_t cols = ain->cols;
    orc_vec_znx_dft a_dft = dft_alloc(&b, n, cols - 1, ain->size);
    for (size_t c = 0; c + 1 < cols; c++) be_dft_apply(&b, 1, 0, &a_dft, c, ain, c + 1);
    gglwe_product_dft(&b, &res_dft, &a_dft, key, dsize);
    be_idft_consume(&b, &res_dft);
    orc_vec_znx_big res_big = {res_dft.data, n, res_dft.cols, res_dft.size};
    be_big_add_small_assign(&b, &res_big, 0, ain, 0);
    for (size_t i = 0; i < res->cols; i++) be_big_normalize(&b, res, res_base2k, 0, i, &res_big, key_base2k, i);
    free(a_dft.data);
    free(res_dft.data);
    free(a_conv.data);
}

/* external_product/glwe.rs:99-141 + :197-271 */
void orc_glwe_external_product(int flavour, const void *mod, orc_vec_znx *res, size_t res_base2k, const orc_vec_znx *a,
                               size_t a_base2k, const orc_vmp_pmat *ggsw, size_t ggsw_base2k, size_t dsize) {
    be_t b = make_be(flavour, mod);
    size_t n = res->n;
    assert(a->cols == ggsw->cols_in && res->cols == ggsw->cols_out);
    orc_vec_znx_dft res_dft = dft_alloc(&b, n, res->cols, ggsw->size);
    orc_vec_znx a_conv = {0};
    const orc_vec_znx *ain = a;
    if (a_base2k != ggsw_base2k) {
        a_conv = conv_base2k(a, a_base2k, ggsw_base2k);
        ain = &a_conv;
    }
    size_t cols = ggsw->cols_in, a_size = ain->size;
    orc_vec_znx_dft a_dft = dft_alloc(&b, n, cols, div_ceil(a_size, dsize));
    if (dsize == 1) {
        a_dft.size = a_size;
        res_dft.size = ggsw->size;
        for (size_t j = 0; j < cols; j++) be_dft_apply(&b, 1, 0, &a_dft, j, ain, j);
        be_vmp(&b, &res_dft, &a_dft, ggsw, 0);
    } else {
        orc_vec_znx_dft tmp = dft_alloc(&b, n, res_dft.cols, ggsw->size);
        for (size_t di = 0; di < dsize; di++) {
            a_dft.size = (a_size + di) / dsize;
            long cut = (long)(dsize - di) - 2;
            res_dft.size = ggsw->size - (size_t)(cut > 0 ? cut : 0);
            for (size_t j = 0; j < cols; j++) be_dft_apply(&b, dsize, dsize - 1 - di, &a_dft, j, ain, j);
            if (di == 0) {
                be_vmp(&b, &res_dft, &a_dft, ggsw, 0);
            } else {
                tmp.size = res_dft.size;
                be_vmp(&b, &tmp, &a_dft, ggsw, di);
                for (size_t c = 0; c < cols; c++) be_dft_add_assign(&b, &res_dft, c, &tmp, c);
            }
        }
        free(tmp.data);
    }
    be_idft_consume(&b, &res_dft);
    orc_vec_znx_big res_big = {res_dft.data, n, res_dft.cols, res_dft.size};
    for (size_t j = 0; j < res->cols; j++) be_big_normalize(&b, res, res_base2k, 0, j, &res_big, ggsw_base2k, j);
    free(a_dft.data);
    free(res_dft.data);
    free(a_conv.data);
}

/* ------------------------------------------------------------------ GLWE tensoring / relinearisation (CKKS multiplication) */
static void be_cnv_prepare(const be_t *b, orc_vec_znx_dft *r, const orc_vec_znx *a, int64_t mask) {
    if (b->flavour == 0) orc_ntt120_cnv_prepare((const orc_ntt120_module *)b->mod, r, a, mask);
    else orc_fft64_cnv_prepare((const orc_fft64_module *)b->mod, r, a, mask);
}
static void be_cnv_pairwise(const be_t *b, size_t off, orc_vec_znx_dft *r, const orc_vec_znx_dft *x, const orc_vec_znx_dft *y, size_t i, size_t j) {
    if (b->flavour == 0) orc_ntt120_cnv_pairwise_apply_dft((const orc_ntt120_module *)b->mod, off, r, 0, x, y, i, j);
    else orc_fft64_cnv_pairwise_apply_dft(off, r, 0, x, y, i, j);
}
/* operations/glwe.rs:921-926 */
static int64_t msb_mask_bottom_limb(size_t base2k, size_t k) {
    size_t r = k % base2k;
    return r == 0 ? ~(int64_t)0 : (int64_t)(~(uint64_t)0 << (base2k - r));
}
/* operations/glwe.rs:928-957 */
static size_t normalize_input_limb_bound_with_offset(size_t full, size_t res_size, size_t res_base2k, size_t in_base2k, int64_t res_offset) {
    int64_t ob = res_offset % (int64_t)in_base2k;
    if (res_offset < 0 && ob != 0) ob += (int64_t)in_base2k;
    return zmin(full, div_ceil(res_size * res_base2k + (size_t)ob, in_base2k));
}
static void znx_limbwise(orc_vec_znx *r, size_t rc, const orc_vec_znx *a, size_t ac, int op) { /* sizes are equal here */
    size_t mn = zmin(r->size, a->size);
    for (size_t j = 0; j < mn; j++) {
        int64_t *x = r->data + r->n * (j * r->cols + rc);
        const int64_t *y = a->data + a->n * (j * a->cols + ac);
        for (size_t i = 0; i < r->n; i++) {
            uint64_t u = (uint64_t)x[i], v = (uint64_t)y[i];
            x[i] = (int64_t)(op == 0 ? v : op == 1 ? u + v : op == 2 ? u - v : 0 - v); /* copy, add_assign, sub_assign, negate */
        }
    }
    if (op == 0 || op == 3)
        for (size_t j = mn; j < r->size; j++) memset(r->data + r->n * (j * r->cols + rc), 0, 8 * r->n);
}

/* operations/glwe.rs:699-818 (glwe_tensor_apply).  res = GLWETensor VecZnx with (rank+1)(rank+2)/2 columns. */
void orc_glwe_tensor_apply(int flavour, const void *mod, size_t cnv_offset, orc_vec_znx *res, size_t res_base2k,
                           const orc_vec_znx *a, size_t a_effective_k, const orc_vec_znx *bb, size_t b_effective_k, size_t ab_base2k) {
    be_t b = make_be(flavour, mod);
    size_t n = res->n, cols = a->cols;
    assert(bb->cols == cols && res->cols == cols * (cols + 1) / 2);
    assert(div_ceil(a_effective_k, ab_base2k) == a->size && div_ceil(b_effective_k, ab_base2k) == bb->size);
    orc_vec_znx_dft a_prep = dft_alloc(&b, n, cols, a->size), b_prep = dft_alloc(&b, n, cols, bb->size);
    be_cnv_prepare(&b, &a_prep, a, msb_mask_bottom_limb(ab_base2k, a_effective_k));
    be_cnv_prepare(&b, &b_prep, bb, msb_mask_bottom_limb(ab_base2k, b_effective_k));
    size_t off_hi;
    int64_t off_lo;
    if (cnv_offset < ab_base2k) {
        off_hi = 0;
        off_lo = -(int64_t)(ab_base2k - (cnv_offset % ab_base2k));
    } else {
        off_hi = cnv_offset / ab_base2k - 1; /* saturating_sub(1): cnv_offset >= ab_base2k here */
        off_lo = (int64_t)(cnv_offset % ab_base2k);
    }
    size_t dft_size = normalize_input_limb_bound_with_offset(a->size + bb->size - off_hi, res->size, res_base2k, ab_base2k, off_lo);
    orc_vec_znx tmp = {(int64_t *)calloc(n * res->size, 8), n, 1, res->size};
    for (size_t i = 0; i < cols; i++) {
        size_t col_i = i * cols - (i * (i + 1) / 2);
        orc_vec_znx_dft res_dft = dft_alloc(&b, n, 1, dft_size);
        be_cnv_pairwise(&b, off_hi, &res_dft, &a_prep, &b_prep, i, i); /* cnv_apply_dft(a_prep col i, b_prep col i) */
        be_idft_consume(&b, &res_dft);
        orc_vec_znx_big res_big = {res_dft.data, n, 1, res_dft.size};
        be_big_normalize(&b, &tmp, res_base2k, off_lo, 0, &res_big, ab_base2k, 0);
        znx_limbwise(res, col_i + i, &tmp, 0, 0);
        for (size_t j = 0; j < cols; j++) {
            if (j == i) continue;
            if (j < i) {
                size_t col_j = j * cols - (j * (j + 1) / 2);
                znx_limbwise(res, col_j + i, &tmp, 0, 2);
            } else {
                znx_limbwise(res, col_i + j, &tmp, 0, 3);
            }
        }
        free(res_dft.data);
    }
    for (size_t i = 0; i < cols; i++) {
        size_t col_i = i * cols - (i * (i + 1) / 2);
        for (size_t j = i + 1; j < cols; j++) {
            orc_vec_znx_dft res_dft = dft_alloc(&b, n, 1, dft_size);
            be_cnv_pairwise(&b, off_hi, &res_dft, &a_prep, &b_prep, i, j);
            be_idft_consume(&b, &res_dft);
            orc_vec_znx_big res_big = {res_dft.data, n, 1, res_dft.size};
            be_big_normalize(&b, &tmp, res_base2k, off_lo, 0, &res_big, ab_base2k, 0);
            znx_limbwise(res, col_i + j, &tmp, 0, 1);
            free(res_dft.data);
        }
    }
    free(tmp.data);
    free(a_prep.data);
    free(b_prep.data);
}

/* operations/glwe.rs:545-610 (glwe_tensor_relinearize): a = GLWETensor (base2k a_base2k), tsk = prepared GGLWE with
 * rank_in = rank(rank+1)/2 rows-columns and rank_out + 1 output columns; `tsk_size` = tsk.size() */
void orc_glwe_tensor_relinearize(int flavour, const void *mod, orc_vec_znx *res, size_t res_base2k, const orc_vec_znx *a,
                                 size_t a_base2k, const orc_vmp_pmat *tsk, size_t key_base2k, size_t dsize) {
    be_t b = make_be(flavour, mod);
    size_t n = res->n, cols = tsk->cols_out, pairs = tsk->cols_in;
    assert(res->cols == cols && a->cols == cols + pairs);
    size_t a_dft_size = div_ceil(a->size * a_base2k, key_base2k);
    orc_vec_znx_dft a_dft = dft_alloc(&b, n, pairs, a_dft_size);
    orc_vec_znx a_conv = {(int64_t *)calloc(n * a_dft_size, 8), n, 1, a_dft_size};
    for (size_t i = 0; i < pairs; i++) {
        if (a_base2k != key_base2k) {
            orc_vec_znx_normalize(&a_conv, key_base2k, 0, 0, a, a_base2k, cols + i, 0);
            be_dft_apply(&b, 1, 0, &a_dft, i, &a_conv, 0);
        } else {
            be_dft_apply(&b, 1, 0, &a_dft, i, a, cols + i);
        }
    }
    orc_vec_znx_dft res_dft = dft_alloc(&b, n, cols, tsk->size);
    gglwe_product_dft(&b, &res_dft, &a_dft, tsk, dsize);
    be_idft_consume(&b, &res_dft);
    orc_vec_znx_big res_big = {res_dft.data, n, res_dft.cols, res_dft.size};
    for (size_t i = 0; i < cols; i++) {
        if (res_base2k == key_base2k) { /* sic: the reference tests res_base2k, not a_base2k (:595) */
            be_big_add_small_assign(&b, &res_big, i, a, i);
        } else {
            orc_vec_znx_normalize(&a_conv, key_base2k, 0, 0, a, a_base2k, i, 0);
            be_big_add_small_assign(&b, &res_big, i, &a_conv, 0);
        }
    }
    for (size_t i = 0; i < res->cols; i++) be_big_normalize(&b, res, res_base2k, 0, i, &res_big, key_base2k, i);
    free(a_conv.data);
    free(a_dft.data);
    free(res_dft.data);
}

/* poulpy-core/src/automorphism/glwe_ct.rs:51-72 (glwe_automorphism_default): key-switch with the automorphism key, then
 * vec_znx_automorphism_assign(key.p()) on every column */
void orc_glwe_automorphism(int flavour, const void *mod, orc_vec_znx *res, size_t res_base2k, const orc_vec_znx *a, size_t a_base2k,
                           const orc_vmp_pmat *key, size_t key_base2k, int64_t p, size_t dsize) {
    orc_glwe_keyswitch(flavour, mod, res, res_base2k, a, a_base2k, key, key_base2k, dsize);
    orc_vec_znx tmp = {(int64_t *)malloc(8 * res->n * res->cols * res->size), res->n, res->cols, res->size};
    memcpy(tmp.data, res->data, 8 * res->n * res->cols * res->size);
    for (size_t i = 0; i < res->cols; i++) orc_vec_znx_automorphism(p, res, i, &tmp, i);
    free(tmp.data);
}

/* ---- automorphism_add_assign and trace --------------------------------------------------------------------------------- */
/* ntt120/vec_znx_big.rs:1499-1529 (i128) and fft64/vec_znx_big.rs:172-188 (i64, through reference/vec_znx/automorphism.rs):
 * X -> X^p in place on every limb of column `col` */
static void be_big_automorphism_assign(const be_t *b, int64_t p, orc_vec_znx_big *r, size_t col) {
    size_t n = r->n, eb = b->flavour == 0 ? 16 : 8, mask = 2 * n - 1, p_2n = (size_t)(p & (int64_t)mask);
    char *tmp = (char *)malloc(n * eb);
    for (size_t limb = 0; limb < r->size; limb++) {
        char *rj = (char *)r->data + eb * n * (limb * r->cols + col);
        memcpy(tmp, rj, n * eb);
        size_t k = 0;
        for (size_t i = 1; i < n; i++) {
            k = (k + p_2n) & mask;
            if (b->flavour == 0) {
                unsigned __int128 v;
                memcpy(&v, tmp + 16 * i, 16);
                if (k >= n) v = (unsigned __int128)0 - v; /* wrapping_neg */
                memcpy(rj + 16 * (k < n ? k : k - n), &v, 16);
            } else {
                uint64_t v;
                memcpy(&v, tmp + 8 * i, 8);
                if (k >= n) v = 0 - v;
                memcpy(rj + 8 * (k < n ? k : k - n), &v, 8);
            }
        }
    }
    free(tmp);
}

/* vec_znx_big_sub_small_assign (op 1: res -= a) / _sub_small_negate_assign (op 2: res = a - res, limbs of res beyond a.size negated):
 * ntt120/vec_znx_big.rs:1285-1318 and the FFT64 twins (i64), wrapping */
static void be_big_small_op(const be_t *b, int op, orc_vec_znx_big *r, size_t rc, const orc_vec_znx *a, size_t ac) {
    if (op == 0) {
        be_big_add_small_assign(b, r, rc, a, ac);
        return;
    }
    size_t n = r->n, mn = zmin(r->size, a->size);
    for (size_t j = 0; j < r->size; j++) {
        if (j >= mn && !(op == 2)) break;
        const int64_t *aj = j < a->size ? a->data + n * (j * a->cols + ac) : NULL;
        for (size_t i = 0; i < n; i++) {
            if (b->flavour == 0) {
                unsigned __int128 *rp = (unsigned __int128 *)r->data + n * (j * r->cols + rc) + i, x = aj ? (unsigned __int128)(__int128)aj[i] : 0;
                if (op == 1) *rp = *rp - x;
                else *rp = x - *rp;
            } else {
                uint64_t *rp = (uint64_t *)r->data + n * (j * r->cols + rc) + i, x = aj ? (uint64_t)aj[i] : 0;
                if (op == 1) *rp = *rp - x;
                else *rp = x - *rp;
            }
        }
    }
}

/* poulpy-core/src/automorphism/glwe_ct.rs:95-275 (glwe_automorphism_add / _sub / _sub_negate and their _assign forms):
 *   res_big = glwe_keyswitch_internal(res_dft(rank+1, key.size), a, key); per column: big_automorphism_assign(p), then
 *   big_add_small_assign (op 0) / big_sub_small_assign (op 1) / big_sub_small_negate_assign (op 2) with the column of a, big_normalize into
 *   res.  a and res share res_base2k; res may alias a. */
void orc_glwe_automorphism_op(int flavour, const void *mod, int op, orc_vec_znx *res, size_t res_base2k, const orc_vec_znx *a,
                              const orc_vmp_pmat *key, size_t key_base2k, int64_t p, size_t dsize) {
    be_t b = make_be(flavour, mod);
    size_t n = res->n;
    assert(res->cols - 1 == key->cols_in && res->cols == key->cols_out && a->cols == res->cols);
    orc_vec_znx_dft res_dft = dft_alloc(&b, n, res->cols, key->size);
    orc_vec_znx a_conv = {0};
    const orc_vec_znx *ain = a;
    if (res_base2k != key_base2k) {
        a_conv = conv_base2k(a, res_base2k, key_base2k);
        ain = &a_conv;
    }
    /* glwe_keyswitch_internal (keyswitching/glwe.rs:207-239) */
    orc_vec_znx_dft a_dft = dft_alloc(&b, n, ain->cols - 1, ain->size);
    for (size_t c = 0; c + 1 < ain->cols; c++) be_dft_apply(&b, 1, 0, &a_dft, c, ain, c + 1);
    gglwe_product_dft(&b, &res_dft, &a_dft, key, dsize);
    be_idft_consume(&b, &res_dft);
    orc_vec_znx_big res_big = {res_dft.data, n, res_dft.cols, res_dft.size};
    be_big_add_small_assign(&b, &res_big, 0, ain, 0);
    for (size_t i = 0; i < res->cols; i++) {
        be_big_automorphism_assign(&b, p, &res_big, i);
        be_big_small_op(&b, op, &res_big, i, ain, i);
        be_big_normalize(&b, res, res_base2k, 0, i, &res_big, key_base2k, i);
    }
    free(a_dft.data);
    free(res_dft.data);
    free(a_conv.data);
}
void orc_glwe_automorphism_add_assign(int flavour, const void *mod, orc_vec_znx *res, size_t res_base2k, const orc_vmp_pmat *key,
                                      size_t key_base2k, int64_t p, size_t dsize) {
    orc_glwe_automorphism_op(flavour, mod, 0, res, res_base2k, res, key, key_base2k, p, dsize);
}

/* GALOISGENERATOR = 5 (poulpy-hal/src/lib.rs); galois_element (layouts/module.rs:214-226) for a positive generator exponent */
static int64_t galois_element_pos(uint64_t e, size_t cyclotomic_order) {
    uint64_t r = 1, x = 5;
    while (e) { /* mod_exp_u64: wrapping square-and-multiply */
        if (e & 1) r *= x;
        x *= x;
        e >>= 1;
    }
    return (int64_t)(r & (uint64_t)(cyclotomic_order - 1));
}
/* poulpy-core/src/glwe_trace.rs:34-44 */
int64_t orc_trace_galois_element(size_t i, size_t n) { return i == 0 ? -1 : galois_element_pos((uint64_t)1 << (i - 1), 2 * n); }

/* poulpy-core/src/glwe_trace.rs:129-175 (glwe_trace_assign_default); keys[i] is the prepared automorphism key of
 * orc_trace_galois_element(i, n), i in [skip, log_n) (entries below `skip` are not read) */
void orc_glwe_trace_assign(int flavour, const void *mod, orc_vec_znx *res, size_t res_base2k, size_t skip, const orc_vmp_pmat *const *keys,
                           size_t key_base2k, size_t dsize) {
    size_t n = res->n, log_n = 0;
    while (((size_t)1 << log_n) < n) log_n++;
    assert(skip <= log_n);
    if (res_base2k != key_base2k) {
        orc_vec_znx res_conv = conv_base2k(res, res_base2k, key_base2k);
        orc_glwe_trace_assign(flavour, mod, &res_conv, key_base2k, skip, keys, key_base2k, dsize);
        for (size_t i = 0; i < res->cols; i++) orc_vec_znx_normalize(res, res_base2k, 0, i, &res_conv, key_base2k, i, 0); /* glwe_normalize */
        free(res_conv.data);
        return;
    }
    for (size_t i = skip; i < log_n; i++) {
        for (size_t c = 0; c < res->cols; c++) orc_vec_znx_rsh_assign(res_base2k, 1, res, c); /* glwe_rsh(1) (operations/glwe.rs:1096-1112) */
        orc_glwe_automorphism_add_assign(flavour, mod, res, res_base2k, keys[i], key_base2k, orc_trace_galois_element(i, n), dsize);
    }
}

/* ---- ggsw_expand_row (poulpy-core/src/conversion/gglwe_to_ggsw.rs:116-268; SURVEY 8f N4: the last step of circuit bootstrapping) ---------- */
/* ggsw: MatZnx(dnum, rank+1, rank+1, size) whose column-0 GLWEs are filled; writes the GLWEs of columns 1..rank.  tsk[c] is the prepared
 * GGLWE of s[c] * s (VmpPMat(dnum_tsk, rank, rank+1, size_tsk)), c < rank. */
void orc_ggsw_expand_row(int flavour, const void *mod, int64_t *ggsw, size_t n, size_t dnum, size_t rank, size_t size, size_t res_base2k,
                         const orc_vmp_pmat *const *tsk, size_t tsk_base2k, size_t dsize) {
    be_t b = make_be(flavour, mod);
    size_t cols = rank + 1, glwe_words = size * cols * n;
    size_t conv = div_ceil(size * res_base2k, tsk_base2k); /* res.max_k().div_ceil(tsk_base2k) */
    for (size_t row = 0; row < dnum; row++) {
        orc_vec_znx mi = {ggsw + (row * cols + 0) * glwe_words, n, cols, size};
        orc_vec_znx_dft a_dft = dft_alloc(&b, n, cols - 1, conv);
        orc_vec_znx a_0 = {(int64_t *)calloc(n * conv, 8), n, 1, conv};
        if (res_base2k == tsk_base2k) {
            for (size_t c = 0; c + 1 < cols; c++) be_dft_apply(&b, 1, 0, &a_dft, c, &mi, c + 1);
            for (size_t j = 0; j < size; j++) memcpy(a_0.data + j * n, mi.data + n * (j * cols + 0), 8 * n); /* vec_znx_copy */
        } else {
            for (size_t c = 0; c + 1 < cols; c++) {
                orc_vec_znx_normalize(&a_0, tsk_base2k, 0, 0, &mi, res_base2k, c + 1, 0);
                be_dft_apply(&b, 1, 0, &a_dft, c, &a_0, 0);
            }
            orc_vec_znx_normalize(&a_0, tsk_base2k, 0, 0, &mi, res_base2k, 0, 0);
        }
        /* ggsw_expand_rows_internal (:182-268) */
        for (size_t col = 1; col < cols; col++) {
            const orc_vmp_pmat *key = tsk[col - 1];
            orc_vec_znx_dft res_dft = dft_alloc(&b, n, cols, key->size); /* zeroed */
            gglwe_product_dft(&b, &res_dft, &a_dft, key, dsize);
            be_idft_consume(&b, &res_dft);
            orc_vec_znx_big res_big = {res_dft.data, n, res_dft.cols, res_dft.size};
            be_big_add_small_assign(&b, &res_big, col, &a_0, 0);
            orc_vec_znx out = {ggsw + (row * cols + col) * glwe_words, n, cols, size};
            for (size_t j = 0; j < cols; j++) be_big_normalize(&b, &out, res_base2k, 0, j, &res_big, tsk_base2k, j);
            free(res_dft.data);
        }
        free(a_dft.data);
        free(a_0.data);
    }
}
