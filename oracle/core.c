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
