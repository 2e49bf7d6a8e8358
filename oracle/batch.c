/*
 * batch.c -- pthread drivers over independent ciphertexts, used only for the CPU baseline in bench.py
 * and for batched parity checks.  TEST INFRASTRUCTURE ONLY.  The reference itself is single-threaded per
 * call (poulpy-bench/src/bench_suite/core/keyswitch.rs:87-92); parallelism over ciphertexts is the caller's.
 */
#include "poulpy_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <unistd.h>

int orc_num_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/* minimal dynamic parallel-for over [0, count) with pthreads (no OpenMP runtime dependency) */
typedef struct {
    void (*fn)(size_t i, void *ctx);
    void *ctx;
    size_t count;
    size_t next;
    pthread_mutex_t mu;
} pf_t;
static void *pf_worker(void *arg) {
    pf_t *p = (pf_t *)arg;
    for (;;) {
        pthread_mutex_lock(&p->mu);
        size_t i = p->next++;
        pthread_mutex_unlock(&p->mu);
        if (i >= p->count) break;
        p->fn(i, p->ctx);
    }
    return NULL;
}
static void parallel_for(size_t count, int threads, void (*fn)(size_t, void *), void *ctx) {
    if (threads <= 0) threads = orc_num_threads();
    if ((size_t)threads > count) threads = (int)(count ? count : 1);
    pf_t p = {fn, ctx, count, 0, PTHREAD_MUTEX_INITIALIZER};
    if (threads == 1) {
        pf_worker(&p);
        return;
    }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)threads);
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, pf_worker, &p);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
}

typedef struct {
    int flavour, ext;
    const void *mod;
    int64_t *res;
    const int64_t *a;
    size_t res_size, res_base2k, a_size, a_base2k, n, cols_a, cols_r, key_base2k, dsize;
    const orc_vmp_pmat *key;
} job_t;
static void job_fn(size_t i, void *ctx) {
    job_t *j = (job_t *)ctx;
    orc_vec_znx av = {(int64_t *)j->a + j->n * j->cols_a * j->a_size * i, j->n, j->cols_a, j->a_size};
    orc_vec_znx rv = {j->res + j->n * j->cols_r * j->res_size * i, j->n, j->cols_r, j->res_size};
    if (j->ext)
        orc_glwe_external_product(j->flavour, j->mod, &rv, j->res_base2k, &av, j->a_base2k, j->key, j->key_base2k, j->dsize);
    else
        orc_glwe_keyswitch(j->flavour, j->mod, &rv, j->res_base2k, &av, j->a_base2k, j->key, j->key_base2k, j->dsize);
}

void orc_glwe_keyswitch_batch(int flavour, const void *mod, int64_t *res, size_t res_size, size_t res_base2k,
                              const int64_t *a, size_t a_size, size_t a_base2k, size_t n, size_t rank_in,
                              size_t rank_out, const orc_vmp_pmat *key, size_t key_base2k, size_t dsize,
                              size_t batch, int threads) {
    job_t j = {flavour, 0, mod, res, a, res_size, res_base2k, a_size, a_base2k, n, rank_in + 1, rank_out + 1,
               key_base2k, dsize, key};
    parallel_for(batch, threads, job_fn, &j);
}

void orc_glwe_external_product_batch(int flavour, const void *mod, int64_t *res, size_t res_size, size_t res_base2k,
                                     const int64_t *a, size_t a_size, size_t a_base2k, size_t n, size_t rank,
                                     const orc_vmp_pmat *ggsw, size_t ggsw_base2k, size_t dsize, size_t batch,
                                     int threads) {
    job_t j = {flavour, 1, mod, res, a, res_size, res_base2k, a_size, a_base2k, n, rank + 1, rank + 1,
               ggsw_base2k, dsize, ggsw};
    parallel_for(batch, threads, job_fn, &j);
}

/* CGGI block-binary blind rotation over a batch of mod-switched LWEs (bench.py's CGGI cpu_baseline): one call of
 * orc_cggi_blind_rotate_block_binary per ciphertext, ciphertexts spread over the host threads. */
typedef struct {
    int flavour;
    const void *mod;
    int64_t *res;
    const int64_t *lwe;
    size_t n, cols, res_size, n_lwe, block_size, base2k;
    const orc_vec_znx *lut;
    const orc_vmp_pmat *brk;
    const orc_svp_ppol *x_pow_a;
} cggi_job_t;
static void cggi_job_fn(size_t i, void *ctx) {
    cggi_job_t *j = (cggi_job_t *)ctx;
    orc_vec_znx rv = {j->res + j->n * j->cols * j->res_size * i, j->n, j->cols, j->res_size};
    orc_cggi_blind_rotate_block_binary(j->flavour, j->mod, &rv, j->lwe + (j->n_lwe + 1) * i, j->n_lwe, j->lut, j->brk, j->x_pow_a,
                                       j->block_size, j->base2k);
}
void orc_cggi_blind_rotate_block_binary_batch(int flavour, const void *mod, int64_t *res, size_t n, size_t cols, size_t res_size,
                                              const int64_t *lwe_2n, size_t n_lwe, const orc_vec_znx *lut, const orc_vmp_pmat *brk,
                                              const orc_svp_ppol *x_pow_a, size_t block_size, size_t base2k, size_t batch, int threads) {
    cggi_job_t j = {flavour, mod, res, lwe_2n, n, cols, res_size, n_lwe, block_size, base2k, lut, brk, x_pow_a};
    parallel_for(batch, threads, cggi_job_fn, &j);
}
