/*
 * ntt120.c -- CPU restatement of poulpy-cpu-ref's NTT120 backend arithmetic.
 * TEST INFRASTRUCTURE ONLY (see poulpy_oracle.h).  Scalar C, same lazy-u64
 * arithmetic and reduction schedule as the reference so that (a) the q120b
 * values can be compared modulo Q_k at every step and (b) the CPU baseline has
 * the reference's cost structure.
 *
 * Restates (paths under poulpy-cpu-ref/src/reference/ntt120/):
 *   primes.rs:80-90, types.rs:240, mod.rs:115-127, arithmetic.rs, ntt.rs,
 *   mat_vec.rs (bbc only), vec_znx_dft.rs, svp.rs, vmp.rs, vec_znx_big.rs
 * and the lazy add/sub leaves of poulpy-cpu-ref/src/ntt120/prim.rs:64-165.
 */
#include "poulpy_oracle.h"

#include <assert.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- primes.rs:80-90 (Primes30) ------------------------------------------ */
const uint32_t ORC_Q[4] = {
    (1u << 30) - 2u * (1u << 17) + 1u, (1u << 30) - 17u * (1u << 17) + 1u,
    (1u << 30) - 23u * (1u << 17) + 1u, (1u << 30) - 42u * (1u << 17) + 1u};
const uint32_t ORC_OMEGA[4] = {1070907127u, 315046632u, 309185662u, 846468380u};
const uint32_t ORC_CRT_CST[4] = {43599465u, 292938863u, 594011630u, 140177212u};
#define LOG_Q 30u

/* types.rs:240 */
static inline uint64_t q_shifted(int k) { return (uint64_t)ORC_Q[k] << 33; }

/* ---- mod.rs:115-127 ------------------------------------------------------- */
static uint64_t pow2_mod(uint64_t e, uint64_t q) {
    uint64_t result = 1, base = 2 % q;
    while (e > 0) {
        if (e & 1) result = (uint64_t)(((u128)result * base) % q);
        base = (uint64_t)(((u128)base * base) % q);
        e >>= 1;
    }
    return result;
}

/* ntt.rs:142-160 */
static uint32_t modq_pow(uint32_t x, int64_t n, uint32_t q) {
    int64_t qm1 = (int64_t)q - 1;
    int64_t np_ = ((n % qm1) + qm1) % qm1;
    uint64_t np = (uint64_t)np_, val = x, q64 = q, res = 1;
    while (np) {
        if (np & 1) res = (res * val) % q64;
        val = (val * val) % q64;
        np >>= 1;
    }
    return (uint32_t)res;
}

static uint64_t ceil_log2_u64(uint64_t x) {
    if (x <= 1) return 0;
    uint64_t fl = 63 - (uint64_t)__builtin_clzll(x);
    return ((x & (x - 1)) == 0) ? fl : fl + 1;
}

/* ---- tables: ntt.rs:60-133 ------------------------------------------------ */
typedef struct {
    uint64_t q2bs[4];
    uint64_t bs, half_bs, mask;
    int reduce;
} step_meta;

typedef struct {
    uint64_t cst[4];
    uint64_t mask, h;
} reduc_meta;

typedef struct {
    size_t n;
    size_t nlevels;
    step_meta *levels;
    uint64_t *powomega;
    reduc_meta reduc;
} ntt_table;

typedef struct {
    uint64_t h;
    uint64_t s2l[4], s2h[4];
} bbc_meta;

struct orc_ntt120_module {
    size_t n;
    ntt_table fwd, inv;
    bbc_meta bbc;
};

/* ntt.rs:175-209 */
static uint64_t fill_reduction_meta(uint64_t bs_start, reduc_meta *out) {
    uint64_t bs_after = UINT64_MAX, min_h = bs_start / 2;
    for (uint64_t h = bs_start / 2; h < bs_start; h++) {
        uint64_t t = 0;
        for (int k = 0; k < 4; k++) {
            uint64_t p = pow2_mod(h, ORC_Q[k]);
            uint64_t pbs = p <= 1 ? 0 : ceil_log2_u64(p);
            uint64_t t1 = bs_start - h + pbs;
            uint64_t t2 = 1 + (t1 > h ? t1 : h);
            if (t < t2) t = t2;
        }
        if (t < bs_after) {
            min_h = h;
            bs_after = t;
        }
    }
    out->mask = (1ull << min_h) - 1;
    out->h = min_h;
    for (int k = 0; k < 4; k++) out->cst[k] = pow2_mod(min_h, ORC_Q[k]);
    return bs_after;
}

/* ntt.rs:212-216 */
static inline uint64_t pack_omega(uint64_t t, uint64_t half_bs, uint64_t q) {
    uint64_t t1 = (uint64_t)(((u128)t << half_bs) % q);
    return (t1 << 32) | t;
}

static void omegas_for(size_t n, uint32_t om[4]) { /* ntt.rs:164-167 */
    for (int k = 0; k < 4; k++) om[k] = modq_pow(ORC_OMEGA[k], (int64_t)((1 << 16) / n), ORC_Q[k]);
}

/* ntt.rs:222-358 */
static void build_fwd(ntt_table *t, size_t n) {
    memset(t, 0, sizeof *t);
    t->n = n;
    uint32_t om[4];
    omegas_for(n, om);
    uint64_t bs = 64;
    uint64_t bs_after = fill_reduction_meta(bs, &t->reduc);
    t->powomega = (uint64_t *)calloc(4 * 2 * (n ? n : 1), sizeof(uint64_t));
    t->levels = (step_meta *)calloc(20, sizeof(step_meta));
    if (n == 1) return;
    size_t po = 0, nl = 0;
    { /* level 0: a[i] *= omega^i */
        uint64_t half = (bs + 1) / 2;
        step_meta *m = &t->levels[nl++];
        m->half_bs = half;
        m->mask = (1ull << half) - 1;
        bs = half + LOG_Q + 1;
        m->bs = bs;
        m->reduce = 0;
        uint64_t pw[4] = {1, 1, 1, 1};
        for (size_t i = 0; i < n; i++) {
            for (int k = 0; k < 4; k++) t->powomega[po + 4 * i + k] = pack_omega(pw[k], half, ORC_Q[k]);
            for (int k = 0; k < 4; k++) pw[k] = (pw[k] * om[k]) % ORC_Q[k];
        }
        po += 4 * n;
    }
    for (size_t nn = n; nn >= 2; nn /= 2) {
        size_t halfnn = nn / 2;
        int do_reduce = (bs == 64);
        if (do_reduce) bs = bs_after;
        step_meta *m = &t->levels[nl++];
        for (int k = 0; k < 4; k++) m->q2bs[k] = (uint64_t)ORC_Q[k] << (bs - LOG_Q);
        uint64_t new_bs;
        if (nn >= 4) {
            uint64_t bs1 = bs + 1;
            uint64_t half = (bs1 + 1) / 2;
            uint64_t bs2 = half + LOG_Q + 1;
            new_bs = bs1 > bs2 ? bs1 : bs2;
            assert(new_bs <= 64);
            m->half_bs = half;
            m->mask = (1ull << half) - 1;
        } else {
            new_bs = bs + 1;
            m->half_bs = 0;
            m->mask = 0;
        }
        m->bs = new_bs;
        m->reduce = do_reduce;
        bs = new_bs;
        if (halfnn > 1) {
            size_t mstep = n / halfnn;
            uint64_t om_m[4], pw[4];
            for (int k = 0; k < 4; k++) pw[k] = om_m[k] = modq_pow(om[k], (int64_t)mstep, ORC_Q[k]);
            for (size_t i = 0; i < halfnn - 1; i++) {
                for (int k = 0; k < 4; k++) t->powomega[po + 4 * i + k] = pack_omega(pw[k], m->half_bs, ORC_Q[k]);
                for (int k = 0; k < 4; k++) pw[k] = (pw[k] * om_m[k]) % ORC_Q[k];
            }
            po += 4 * (halfnn - 1);
        }
    }
    t->nlevels = nl;
}

/* ntt.rs:365-510 */
static void build_inv(ntt_table *t, size_t n) {
    memset(t, 0, sizeof *t);
    t->n = n;
    uint32_t om[4];
    omegas_for(n, om);
    uint64_t bs = 64;
    uint64_t bs_after = fill_reduction_meta(bs, &t->reduc);
    t->powomega = (uint64_t *)calloc(4 * 2 * (n ? n : 1), sizeof(uint64_t));
    t->levels = (step_meta *)calloc(20, sizeof(step_meta));
    if (n == 1) return;
    size_t po = 0, nl = 0;
    { /* nn = 2 */
        int do_reduce = (bs == 64);
        if (do_reduce) bs = bs_after;
        step_meta *m = &t->levels[nl++];
        for (int k = 0; k < 4; k++) m->q2bs[k] = (uint64_t)ORC_Q[k] << (bs - LOG_Q);
        m->bs = bs + 1;
        m->half_bs = 0;
        m->mask = 0;
        m->reduce = do_reduce;
        bs = bs + 1;
    }
    for (size_t nn = 4; nn <= n; nn *= 2) {
        size_t halfnn = nn / 2;
        int do_reduce = (bs == 64);
        if (do_reduce) bs = bs_after;
        uint64_t half = (bs + 1) / 2;
        uint64_t bs_mult = half + LOG_Q + 1;
        uint64_t new_bs = 1 + (bs > bs_mult ? bs : bs_mult);
        assert(new_bs <= 64);
        step_meta *m = &t->levels[nl++];
        for (int k = 0; k < 4; k++) m->q2bs[k] = (uint64_t)ORC_Q[k] << (bs_mult - LOG_Q);
        m->bs = new_bs;
        m->half_bs = half;
        m->mask = (1ull << half) - 1;
        m->reduce = do_reduce;
        bs = new_bs;
        size_t mstep = n / halfnn;
        uint64_t om_m[4], pw[4];
        for (int k = 0; k < 4; k++) pw[k] = om_m[k] = modq_pow(om[k], -(int64_t)mstep, ORC_Q[k]);
        for (size_t i = 0; i < halfnn - 1; i++) {
            for (int k = 0; k < 4; k++) t->powomega[po + 4 * i + k] = pack_omega(pw[k], half, ORC_Q[k]);
            for (int k = 0; k < 4; k++) pw[k] = (pw[k] * om_m[k]) % ORC_Q[k];
        }
        po += 4 * (halfnn - 1);
    }
    { /* last: a[i] *= omega^-i * n^-1 */
        int do_reduce = (bs == 64);
        if (do_reduce) bs = bs_after;
        uint64_t half = (bs + 1) / 2;
        uint64_t new_bs = half + LOG_Q + 1;
        assert(new_bs <= 64);
        step_meta *m = &t->levels[nl++];
        for (int k = 0; k < 4; k++) m->q2bs[k] = (uint64_t)ORC_Q[k] << (new_bs - LOG_Q);
        m->bs = new_bs;
        m->half_bs = half;
        m->mask = (1ull << half) - 1;
        m->reduce = do_reduce;
        for (int k = 0; k < 4; k++) {
            uint64_t q = ORC_Q[k];
            uint64_t inv_n = modq_pow((uint32_t)n, -1, ORC_Q[k]);
            uint64_t om_inv = modq_pow(om[k], -1, ORC_Q[k]);
            uint64_t pw = inv_n;
            for (size_t i = 0; i < n; i++) {
                t->powomega[po + 4 * i + k] = pack_omega(pw, half, q);
                pw = (pw * om_inv) % q;
            }
        }
        po += 4 * n;
    }
    t->nlevels = nl;
}

/* mat_vec.rs:175-213 (BbcMeta::new) */
static double bit_size_red(uint64_t e) {
    double mx = 0.0;
    for (int k = 0; k < 4; k++) {
        uint64_t v = pow2_mod(e, ORC_Q[k]);
        if (v > 1) {
            double b = log2((double)v);
            if (b > mx) mx = b;
        }
    }
    return mx;
}
static void build_bbc(bbc_meta *m) {
    double ell_bs = log2(10000.0);
    double p32 = bit_size_red(32);
    double s1 = 32.0 + ell_bs;
    double best = 1e300;
    uint64_t min_h = 16;
    for (uint64_t h = 16; h < 32; h++) {
        double s2l = p32 + (double)h;
        double s2h = (s1 - (double)h) + bit_size_red(32 + h);
        double r = log2(pow(2.0, s1) + pow(2.0, s2l) + pow(2.0, s2h));
        if (r < best) {
            best = r;
            min_h = h;
        }
    }
    m->h = min_h;
    for (int k = 0; k < 4; k++) {
        m->s2l[k] = pow2_mod(32, ORC_Q[k]);
        m->s2h[k] = pow2_mod(32 + min_h, ORC_Q[k]);
    }
}

orc_ntt120_module *orc_ntt120_new(size_t n) {
    assert(n >= 1 && n <= (1u << 16) && (n & (n - 1)) == 0);
    orc_ntt120_module *m = (orc_ntt120_module *)calloc(1, sizeof *m);
    m->n = n;
    build_fwd(&m->fwd, n);
    build_inv(&m->inv, n);
    build_bbc(&m->bbc);
    return m;
}
void orc_ntt120_free(orc_ntt120_module *m) {
    if (!m) return;
    free(m->fwd.levels);
    free(m->fwd.powomega);
    free(m->inv.levels);
    free(m->inv.powomega);
    free(m);
}
size_t orc_ntt120_n(const orc_ntt120_module *m) { return m->n; }
static size_t dump_levels(const ntt_table *t, uint64_t *bs, int *reduce, size_t cap) {
    for (size_t i = 0; i < t->nlevels && i < cap; i++) {
        bs[i] = t->levels[i].bs;
        reduce[i] = t->levels[i].reduce;
    }
    return t->nlevels;
}
size_t orc_ntt120_fwd_levels(const orc_ntt120_module *m, uint64_t *bs, int *reduce, size_t cap) {
    return dump_levels(&m->fwd, bs, reduce, cap);
}
size_t orc_ntt120_inv_levels(const orc_ntt120_module *m, uint64_t *bs, int *reduce, size_t cap) {
    return dump_levels(&m->inv, bs, reduce, cap);
}
uint64_t orc_ntt120_reduc_h(const orc_ntt120_module *m) { return m->fwd.reduc.h; }
uint64_t orc_ntt120_bbc_h(const orc_ntt120_module *m) { return m->bbc.h; }

/* ---- arithmetic.rs -------------------------------------------------------- */
void orc_ntt120_b_from_znx64(size_t nn, uint64_t *res, const int64_t *x) { /* :39-60 */
    uint64_t oq[4];
    for (int k = 0; k < 4; k++) oq[k] = (uint64_t)ORC_Q[k] - ((1ull << 63) % ORC_Q[k]);
    const uint64_t mask_lo = 0x7FFFFFFFFFFFFFFFull;
    for (size_t j = 0; j < nn; j++) {
        uint64_t xu = (uint64_t)x[j];
        int neg = xu > mask_lo;
        uint64_t xl = xu & mask_lo;
        for (int k = 0; k < 4; k++) res[4 * j + k] = xl + (neg ? oq[k] : 0);
    }
}
void orc_ntt120_c_from_b(size_t nn, uint32_t *res, const uint64_t *x) { /* :202-214 */
    for (size_t j = 0; j < nn; j++)
        for (int k = 0; k < 4; k++) {
            uint64_t q = ORC_Q[k];
            uint64_t r = x[4 * j + k] % q;
            res[8 * j + 2 * k] = (uint32_t)r;
            res[8 * j + 2 * k + 1] = (uint32_t)((r << 32) % q);
        }
}
void orc_ntt120_b_to_znx128(size_t nn, i128 *res, const uint64_t *x) { /* :119-140 */
    i128 q[4], qm[4], crt[4];
    for (int k = 0; k < 4; k++) {
        q[k] = ORC_Q[k];
        crt[k] = ORC_CRT_CST[k];
    }
    i128 total = q[0] * q[1] * q[2] * q[3];
    qm[0] = q[1] * q[2] * q[3];
    qm[1] = q[0] * q[2] * q[3];
    qm[2] = q[0] * q[1] * q[3];
    qm[3] = q[0] * q[1] * q[2];
    i128 half = (total + 1) / 2;
    for (size_t j = 0; j < nn; j++) {
        i128 tmp = 0;
        for (int k = 0; k < 4; k++) {
            i128 xk = (i128)(x[4 * j + k] % (uint64_t)ORC_Q[k]);
            i128 t = (xk * crt[k]) % q[k];
            tmp += t * qm[k];
        }
        tmp %= total;
        res[j] = tmp >= half ? tmp - total : tmp;
    }
}

/* ---- ntt.rs:527-544 ------------------------------------------------------- */
static inline uint64_t split_precompmul(uint64_t inp, uint64_t po, uint64_t half_bs, uint64_t mask) {
    uint64_t lo = inp & mask;
    uint64_t t = po & 0xFFFFFFFFull;
    uint64_t t1 = po >> 32;
    return lo * t + (inp >> half_bs) * t1;
}
static inline uint64_t modq_red(uint64_t x, uint64_t h, uint64_t mask, uint64_t cst) {
    return (x & mask) + (x >> h) * cst;
}

/* ---- "cpu-avx-style" leaves ------------------------------------------------------------------------------------------------------
 * poulpy-cpu-avx keeps the four primes of one coefficient in one __m256i (4 x u64 lanes) and runs the SAME lazy arithmetic with
 * _mm256_mul_epu32 (poulpy-cpu-avx/src/ntt120/ntt.rs:81-110 for the butterflies, mat_vec_avx.rs for the bbc products).  The functions
 * below are that data path: lane k computes exactly what the scalar loop over k computes (same split products, same reduction schedule),
 * so every intermediate u64 is identical to the scalar port and the switch changes speed only (tests/test_oracle_kat.py checks it).
 * orc_ntt120_set_simd(1) selects them; 0 (default) = the scalar restatement of poulpy-cpu-ref. */
#include <immintrin.h>
static int g_simd = 0;
void orc_ntt120_set_simd(int on) { g_simd = on ? 1 : 0; }
int orc_ntt120_get_simd(void) { return g_simd; }

static inline __m256i v_split_precompmul(__m256i inp, __m256i po, __m128i half_bs, __m256i mask) {
    __m256i lo = _mm256_and_si256(inp, mask);
    __m256i hi = _mm256_srl_epi64(inp, half_bs);
    __m256i t1 = _mm256_srli_epi64(po, 32);
    /* mul_epu32 multiplies the low 32 bits of every 64-bit lane: lo < 2^half_bs <= 2^32, hi < 2^32, t = low word of po, t1 < 2^32 */
    return _mm256_add_epi64(_mm256_mul_epu32(lo, po), _mm256_mul_epu32(hi, t1));
}
static inline __m256i v_modq_red(__m256i x, __m128i h, __m256i mask, __m256i cst) {
    return _mm256_add_epi64(_mm256_and_si256(x, mask), _mm256_mul_epu32(_mm256_srl_epi64(x, h), cst));
}
static void fwd_block_avx(uint64_t *d, size_t blk, size_t halfnn, const step_meta *m, const reduc_meta *r, const uint64_t *po) {
    const __m256i q2bs = _mm256_loadu_si256((const __m256i *)m->q2bs), mask = _mm256_set1_epi64x((long long)m->mask);
    const __m256i rmask = _mm256_set1_epi64x((long long)r->mask), rcst = _mm256_loadu_si256((const __m256i *)r->cst);
    const __m128i hb = _mm_cvtsi64_si128((long long)m->half_bs), rh = _mm_cvtsi64_si128((long long)r->h);
    for (size_t i = 0; i < halfnn; i++) {
        __m256i *pa = (__m256i *)(d + 4 * (blk + i)), *pb = (__m256i *)(d + 4 * (blk + halfnn + i));
        __m256i a = _mm256_loadu_si256(pa), b = _mm256_loadu_si256(pb);
        if (m->reduce) {
            a = v_modq_red(a, rh, rmask, rcst);
            b = v_modq_red(b, rh, rmask, rcst);
        }
        _mm256_storeu_si256(pa, _mm256_add_epi64(a, b));
        __m256i b1 = _mm256_sub_epi64(_mm256_add_epi64(a, q2bs), b);
        if (i) b1 = v_split_precompmul(b1, _mm256_loadu_si256((const __m256i *)(po + 4 * (i - 1))), hb, mask);
        _mm256_storeu_si256(pb, b1);
    }
}
static void inv_block_avx(uint64_t *d, size_t blk, size_t halfnn, const step_meta *m, const reduc_meta *r, const uint64_t *po) {
    const __m256i q2bs = _mm256_loadu_si256((const __m256i *)m->q2bs), mask = _mm256_set1_epi64x((long long)m->mask);
    const __m256i rmask = _mm256_set1_epi64x((long long)r->mask), rcst = _mm256_loadu_si256((const __m256i *)r->cst);
    const __m128i hb = _mm_cvtsi64_si128((long long)m->half_bs), rh = _mm_cvtsi64_si128((long long)r->h);
    for (size_t i = 0; i < halfnn; i++) {
        __m256i *pa = (__m256i *)(d + 4 * (blk + i)), *pb = (__m256i *)(d + 4 * (blk + halfnn + i));
        __m256i a = _mm256_loadu_si256(pa), b = _mm256_loadu_si256(pb);
        if (m->reduce) {
            a = v_modq_red(a, rh, rmask, rcst);
            b = v_modq_red(b, rh, rmask, rcst);
        }
        if (i) b = v_split_precompmul(b, _mm256_loadu_si256((const __m256i *)(po + 4 * (i - 1))), hb, mask);
        _mm256_storeu_si256(pa, _mm256_add_epi64(a, b));
        _mm256_storeu_si256(pb, _mm256_sub_epi64(_mm256_add_epi64(a, q2bs), b));
    }
}
static void twist_pass_avx(uint64_t *d, size_t n, const uint64_t *po, const step_meta *m, const reduc_meta *r, int reduce) {
    const __m256i mask = _mm256_set1_epi64x((long long)m->mask);
    const __m256i rmask = _mm256_set1_epi64x((long long)r->mask), rcst = _mm256_loadu_si256((const __m256i *)r->cst);
    const __m128i hb = _mm_cvtsi64_si128((long long)m->half_bs), rh = _mm_cvtsi64_si128((long long)r->h);
    for (size_t i = 0; i < n; i++) {
        __m256i x = _mm256_loadu_si256((const __m256i *)(d + 4 * i));
        if (reduce) x = v_modq_red(x, rh, rmask, rcst);
        _mm256_storeu_si256((__m256i *)(d + 4 * i), v_split_precompmul(x, _mm256_loadu_si256((const __m256i *)(po + 4 * i)), hb, mask));
    }
}

/* ntt.rs:695-749 */
static void fwd_block(uint64_t *d, size_t blk, size_t halfnn, const step_meta *m, const reduc_meta *r,
                      const uint64_t *po) {
    for (size_t i = 0; i < halfnn; i++)
        for (int k = 0; k < 4; k++) {
            size_t ia = 4 * (blk + i) + k, ib = 4 * (blk + halfnn + i) + k;
            uint64_t a = d[ia], b = d[ib];
            if (m->reduce) {
                a = modq_red(a, r->h, r->mask, r->cst[k]);
                b = modq_red(b, r->h, r->mask, r->cst[k]);
            }
            d[ia] = a + b;
            uint64_t b1 = a + m->q2bs[k] - b;
            d[ib] = (i == 0) ? b1 : split_precompmul(b1, po[4 * (i - 1) + k], m->half_bs, m->mask);
        }
}
/* ntt.rs:757-811 */
static void inv_block(uint64_t *d, size_t blk, size_t halfnn, const step_meta *m, const reduc_meta *r,
                      const uint64_t *po) {
    for (size_t i = 0; i < halfnn; i++)
        for (int k = 0; k < 4; k++) {
            size_t ia = 4 * (blk + i) + k, ib = 4 * (blk + halfnn + i) + k;
            uint64_t a = d[ia], b = d[ib];
            if (m->reduce) {
                a = modq_red(a, r->h, r->mask, r->cst[k]);
                b = modq_red(b, r->h, r->mask, r->cst[k]);
            }
            uint64_t bo = (i == 0) ? b : split_precompmul(b, po[4 * (i - 1) + k], m->half_bs, m->mask);
            d[ia] = a + bo;
            d[ib] = a + m->q2bs[k] - bo;
        }
}

void orc_ntt120_ntt(const orc_ntt120_module *mod, uint64_t *d) { /* ntt.rs:558-602 */
    const ntt_table *t = &mod->fwd;
    size_t n = t->n;
    if (n == 1) return;
    size_t po = 0, mi = 0;
    {
        const step_meta *m = &t->levels[mi++];
        if (g_simd) twist_pass_avx(d, n, t->powomega + po, m, &t->reduc, 0);
        else
        for (size_t i = 0; i < n; i++)
            for (int k = 0; k < 4; k++)
                d[4 * i + k] = split_precompmul(d[4 * i + k], t->powomega[po + 4 * i + k], m->half_bs, m->mask);
        po += 4 * n;
    }
    for (size_t nn = n; nn >= 2; nn /= 2) {
        size_t halfnn = nn / 2;
        const step_meta *m = &t->levels[mi++];
        if (g_simd) for (size_t blk = 0; blk < n; blk += nn) fwd_block_avx(d, blk, halfnn, m, &t->reduc, t->powomega + po);
        else
        for (size_t blk = 0; blk < n; blk += nn) fwd_block(d, blk, halfnn, m, &t->reduc, t->powomega + po);
        po += 4 * (halfnn > 0 ? halfnn - 1 : 0);
    }
}

void orc_ntt120_intt(const orc_ntt120_module *mod, uint64_t *d) { /* ntt.rs:617-684 */
    const ntt_table *t = &mod->inv;
    size_t n = t->n;
    if (n == 1) return;
    size_t po = 0, mi = 0;
    {
        const step_meta *m = &t->levels[mi++];
        if (g_simd) for (size_t blk = 0; blk < n; blk += 2) inv_block_avx(d, blk, 1, m, &t->reduc, t->powomega + po);
        else
        for (size_t blk = 0; blk < n; blk += 2) inv_block(d, blk, 1, m, &t->reduc, t->powomega + po);
    }
    for (size_t nn = 4; nn <= n; nn *= 2) {
        size_t halfnn = nn / 2;
        const step_meta *m = &t->levels[mi++];
        if (g_simd) for (size_t blk = 0; blk < n; blk += nn) inv_block_avx(d, blk, halfnn, m, &t->reduc, t->powomega + po);
        else
        for (size_t blk = 0; blk < n; blk += nn) inv_block(d, blk, halfnn, m, &t->reduc, t->powomega + po);
        po += 4 * (halfnn - 1);
    }
    {
        const step_meta *m = &t->levels[mi];
        if (g_simd) twist_pass_avx(d, n, t->powomega + po, m, &t->reduc, m->reduce);
        else
        for (size_t i = 0; i < n; i++)
            for (int k = 0; k < 4; k++) {
                uint64_t x = d[4 * i + k];
                if (m->reduce) x = modq_red(x, t->reduc.h, t->reduc.mask, t->reduc.cst[k]);
                d[4 * i + k] = split_precompmul(x, t->powomega[po + 4 * i + k], m->half_bs, m->mask);
            }
    }
}

/* ---- mat_vec.rs:343-447 (bbc) --------------------------------------------- */
static inline void accum_mul_q120_bc(uint64_t s[8], const uint32_t *x, const uint32_t *y) {
    for (int i = 0; i < 4; i++) {
        uint64_t xl = x[2 * i], xh = x[2 * i + 1], yl = y[2 * i], yh = y[2 * i + 1];
        uint64_t lo = xl * yl, hi = xh * yh;
        s[2 * i] += (lo & 0xFFFFFFFFull) + (hi & 0xFFFFFFFFull);
        s[2 * i + 1] += (lo >> 32) + (hi >> 32);
    }
}
/* mat_vec_avx.rs: x = 4 lazy residues as u64 (low word xl, high word xh), y = 4 x (r, r 2^32 mod q) as u32 pairs; the sums of the low and
 * of the high product words are kept in two vectors (the scalar s[2i] / s[2i+1]) */
static inline void v_accum_mul_q120_bc(__m256i *s_lo, __m256i *s_hi, const uint32_t *x, const uint32_t *y) {
    const __m256i lomask = _mm256_set1_epi64x(0xFFFFFFFFll);
    __m256i vx = _mm256_loadu_si256((const __m256i *)x), vy = _mm256_loadu_si256((const __m256i *)y);
    __m256i lo = _mm256_mul_epu32(vx, vy), hi = _mm256_mul_epu32(_mm256_srli_epi64(vx, 32), _mm256_srli_epi64(vy, 32));
    *s_lo = _mm256_add_epi64(*s_lo, _mm256_add_epi64(_mm256_and_si256(lo, lomask), _mm256_and_si256(hi, lomask)));
    *s_hi = _mm256_add_epi64(*s_hi, _mm256_add_epi64(_mm256_srli_epi64(lo, 32), _mm256_srli_epi64(hi, 32)));
}
static inline void v_accum_to_q120b(uint64_t res[4], __m256i s_lo, __m256i s_hi, const bbc_meta *m) {
    const __m256i mask2 = _mm256_set1_epi64x((long long)((1ull << m->h) - 1));
    const __m128i h = _mm_cvtsi64_si128((long long)m->h);
    __m256i s2l = _mm256_and_si256(s_hi, mask2), s2h = _mm256_srl_epi64(s_hi, h);
    __m256i r = _mm256_add_epi64(s_lo, _mm256_add_epi64(_mm256_mul_epu32(s2l, _mm256_loadu_si256((const __m256i *)m->s2l)),
                                                       _mm256_mul_epu32(s2h, _mm256_loadu_si256((const __m256i *)m->s2h))));
    _mm256_storeu_si256((__m256i *)res, r);
}
static inline void accum_to_q120b(uint64_t res[4], const uint64_t s[8], const bbc_meta *m) {
    uint64_t mask2 = (1ull << m->h) - 1;
    for (int k = 0; k < 4; k++) {
        uint64_t s2l = s[2 * k + 1] & mask2, s2h = s[2 * k + 1] >> m->h;
        res[k] = s[2 * k] + s2l * m->s2l[k] + s2h * m->s2h[k];
    }
}
/* mat_vec.rs:374-387 */
static void mat1col_bbc(const bbc_meta *m, size_t ell, uint64_t *res, const uint32_t *x, const uint32_t *y) {
    uint64_t s[8] = {0};
    for (size_t i = 0; i < ell; i++) accum_mul_q120_bc(s, x + 8 * i, y + 8 * i);
    accum_to_q120b(res, s, m);
}
/* mat_vec.rs:391-418 */
static void mat1col_x2_bbc(const bbc_meta *m, size_t ell, uint64_t *res, const uint32_t *x, const uint32_t *y) {
    if (g_simd) {
        __m256i l0 = _mm256_setzero_si256(), h0 = l0, l1 = l0, h1 = l0;
        for (size_t i = 0; i < ell; i++) {
            v_accum_mul_q120_bc(&l0, &h0, x + 16 * i, y + 16 * i);
            v_accum_mul_q120_bc(&l1, &h1, x + 16 * i + 8, y + 16 * i + 8);
        }
        v_accum_to_q120b(res, l0, h0, m);
        v_accum_to_q120b(res + 4, l1, h1, m);
        return;
    }
    uint64_t s[2][8] = {{0}};
    for (size_t i = 0; i < ell; i++) {
        accum_mul_q120_bc(s[0], x + 16 * i, y + 16 * i);
        accum_mul_q120_bc(s[1], x + 16 * i + 8, y + 16 * i + 8);
    }
    accum_to_q120b(res, s[0], m);
    accum_to_q120b(res + 4, s[1], m);
}
/* mat_vec.rs:423-447 */
static void mat2cols_x2_bbc(const bbc_meta *m, size_t ell, uint64_t *res, const uint32_t *x, const uint32_t *y) {
    if (g_simd) {
        __m256i l[4], h[4];
        for (int o = 0; o < 4; o++) l[o] = h[o] = _mm256_setzero_si256();
        for (size_t i = 0; i < ell; i++) {
            const uint32_t *x0 = x + 16 * i, *x1 = x + 16 * i + 8;
            v_accum_mul_q120_bc(&l[0], &h[0], x0, y + 32 * i);
            v_accum_mul_q120_bc(&l[1], &h[1], x1, y + 32 * i + 8);
            v_accum_mul_q120_bc(&l[2], &h[2], x0, y + 32 * i + 16);
            v_accum_mul_q120_bc(&l[3], &h[3], x1, y + 32 * i + 24);
        }
        for (int o = 0; o < 4; o++) v_accum_to_q120b(res + 4 * o, l[o], h[o], m);
        return;
    }
    uint64_t s[4][8] = {{0}};
    for (size_t i = 0; i < ell; i++) {
        const uint32_t *x0 = x + 16 * i, *x1 = x + 16 * i + 8;
        accum_mul_q120_bc(s[0], x0, y + 32 * i);
        accum_mul_q120_bc(s[1], x1, y + 32 * i + 8);
        accum_mul_q120_bc(s[2], x0, y + 32 * i + 16);
        accum_mul_q120_bc(s[3], x1, y + 32 * i + 24);
    }
    for (int o = 0; o < 4; o++) accum_to_q120b(res + 4 * o, s[o], m);
}

/* ---- vec_znx_dft.rs ------------------------------------------------------- */
static inline uint64_t *dft_limb(const orc_vec_znx_dft *v, size_t col, size_t limb) {
    assert(col < v->cols && limb < v->size);
    return (uint64_t *)v->data + 4 * v->n * (limb * v->cols + col);
}
static inline i128 *big_limb(const orc_vec_znx_big *v, size_t col, size_t limb) {
    assert(col < v->cols && limb < v->size);
    return (i128 *)v->data + v->n * (limb * v->cols + col);
}
static inline int64_t *znx_limb(const orc_vec_znx *v, size_t col, size_t limb) {
    assert(col < v->cols && limb < v->size);
    return v->data + v->n * (limb * v->cols + col);
}
static inline size_t zmin(size_t a, size_t b) { return a < b ? a : b; }
static inline size_t div_ceil(size_t a, size_t b) { return (a + b - 1) / b; }

/* :177-215 */
void orc_ntt120_vec_znx_dft_apply(const orc_ntt120_module *m, size_t step, size_t offset, orc_vec_znx_dft *res,
                                  size_t res_col, const orc_vec_znx *a, size_t a_col) {
    size_t n = res->n;
    size_t steps = div_ceil(a->size, step);
    size_t min_steps = zmin(res->size, steps);
    for (size_t j = 0; j < min_steps; j++) {
        size_t limb = offset + j * step;
        uint64_t *r = dft_limb(res, res_col, j);
        if (limb < a->size) {
            orc_ntt120_b_from_znx64(n, r, znx_limb(a, a_col, limb));
            orc_ntt120_ntt(m, r);
        } else {
            memset(r, 0, 32 * n);
        }
    }
    for (size_t j = min_steps; j < res->size; j++) memset(dft_limb(res, res_col, j), 0, 32 * n);
}

/* :236-268 */
void orc_ntt120_vec_znx_idft_apply(const orc_ntt120_module *m, orc_vec_znx_big *res, size_t res_col,
                                   const orc_vec_znx_dft *a, size_t a_col) {
    size_t n = res->n;
    size_t min_size = zmin(res->size, a->size);
    uint64_t *tmp = (uint64_t *)malloc(32 * n);
    for (size_t j = 0; j < min_size; j++) {
        memcpy(tmp, dft_limb(a, a_col, j), 32 * n);
        orc_ntt120_intt(m, tmp);
        orc_ntt120_b_to_znx128(n, big_limb(res, res_col, j), tmp);
    }
    for (size_t j = min_size; j < res->size; j++) memset(big_limb(res, res_col, j), 0, 16 * n);
    free(tmp);
}
/* :274-303 */
void orc_ntt120_vec_znx_idft_apply_tmpa(const orc_ntt120_module *m, orc_vec_znx_big *res, size_t res_col,
                                        orc_vec_znx_dft *a, size_t a_col) {
    size_t n = res->n;
    size_t min_size = zmin(res->size, a->size);
    for (size_t j = 0; j < min_size; j++) {
        orc_ntt120_intt(m, dft_limb(a, a_col, j));
        orc_ntt120_b_to_znx128(n, big_limb(res, res_col, j), dft_limb(a, a_col, j));
    }
    for (size_t j = min_size; j < res->size; j++) memset(big_limb(res, res_col, j), 0, 16 * n);
}

/* :306-325 */
static inline uint64_t barrett_u61(uint64_t x, uint64_t q, uint64_t mu) {
    uint64_t qa = (uint64_t)(((u128)x * mu) >> 61);
    uint64_t r = x - qa * q;
    if (r >= q) r -= q;
    if (r >= q) r -= q;
    return r;
}
static inline uint64_t reduce_q120b_crt(uint64_t x, uint64_t q, uint64_t mu, uint64_t p32, uint64_t p16, uint64_t crt) {
    uint64_t xh = x >> 32;
    uint64_t xhr = xh >= q ? xh - q : xh;
    uint64_t xl = x & 0xFFFFFFFFull;
    uint64_t tmp = xhr * p32 + (xl >> 16) * p16 + (xl & 0xFFFF) * crt;
    return barrett_u61(tmp, q, mu);
}
/* :327-409 (compact_all_blocks_scalar + consume) */
void orc_ntt120_vec_znx_idft_apply_consume(const orc_ntt120_module *m, orc_vec_znx_dft *a) {
    size_t n = a->n, n_blocks = a->cols * a->size;
    uint64_t *p = (uint64_t *)a->data;
    uint64_t q64[4], mu[4], crt[4], p32[4], p16[4];
    u128 q[4], qm[4];
    for (int k = 0; k < 4; k++) {
        q64[k] = ORC_Q[k];
        mu[k] = (1ull << 61) / q64[k];
        crt[k] = ORC_CRT_CST[k];
        q[k] = q64[k];
    }
    for (int k = 0; k < 4; k++) {
        uint64_t pow32 = (uint64_t)(((u128)1 << 32) % q64[k]);
        p32[k] = barrett_u61(pow32 * crt[k], q64[k], mu[k]);
        p16[k] = barrett_u61((1ull << 16) * crt[k], q64[k], mu[k]);
    }
    u128 total = q[0] * q[1] * q[2] * q[3];
    qm[0] = q[1] * q[2] * q[3];
    qm[1] = q[0] * q[2] * q[3];
    qm[2] = q[0] * q[1] * q[3];
    qm[3] = q[0] * q[1] * q[2];
    u128 half_q = (total + 1) / 2;
    u128 tq[4] = {0, total, total * 2, total * 3};
    for (size_t b = 0; b < n_blocks; b++) {
        uint64_t *src = p + 4 * n * b;
        uint64_t *dst = p + 2 * n * b;
        orc_ntt120_intt(m, src);
        for (size_t c = 0; c < n; c++) {
            uint64_t t0 = reduce_q120b_crt(src[4 * c], q64[0], mu[0], p32[0], p16[0], crt[0]);
            uint64_t t1 = reduce_q120b_crt(src[4 * c + 1], q64[1], mu[1], p32[1], p16[1], crt[1]);
            uint64_t t2 = reduce_q120b_crt(src[4 * c + 2], q64[2], mu[2], p32[2], p16[2], crt[2]);
            uint64_t t3 = reduce_q120b_crt(src[4 * c + 3], q64[3], mu[3], p32[3], p16[3], crt[3]);
            u128 v = (u128)t0 * qm[0] + (u128)t1 * qm[1] + (u128)t2 * qm[2] + (u128)t3 * qm[3];
            size_t qa = (size_t)(v >> 120);
            v -= tq[qa];
            if (v >= total) v -= total;
            i128 val = v >= half_q ? (i128)v - (i128)total : (i128)v;
            memcpy(dst + 2 * c, &val, 16);
        }
    }
}

/* prim.rs:64-165 lazy leaves */
static void l_add(size_t n, uint64_t *r, const uint64_t *a, const uint64_t *b) {
    for (size_t j = 0; j < n; j++)
        for (int k = 0; k < 4; k++) r[4 * j + k] = a[4 * j + k] % q_shifted(k) + b[4 * j + k] % q_shifted(k);
}
static void l_sub(size_t n, uint64_t *r, const uint64_t *a, const uint64_t *b) {
    for (size_t j = 0; j < n; j++)
        for (int k = 0; k < 4; k++)
            r[4 * j + k] = a[4 * j + k] % q_shifted(k) + (q_shifted(k) - b[4 * j + k] % q_shifted(k));
}
static void l_neg(size_t n, uint64_t *r, const uint64_t *a) {
    for (size_t j = 0; j < n; j++)
        for (int k = 0; k < 4; k++) r[4 * j + k] = q_shifted(k) - a[4 * j + k] % q_shifted(k);
}

/* :418-470 */
void orc_ntt120_vec_znx_dft_add_into(orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx_dft *a, size_t a_col,
                                     const orc_vec_znx_dft *b, size_t b_col) {
    size_t n = res->n, rs = res->size;
    const orc_vec_znx_dft *lo = a->size <= b->size ? a : b, *hi = a->size <= b->size ? b : a;
    size_t hi_col = a->size <= b->size ? b_col : a_col;
    size_t sum = zmin(lo->size, rs), cpy = zmin(hi->size, rs);
    for (size_t j = 0; j < sum; j++) l_add(n, dft_limb(res, res_col, j), dft_limb(a, a_col, j), dft_limb(b, b_col, j));
    for (size_t j = sum; j < cpy; j++) memcpy(dft_limb(res, res_col, j), dft_limb(hi, hi_col, j), 32 * n);
    for (size_t j = cpy; j < rs; j++) memset(dft_limb(res, res_col, j), 0, 32 * n);
}
/* :472-486 */
void orc_ntt120_vec_znx_dft_add_assign(orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx_dft *a, size_t a_col) {
    size_t sum = zmin(res->size, a->size);
    for (size_t j = 0; j < sum; j++)
        l_add(res->n, dft_limb(res, res_col, j), dft_limb(res, res_col, j), dft_limb(a, a_col, j));
}
/* :488-520 */
void orc_ntt120_vec_znx_dft_add_scaled_assign(orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx_dft *a,
                                              size_t a_col, int64_t a_scale) {
    size_t n = res->n, rs = res->size, as = a->size;
    if (a_scale > 0) {
        size_t shift = zmin((size_t)a_scale, as);
        size_t m = zmin(as, rs);
        size_t sum = m > shift ? m - shift : 0;
        for (size_t j = 0; j < sum; j++)
            l_add(n, dft_limb(res, res_col, j), dft_limb(res, res_col, j), dft_limb(a, a_col, j + shift));
    } else if (a_scale < 0) {
        size_t shift = zmin((size_t)(-a_scale), rs);
        size_t sum = zmin(as, rs - shift);
        for (size_t j = 0; j < sum; j++)
            l_add(n, dft_limb(res, res_col, j + shift), dft_limb(res, res_col, j + shift), dft_limb(a, a_col, j));
    } else {
        orc_ntt120_vec_znx_dft_add_assign(res, res_col, a, a_col);
    }
}
/* :522-580 */
void orc_ntt120_vec_znx_dft_sub(orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx_dft *a, size_t a_col,
                                const orc_vec_znx_dft *b, size_t b_col) {
    size_t n = res->n, rs = res->size;
    if (a->size <= b->size) {
        size_t sum = zmin(a->size, rs), cpy = zmin(b->size, rs);
        for (size_t j = 0; j < sum; j++) l_sub(n, dft_limb(res, res_col, j), dft_limb(a, a_col, j), dft_limb(b, b_col, j));
        for (size_t j = sum; j < cpy; j++) l_neg(n, dft_limb(res, res_col, j), dft_limb(b, b_col, j));
        for (size_t j = cpy; j < rs; j++) memset(dft_limb(res, res_col, j), 0, 32 * n);
    } else {
        size_t sum = zmin(b->size, rs), cpy = zmin(a->size, rs);
        for (size_t j = 0; j < sum; j++) l_sub(n, dft_limb(res, res_col, j), dft_limb(a, a_col, j), dft_limb(b, b_col, j));
        for (size_t j = sum; j < cpy; j++) memcpy(dft_limb(res, res_col, j), dft_limb(a, a_col, j), 32 * n);
        for (size_t j = cpy; j < rs; j++) memset(dft_limb(res, res_col, j), 0, 32 * n);
    }
}
/* :582-596 */
void orc_ntt120_vec_znx_dft_sub_assign(orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx_dft *a, size_t a_col) {
    size_t sum = zmin(res->size, a->size);
    for (size_t j = 0; j < sum; j++)
        l_sub(res->n, dft_limb(res, res_col, j), dft_limb(res, res_col, j), dft_limb(a, a_col, j));
}
/* :598-616 */
void orc_ntt120_vec_znx_dft_sub_negate_assign(orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx_dft *a,
                                              size_t a_col) {
    size_t rs = res->size, sum = zmin(rs, a->size);
    for (size_t j = 0; j < sum; j++)
        l_sub(res->n, dft_limb(res, res_col, j), dft_limb(a, a_col, j), dft_limb(res, res_col, j));
    for (size_t j = sum; j < rs; j++) l_neg(res->n, dft_limb(res, res_col, j), dft_limb(res, res_col, j));
}
/* :618-645 */
void orc_ntt120_vec_znx_dft_copy(size_t step, size_t offset, orc_vec_znx_dft *res, size_t res_col,
                                 const orc_vec_znx_dft *a, size_t a_col) {
    size_t n = res->n;
    size_t steps = div_ceil(a->size, step);
    size_t min_steps = zmin(res->size, steps);
    for (size_t j = 0; j < min_steps; j++) {
        size_t limb = offset + j * step;
        if (limb < a->size)
            memcpy(dft_limb(res, res_col, j), dft_limb(a, a_col, limb), 32 * n);
        else
            memset(dft_limb(res, res_col, j), 0, 32 * n);
    }
    for (size_t j = min_steps; j < res->size; j++) memset(dft_limb(res, res_col, j), 0, 32 * n);
}
/* :647-652 */
void orc_ntt120_vec_znx_dft_zero(orc_vec_znx_dft *res, size_t res_col) {
    for (size_t j = 0; j < res->size; j++) memset(dft_limb(res, res_col, j), 0, 32 * res->n);
}

/* ---- svp.rs ---------------------------------------------------------------- */
/* :52-70 */
void orc_ntt120_svp_prepare(const orc_ntt120_module *m, orc_svp_ppol *res, size_t res_col, const orc_scalar_znx *a,
                            size_t a_col) {
    size_t n = res->n;
    uint64_t *tmp = (uint64_t *)malloc(32 * n);
    orc_ntt120_b_from_znx64(n, tmp, a->data + n * a_col);
    orc_ntt120_ntt(m, tmp);
    orc_ntt120_c_from_b(n, (uint32_t *)res->data + 8 * n * res_col, tmp);
    free(tmp);
}
/* :87-133 */
void orc_ntt120_svp_apply_dft_to_dft(const orc_ntt120_module *m, orc_vec_znx_dft *res, size_t res_col,
                                     const orc_svp_ppol *a, size_t a_col, const orc_vec_znx_dft *b, size_t b_col) {
    size_t n = res->n;
    size_t min_size = zmin(res->size, b->size);
    const uint32_t *a32 = (const uint32_t *)a->data + 8 * n * a_col;
    for (size_t j = 0; j < min_size; j++) {
        uint64_t *r = dft_limb(res, res_col, j);
        const uint32_t *b32 = (const uint32_t *)dft_limb(b, b_col, j);
        for (size_t i = 0; i < n; i++) mat1col_bbc(&m->bbc, 1, r + 4 * i, b32 + 8 * i, a32 + 8 * i);
    }
    for (size_t j = min_size; j < res->size; j++) memset(dft_limb(res, res_col, j), 0, 32 * n);
}
/* :148-180 */
void orc_ntt120_svp_apply_dft_to_dft_assign(const orc_ntt120_module *m, orc_vec_znx_dft *res, size_t res_col,
                                            const orc_svp_ppol *a, size_t a_col) {
    size_t n = res->n;
    const uint32_t *a32 = (const uint32_t *)a->data + 8 * n * a_col;
    for (size_t j = 0; j < res->size; j++) {
        uint64_t *r = dft_limb(res, res_col, j);
        for (size_t i = 0; i < n; i++) {
            uint64_t x[4], prod[4];
            memcpy(x, r + 4 * i, 32);
            mat1col_bbc(&m->bbc, 1, prod, (const uint32_t *)x, a32 + 8 * i);
            memcpy(r + 4 * i, prod, 32);
        }
    }
}

/* ---- vmp.rs ---------------------------------------------------------------- */
/* :64-119 */
void orc_ntt120_vmp_prepare(const orc_ntt120_module *m, orc_vmp_pmat *res, const orc_mat_znx *a) {
    size_t n = res->n;
    assert(a->n == n && res->cols_in == a->cols_in && res->rows == a->rows && res->cols_out == a->cols_out &&
           res->size == a->size);
    size_t nrows = a->cols_in * a->rows, ncols = a->cols_out * a->size;
    size_t n_blks = n / 2, offset = nrows * ncols * 16;
    uint32_t *pm = (uint32_t *)res->data;
    uint64_t *tmp = (uint64_t *)malloc(32 * n);
    uint32_t *tc = (uint32_t *)malloc(32 * n);
    for (size_t row_i = 0; row_i < nrows; row_i++)
        for (size_t col_i = 0; col_i < ncols; col_i++) {
            size_t pos = n * (row_i * ncols + col_i);
            orc_ntt120_b_from_znx64(n, tmp, a->data + pos);
            orc_ntt120_ntt(m, tmp);
            orc_ntt120_c_from_b(n, tc, tmp);
            size_t dst_base = (col_i == ncols - 1 && (ncols % 2) != 0)
                                  ? col_i * nrows * 16 + row_i * 16
                                  : (col_i / 2) * (nrows * 32) + row_i * 32 + (col_i % 2) * 16;
            for (size_t blk = 0; blk < n_blks; blk++) memcpy(pm + dst_base + blk * offset, tc + 16 * blk, 64);
        }
    free(tmp);
    free(tc);
}

/* :169-288 (OVERWRITE = true path; the add path is only used by convolution) */
static void vmp_core(size_t n, uint64_t *res, size_t res_size, const uint64_t *a, size_t a_size, const uint32_t *pmat,
                     size_t limb_offset, size_t nrows, size_t ncols, const bbc_meta *meta) {
    size_t n_blks = n / 2;
    size_t row_max = zmin(nrows, a_size);
    size_t col_max = zmin(ncols, res_size + limb_offset);
    if (limb_offset >= col_max) {
        memset(res, 0, 32 * n * res_size);
        return;
    }
    uint64_t out[16];
    uint64_t *ext = (uint64_t *)malloc(64 * (row_max ? row_max : 1));
    size_t offset = nrows * ncols * 16;
    for (size_t blk = 0; blk < n_blks; blk++) {
        const uint32_t *mb = pmat + blk * offset;
        for (size_t r = 0; r < row_max; r++) memcpy(ext + 8 * r, a + 4 * n * r + 8 * blk, 64); /* mat_vec.rs:472-483 */
        const uint32_t *e32 = (const uint32_t *)ext;
        if (limb_offset % 2 == 0) {
            size_t col_res = 0;
            for (size_t col_pmat = limb_offset; col_pmat + 1 < col_max; col_pmat += 2, col_res += 2) {
                mat2cols_x2_bbc(meta, row_max, out, e32, mb + col_pmat * (nrows * 16));
                memcpy(res + col_res * 4 * n + 8 * blk, out, 64);
                memcpy(res + (col_res + 1) * 4 * n + 8 * blk, out + 8, 64);
            }
        } else {
            mat2cols_x2_bbc(meta, row_max, out, e32, mb + (limb_offset - 1) * (nrows * 16));
            memcpy(res + 8 * blk, out + 8, 64);
            size_t col_res = 1;
            for (size_t col_pmat = limb_offset + 1; col_pmat + 1 < col_max; col_pmat += 2, col_res += 2) {
                mat2cols_x2_bbc(meta, row_max, out, e32, mb + col_pmat * (nrows * 16));
                memcpy(res + col_res * 4 * n + 8 * blk, out, 64);
                memcpy(res + (col_res + 1) * 4 * n + 8 * blk, out + 8, 64);
            }
        }
        if (col_max % 2 != 0) {
            size_t last = col_max - 1;
            if (last >= limb_offset) {
                mat1col_x2_bbc(meta, row_max, out, e32, mb + last * (nrows * 16));
                memcpy(res + (last - limb_offset) * 4 * n + 8 * blk, out, 64);
            }
        }
    }
    for (size_t c = col_max - limb_offset; c < res_size; c++) memset(res + c * 4 * n, 0, 32 * n);
    free(ext);
}

/* :301-341 */
void orc_ntt120_vmp_apply_dft_to_dft(const orc_ntt120_module *m, orc_vec_znx_dft *res, const orc_vec_znx_dft *a,
                                     const orc_vmp_pmat *pmat, size_t limb_offset) {
    size_t n = res->n;
    assert(pmat->n == n && a->n == n);
    size_t nrows = pmat->cols_in * pmat->rows, ncols = pmat->cols_out * pmat->size;
    vmp_core(n, (uint64_t *)res->data, res->cols * res->size, (const uint64_t *)a->data, a->cols * a->size,
             (const uint32_t *)pmat->data, limb_offset * pmat->cols_out, nrows, ncols, &m->bbc);
}

/* ---- vec_znx_big.rs (i128) ------------------------------------------------- */
/* wrapping i128 helpers: C signed overflow is UB, so go through u128 */
static inline i128 wadd(i128 a, i128 b) { return (i128)((u128)a + (u128)b); }
static inline i128 wsub(i128 a, i128 b) { return (i128)((u128)a - (u128)b); }
static inline i128 wshl(i128 a, unsigned s) { return (i128)((u128)a << s); }
/* reference/znx/normalization.rs:14-21 */
static inline i128 get_digit(size_t k, i128 x) { return (i128)((u128)x << (128 - k)) >> (128 - k); }
static inline i128 get_carry(size_t k, i128 x, i128 d) { return wsub(x, d) >> k; }

/* :1128-1140 */
void orc_ntt120_vec_znx_big_add_small_assign(orc_vec_znx_big *res, size_t res_col, const orc_vec_znx *a, size_t a_col) {
    size_t sum = zmin(res->size, a->size);
    for (size_t j = 0; j < sum; j++) {
        i128 *r = big_limb(res, res_col, j);
        const int64_t *x = znx_limb(a, a_col, j);
        for (size_t i = 0; i < res->n; i++) r[i] = wadd(r[i], (i128)x[i]);
    }
}
/* :1243-1261 */
void orc_ntt120_vec_znx_big_from_small(orc_vec_znx_big *res, size_t res_col, const orc_vec_znx *a, size_t a_col) {
    size_t mn = zmin(res->size, a->size);
    for (size_t j = 0; j < mn; j++) {
        i128 *r = big_limb(res, res_col, j);
        const int64_t *x = znx_limb(a, a_col, j);
        for (size_t i = 0; i < res->n; i++) r[i] = (i128)x[i];
    }
    for (size_t j = mn; j < res->size; j++) memset(big_limb(res, res_col, j), 0, 16 * res->n);
}
void orc_ntt120_vec_znx_big_add_assign(orc_vec_znx_big *res, size_t res_col, const orc_vec_znx_big *a, size_t a_col) {
    size_t sum = zmin(res->size, a->size);
    for (size_t j = 0; j < sum; j++) {
        i128 *r = big_limb(res, res_col, j);
        const i128 *x = big_limb(a, a_col, j);
        for (size_t i = 0; i < res->n; i++) r[i] = wadd(r[i], x[i]);
    }
}
void orc_ntt120_vec_znx_big_sub_assign(orc_vec_znx_big *res, size_t res_col, const orc_vec_znx_big *a, size_t a_col) {
    size_t sum = zmin(res->size, a->size);
    for (size_t j = 0; j < sum; j++) {
        i128 *r = big_limb(res, res_col, j);
        const i128 *x = big_limb(a, a_col, j);
        for (size_t i = 0; i < res->n; i++) r[i] = wsub(r[i], x[i]);
    }
}
void orc_ntt120_vec_znx_big_negate_assign(orc_vec_znx_big *res, size_t res_col) {
    for (size_t j = 0; j < res->size; j++) {
        i128 *r = big_limb(res, res_col, j);
        for (size_t i = 0; i < res->n; i++) r[i] = wsub(0, r[i]);
    }
}

/* normalize: instantiate the shared template for i128 (vec_znx_big.rs:47-803) */
#define NT i128
#define NUT u128
#define NBITS 128
#define NBIG orc_vec_znx_big
#define NLIMB(v, c, l) big_limb(v, c, l)
#define NF(name) name##_i128
#include "normalize_impl.inc"
#undef NT
#undef NUT
#undef NBITS
#undef NBIG
#undef NLIMB
#undef NF

/* :1383-1461 */
void orc_ntt120_vec_znx_big_normalize(orc_vec_znx *res, size_t res_base2k, int64_t res_offset, size_t res_col,
                                      const orc_vec_znx_big *a, size_t a_base2k, size_t a_col, int op) {
    i128 *scratch = (i128 *)malloc(3 * res->n * sizeof(i128));
    if (res_base2k == a_base2k)
        normalize_inter_i128(res_base2k, res, res_offset, res_col, a, a_col, scratch, op);
    else
        normalize_cross_i128(res, res_base2k, res_offset, res_col, a, a_base2k, a_col, scratch, op);
    free(scratch);
}

/* ---- convolution.rs (bivariate convolution, NTT120) ------------------------------------------------------------------
 * CnvPVecL / CnvPVecR are opaque prepared layouts; this restatement stores both as q120b limbs in the VecZnxDft layout
 * (limb-major, column-minor) and forms the q120c view of the right operand on the fly (c_from_b), which is what
 * prepare_right stores (convolution.rs:120-157).  Products go through the same bbc kernel (mat1col_x2_bbc). */

/* arithmetic.rs:64-86 (b_from_znx64_masked_ref): the i64 is ANDed with the mask, then mapped like b_from_znx64 */
static void b_from_znx64_masked(size_t nn, uint64_t *res, const int64_t *x, int64_t mask) {
    int64_t *t = (int64_t *)malloc(8 * nn);
    for (size_t i = 0; i < nn; i++) t[i] = x[i] & mask;
    orc_ntt120_b_from_znx64(nn, res, t);
    free(t);
}

/* convolution.rs:66-100 (prepare_left), :120-157 (prepare_right), :177-236 (prepare_self): all columns of res, limbs
 * [0, min(res.size, a.size)), the last active limb masked, the remaining limbs zero */
void orc_ntt120_cnv_prepare(const orc_ntt120_module *m, orc_vec_znx_dft *res, const orc_vec_znx *a, int64_t mask) {
    size_t n = res->n, min_size = zmin(res->size, a->size);
    for (size_t col = 0; col < res->cols; col++) {
        for (size_t j = 0; j < min_size; j++) {
            uint64_t *r = dft_limb(res, col, j);
            if (j + 1 == min_size) b_from_znx64_masked(n, r, znx_limb(a, col, j), mask);
            else orc_ntt120_b_from_znx64(n, r, znx_limb(a, col, j));
            orc_ntt120_ntt(m, r);
        }
        for (size_t j = min_size; j < res->size; j++) memset(dft_limb(res, col, j), 0, 32 * n);
    }
}

/* pack one x2 block (two NTT coefficients) of every limb: left = canonical residues with a zero high half
 * (poulpy-cpu-ref/src/ntt120/prim.rs:223-240, pairwise :259-280), right = q120c in REVERSED limb order (:245-254, :285-299) */
static void cnv_pack_left(uint32_t *dst, const orc_vec_znx_dft *a, size_t col_i, size_t col_j, size_t blk) {
    for (size_t row = 0; row < a->size; row++) {
        const uint64_t *x = dft_limb(a, col_i, row) + 8 * blk, *y = col_j == col_i ? NULL : dft_limb(a, col_j, row) + 8 * blk;
        for (int c = 0; c < 2; c++)
            for (int k = 0; k < 4; k++) {
                uint64_t q = ORC_Q[k], s = x[4 * c + k] % q;
                if (y) {
                    s += y[4 * c + k] % q;
                    if (s >= q) s -= q;
                }
                dst[16 * row + 8 * c + 2 * k] = (uint32_t)s;
                dst[16 * row + 8 * c + 2 * k + 1] = 0;
            }
    }
}
static void cnv_pack_right(uint32_t *dst, const orc_vec_znx_dft *b, size_t col_i, size_t col_j, size_t blk) {
    for (size_t row = 0; row < b->size; row++) {
        size_t src = b->size - 1 - row;
        uint32_t ci[16], cj[16];
        orc_ntt120_c_from_b(2, ci, dft_limb(b, col_i, src) + 8 * blk);
        if (col_j != col_i) {
            orc_ntt120_c_from_b(2, cj, dft_limb(b, col_j, src) + 8 * blk);
            for (int t = 0; t < 16; t++) ci[t] += cj[t];
        }
        memcpy(dst + 16 * row, ci, 64);
    }
}
/* convolution.rs:256-335 (apply_dft) and :441-557 (pairwise; col_i == col_j falls back to apply_dft with both columns equal):
 * res[res_col, k] = sum_j a[k_abs - j] (.) b[j], k_abs = k + min(cnv_offset, bound) */
static void cnv_apply_core(const orc_ntt120_module *m, size_t cnv_offset, orc_vec_znx_dft *res, size_t res_col,
                           const orc_vec_znx_dft *a, size_t a_i, size_t a_j, const orc_vec_znx_dft *b, size_t b_i, size_t b_j) {
    size_t n = res->n, res_size = res->size, a_size = a->size, b_size = b->size;
    if (res_size == 0 || a_size == 0 || b_size == 0) {
        for (size_t j = 0; j < res_size; j++) memset(dft_limb(res, res_col, j), 0, 32 * n);
        return;
    }
    size_t bound = a_size + b_size - 1, offset = zmin(cnv_offset, bound);
    size_t min_size = zmin(res_size, bound + 1 > offset ? bound + 1 - offset : 0);
    uint32_t *a_tmp = (uint32_t *)malloc(64 * a_size), *b_tmp = (uint32_t *)malloc(64 * b_size);
    for (size_t blk = 0; blk < n / 2; blk++) {
        cnv_pack_left(a_tmp, a, a_i, a_j, blk);
        cnv_pack_right(b_tmp, b, b_i, b_j, blk);
        for (size_t k = 0; k < min_size; k++) {
            size_t k_abs = k + offset;
            size_t j_min = k_abs > a_size - 1 ? k_abs - (a_size - 1) : 0;
            size_t j_max = zmin(k_abs + 1, b_size);
            uint64_t *r = dft_limb(res, res_col, k) + 8 * blk;
            if (j_max <= j_min) { /* ell == 0: the bbc kernel writes the reduction of an empty sum */
                memset(r, 0, 64);
                continue;
            }
            size_t ell = j_max - j_min, a_start = k_abs + 1 - j_max, b_start = b_size - j_max;
            mat1col_x2_bbc(&m->bbc, ell, r, a_tmp + 16 * a_start, b_tmp + 16 * b_start);
        }
    }
    free(a_tmp);
    free(b_tmp);
    for (size_t j = min_size; j < res_size; j++) memset(dft_limb(res, res_col, j), 0, 32 * n);
}
void orc_ntt120_cnv_apply_dft(const orc_ntt120_module *m, size_t cnv_offset, orc_vec_znx_dft *res, size_t res_col,
                              const orc_vec_znx_dft *a, size_t a_col, const orc_vec_znx_dft *b, size_t b_col) {
    cnv_apply_core(m, cnv_offset, res, res_col, a, a_col, a_col, b, b_col, b_col);
}
void orc_ntt120_cnv_pairwise_apply_dft(const orc_ntt120_module *m, size_t cnv_offset, orc_vec_znx_dft *res, size_t res_col,
                                       const orc_vec_znx_dft *a, const orc_vec_znx_dft *b, size_t col_i, size_t col_j) {
    cnv_apply_core(m, cnv_offset, res, res_col, a, col_i, col_j, b, col_i, col_j);
}
/* convolution.rs:361-410: coefficient-domain product with a constant limb vector, i128 accumulators */
void orc_ntt120_cnv_by_const_apply(size_t cnv_offset, orc_vec_znx_big *res, size_t res_col, const orc_vec_znx *a, size_t a_col,
                                   const int64_t *b, size_t b_size) {
    size_t n = res->n, res_size = res->size, a_size = a->size;
    if (res_size == 0 || a_size == 0 || b_size == 0) {
        for (size_t j = 0; j < res_size; j++) memset(big_limb(res, res_col, j), 0, 16 * n);
        return;
    }
    size_t bound = a_size + b_size - 1, offset = zmin(cnv_offset, bound);
    size_t min_size = zmin(res_size, bound + 1 > offset ? bound + 1 - offset : 0);
    for (size_t k = 0; k < min_size; k++) {
        size_t k_abs = k + offset;
        size_t j_min = k_abs > a_size - 1 ? k_abs - (a_size - 1) : 0, j_max = zmin(k_abs + 1, b_size);
        i128 *r = big_limb(res, res_col, k);
        for (size_t i = 0; i < n; i++) {
            i128 acc = 0;
            for (size_t j = j_min; j < j_max; j++) acc = wadd(acc, (i128)znx_limb(a, a_col, k_abs - j)[i] * (i128)b[j]);
            r[i] = acc;
        }
    }
    for (size_t j = min_size; j < res_size; j++) memset(big_limb(res, res_col, j), 0, 16 * n);
}
