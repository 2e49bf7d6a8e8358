/*
 * fft64.c -- CPU restatement of poulpy-cpu-ref's FFT64 backend arithmetic.
 * TEST INFRASTRUCTURE ONLY (see poulpy_oracle.h).
 *
 * Compile with -ffp-contract=off: the reference is Rust f64 without FMA
 * contraction, and the butterflies below keep its operation order.
 *
 * Restates (paths under poulpy-cpu-ref/src/reference/fft64/):
 *   reim/mod.rs:50-64 (frac_rev_bits), reim/table_fft.rs, reim/table_ifft.rs,
 *   reim/fft_ref.rs, reim/ifft_ref.rs, reim/conversion.rs, reim/fft_vec.rs,
 *   reim4/arithmetic_ref.rs, vec_znx_dft.rs, svp.rs, vmp.rs, vec_znx_big.rs
 */
#include "poulpy_oracle.h"

#include <assert.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

struct orc_fft64_module {
    size_t n, m;
    double *omg_f, *omg_i; /* 2m doubles each */
};

static const double TWO_PI = 2.0 * M_PI;

/* reim/mod.rs:50-64 */
static double frac_rev_bits(size_t x) {
    if (x == 0) return 0.0;
    if (x == 1) return 0.5;
    if (x % 2 == 0) return frac_rev_bits(x >> 1) * 0.5;
    return frac_rev_bits(x >> 1) * 0.5 + 0.5;
}
static size_t log2_ceil(size_t m) { /* (usize::BITS - (m-1).leading_zeros()) */
    size_t l = 0;
    while (((size_t)1 << l) < m) l++;
    return l;
}

/* ---- table_fft.rs:85-216 --------------------------------------------------- */
static size_t fill_fft2(double j, double *omg, size_t pos) {
    double a = j / 2.0;
    omg[pos] = cos(TWO_PI * a);
    omg[pos + 1] = sin(TWO_PI * a);
    return pos + 2;
}
static size_t fill_fft4(double j, double *omg, size_t pos) {
    double a1 = j / 2.0, a2 = j / 4.0;
    omg[pos] = cos(TWO_PI * a1);
    omg[pos + 1] = sin(TWO_PI * a1);
    omg[pos + 2] = cos(TWO_PI * a2);
    omg[pos + 3] = sin(TWO_PI * a2);
    return pos + 4;
}
static size_t fill_fft8(double j, double *omg, size_t pos) {
    const double e8 = 1.0 / 8.0;
    double a1 = j / 2.0, a2 = j / 4.0, a4 = j / 8.0;
    double *o = omg + pos;
    o[0] = cos(TWO_PI * a1);
    o[1] = sin(TWO_PI * a1);
    o[2] = cos(TWO_PI * a2);
    o[3] = sin(TWO_PI * a2);
    o[4] = cos(TWO_PI * a4);
    o[5] = cos(TWO_PI * (a4 + e8));
    o[6] = sin(TWO_PI * a4);
    o[7] = sin(TWO_PI * (a4 + e8));
    return pos + 8;
}
static size_t fill_fft16(double j, double *omg, size_t pos) {
    const double e8 = 1.0 / 8.0, e16 = 1.0 / 16.0;
    double a1 = j / 2.0, a2 = j / 4.0, a4 = j / 8.0, a8 = j / 16.0;
    double *o = omg + pos;
    o[0] = cos(TWO_PI * a1);
    o[1] = sin(TWO_PI * a1);
    o[2] = cos(TWO_PI * a2);
    o[3] = sin(TWO_PI * a2);
    o[4] = cos(TWO_PI * a4);
    o[5] = sin(TWO_PI * a4);
    o[6] = cos(TWO_PI * (a4 + e8));
    o[7] = sin(TWO_PI * (a4 + e8));
    o[8] = cos(TWO_PI * a8);
    o[9] = cos(TWO_PI * (a8 + e8));
    o[10] = cos(TWO_PI * (a8 + e16));
    o[11] = cos(TWO_PI * (a8 + e8 + e16));
    o[12] = sin(TWO_PI * a8);
    o[13] = sin(TWO_PI * (a8 + e8));
    o[14] = sin(TWO_PI * (a8 + e16));
    o[15] = sin(TWO_PI * (a8 + e8 + e16));
    return pos + 16;
}
static size_t fill_fft_bfs16(size_t m, double j, double *omg, size_t pos) {
    size_t log_m = log2_ceil(m);
    size_t mm = m;
    double jj = j;
    if (log_m % 2 != 0) {
        size_t h = mm >> 1;
        double j2 = jj * 0.5;
        omg[pos] = cos(TWO_PI * j2);
        omg[pos + 1] = sin(TWO_PI * j2);
        pos += 2;
        mm = h;
        jj = j2;
    }
    while (mm > 16) {
        size_t h = mm >> 2;
        double j4 = jj * (1.0 / 4.0);
        for (size_t i = 0; i < m; i += mm) {
            double rs0 = j4 + frac_rev_bits(i / mm) * (1.0 / 4.0);
            double rs1 = 2.0 * rs0;
            omg[pos] = cos(TWO_PI * rs1);
            omg[pos + 1] = sin(TWO_PI * rs1);
            omg[pos + 2] = cos(TWO_PI * rs0);
            omg[pos + 3] = sin(TWO_PI * rs0);
            pos += 4;
        }
        mm = h;
        jj = j4;
    }
    for (size_t i = 0; i < m; i += 16) {
        double jl = jj + frac_rev_bits(i >> 4);
        fill_fft16(jl, omg, pos);
        pos += 16;
    }
    return pos;
}
static size_t fill_fft_rec16(size_t m, double j, double *omg, size_t pos) {
    if (m <= 2048) return fill_fft_bfs16(m, j, omg, pos);
    size_t h = m >> 1;
    double s = j * 0.5;
    omg[pos] = cos(TWO_PI * s);
    omg[pos + 1] = sin(TWO_PI * s);
    pos += 2;
    pos = fill_fft_rec16(h, s, omg, pos);
    pos = fill_fft_rec16(h, s + 0.5, omg, pos);
    return pos;
}

/* ---- table_ifft.rs:84-216 --------------------------------------------------- */
static size_t fill_ifft2(double j, double *omg, size_t pos) {
    double a = j / 4.0;          /* j / exp2(2) */
    double four_pi = 4.0 * M_PI; /* exp2(2) * PI */
    omg[pos] = cos(four_pi * a);
    omg[pos + 1] = -sin(four_pi * a);
    return pos + 2;
}
static size_t fill_ifft4(double j, double *omg, size_t pos) {
    double a1 = j / 2.0, a2 = j / 4.0;
    omg[pos] = cos(TWO_PI * a2);
    omg[pos + 1] = -sin(TWO_PI * a2);
    omg[pos + 2] = cos(TWO_PI * a1);
    omg[pos + 3] = -sin(TWO_PI * a1);
    return pos + 4;
}
static size_t fill_ifft8(double j, double *omg, size_t pos) {
    const double e8 = 1.0 / 8.0;
    double a1 = j / 2.0, a2 = j / 4.0, a4 = j / 8.0;
    double *o = omg + pos;
    o[0] = cos(TWO_PI * a4);
    o[1] = cos(TWO_PI * (a4 + e8));
    o[2] = -sin(TWO_PI * a4);
    o[3] = -sin(TWO_PI * (a4 + e8));
    o[4] = cos(TWO_PI * a2);
    o[5] = -sin(TWO_PI * a2);
    o[6] = cos(TWO_PI * a1);
    o[7] = -sin(TWO_PI * a1);
    return pos + 8;
}
static size_t fill_ifft16(double j, double *omg, size_t pos) {
    const double e8 = 1.0 / 8.0, e16 = 1.0 / 16.0;
    double a1 = j / 2.0, a2 = j / 4.0, a4 = j / 8.0, a8 = j / 16.0;
    double *o = omg + pos;
    o[0] = cos(TWO_PI * a8);
    o[1] = cos(TWO_PI * (a8 + e8));
    o[2] = cos(TWO_PI * (a8 + e16));
    o[3] = cos(TWO_PI * (a8 + e8 + e16));
    o[4] = -sin(TWO_PI * a8);
    o[5] = -sin(TWO_PI * (a8 + e8));
    o[6] = -sin(TWO_PI * (a8 + e16));
    o[7] = -sin(TWO_PI * (a8 + e8 + e16));
    o[8] = cos(TWO_PI * a4);
    o[9] = -sin(TWO_PI * a4);
    o[10] = cos(TWO_PI * (a4 + e8));
    o[11] = -sin(TWO_PI * (a4 + e8));
    o[12] = cos(TWO_PI * a2);
    o[13] = -sin(TWO_PI * a2);
    o[14] = cos(TWO_PI * a1);
    o[15] = -sin(TWO_PI * a1);
    return pos + 16;
}
static size_t fill_ifft_bfs16(size_t m, double j, double *omg, size_t pos) {
    size_t log_m = log2_ceil(m);
    double jj = j * 16.0 / (double)m;
    for (size_t i = 0; i < m; i += 16) {
        double jl = jj + frac_rev_bits(i >> 4);
        fill_ifft16(jl, omg, pos);
        pos += 16;
    }
    size_t h = 16, m_half = m >> 1;
    while (h < m_half) {
        size_t mm = h << 2;
        for (size_t i = 0; i < m; i += mm) {
            double rs0 = jj + frac_rev_bits(i / mm) / 4.0;
            double rs1 = 2.0 * rs0;
            omg[pos] = cos(TWO_PI * rs0);
            omg[pos + 1] = -sin(TWO_PI * rs0);
            omg[pos + 2] = cos(TWO_PI * rs1);
            omg[pos + 3] = -sin(TWO_PI * rs1);
            pos += 4;
        }
        h = mm;
        jj = jj * 4.0;
    }
    if (log_m % 2 != 0) {
        omg[pos] = cos(TWO_PI * jj);
        omg[pos + 1] = -sin(TWO_PI * jj);
        pos += 2;
        jj = jj * 2.0;
    }
    assert(jj == j);
    return pos;
}
static size_t fill_ifft_rec16(size_t m, double j, double *omg, size_t pos) {
    if (m <= 2048) return fill_ifft_bfs16(m, j, omg, pos);
    size_t h = m >> 1;
    double s = j / 2.0;
    pos = fill_ifft_rec16(h, s, omg, pos);
    pos = fill_ifft_rec16(h, s + 0.5, omg, pos);
    omg[pos] = cos(TWO_PI * s);
    omg[pos + 1] = -sin(TWO_PI * s);
    pos += 2;
    return pos;
}

/* table_fft.rs:38-69, table_ifft.rs:39-82, fft64/module.rs:62-69 */
orc_fft64_module *orc_fft64_new(size_t n) {
    assert(n >= 2 && (n & (n - 1)) == 0);
    orc_fft64_module *md = (orc_fft64_module *)calloc(1, sizeof *md);
    size_t m = n / 2;
    md->n = n;
    md->m = m;
    md->omg_f = (double *)calloc(2 * m + 16, sizeof(double));
    md->omg_i = (double *)calloc(2 * m + 16, sizeof(double));
    const double quarter = 0.25;
    if (m <= 16) {
        if (m == 2) {
            fill_fft2(quarter, md->omg_f, 0);
            fill_ifft2(quarter, md->omg_i, 0);
        } else if (m == 4) {
            fill_fft4(quarter, md->omg_f, 0);
            fill_ifft4(quarter, md->omg_i, 0);
        } else if (m == 8) {
            fill_fft8(quarter, md->omg_f, 0);
            fill_ifft8(quarter, md->omg_i, 0);
        } else if (m == 16) {
            fill_fft16(quarter, md->omg_f, 0);
            fill_ifft16(quarter, md->omg_i, 0);
        }
    } else if (m <= 2048) {
        fill_fft_bfs16(m, quarter, md->omg_f, 0);
        fill_ifft_bfs16(m, quarter, md->omg_i, 0);
    } else {
        fill_fft_rec16(m, quarter, md->omg_f, 0);
        fill_ifft_rec16(m, quarter, md->omg_i, 0);
    }
    return md;
}
void orc_fft64_free(orc_fft64_module *m) {
    if (!m) return;
    free(m->omg_f);
    free(m->omg_i);
    free(m);
}
const double *orc_fft64_omg(const orc_fft64_module *m, int inverse, size_t *len) {
    if (len) *len = 2 * m->m;
    return inverse ? m->omg_i : m->omg_f;
}

/* ---- fft_ref.rs ---------------------------------------------------------------- */
/* :60-67 */
static inline void ctw(double *ra, double *ia, double *rb, double *ib, double wr, double wi) {
    double dr = *rb * wr - *ib * wi;
    double di = *rb * wi + *ib * wr;
    *rb = *ra - dr;
    *ib = *ia - di;
    *ra = *ra + dr;
    *ia = *ia + di;
}
/* :70-77 */
static inline void citw(double *ra, double *ia, double *rb, double *ib, double wr, double wi) {
    double dr = *rb * wi + *ib * wr;
    double di = *rb * wr - *ib * wi;
    *rb = *ra + dr;
    *ib = *ia - di;
    *ra = *ra - dr;
    *ia = *ia + di;
}
#define CT(a, b, wr, wi) ctw(&re[a], &im[a], &re[b], &im[b], wr, wi)
#define CIT(a, b, wr, wi) citw(&re[a], &im[a], &re[b], &im[b], wr, wi)

static void fft2(double *re, double *im, const double *o) { CT(0, 1, o[0], o[1]); }
static void fft4(double *re, double *im, const double *o) {
    CT(0, 2, o[0], o[1]);
    CT(1, 3, o[0], o[1]);
    CT(0, 1, o[2], o[3]);
    CIT(2, 3, o[2], o[3]);
}
static void fft8(double *re, double *im, const double *o) {
    for (int i = 0; i < 4; i++) CT(i, i + 4, o[0], o[1]);
    CT(0, 2, o[2], o[3]);
    CT(1, 3, o[2], o[3]);
    CIT(4, 6, o[2], o[3]);
    CIT(5, 7, o[2], o[3]);
    CT(0, 1, o[4], o[6]);
    CIT(2, 3, o[4], o[6]);
    CT(4, 5, o[5], o[7]);
    CIT(6, 7, o[5], o[7]);
}
/* :143-244 */
static void fft16(double *re, double *im, const double *o) {
    for (int i = 0; i < 8; i++) CT(i, i + 8, o[0], o[1]);
    for (int i = 0; i < 4; i++) CT(i, i + 4, o[2], o[3]);
    for (int i = 8; i < 12; i++) CIT(i, i + 4, o[2], o[3]);
    CT(0, 2, o[4], o[5]);
    CT(1, 3, o[4], o[5]);
    CT(8, 10, o[6], o[7]);
    CT(9, 11, o[6], o[7]);
    CIT(4, 6, o[4], o[5]);
    CIT(5, 7, o[4], o[5]);
    CIT(12, 14, o[6], o[7]);
    CIT(13, 15, o[6], o[7]);
    CT(0, 1, o[8], o[12]);
    CT(4, 5, o[9], o[13]);
    CT(8, 9, o[10], o[14]);
    CT(12, 13, o[11], o[15]);
    CIT(2, 3, o[8], o[12]);
    CIT(6, 7, o[9], o[13]);
    CIT(10, 11, o[10], o[14]);
    CIT(14, 15, o[11], o[15]);
}
/* :280-293 */
static void twiddle_fft(size_t h, double *re, double *im, const double *o) {
    for (size_t i = 0; i < h; i++) ctw(&re[i], &im[i], &re[h + i], &im[h + i], o[0], o[1]);
}
/* :296-316 */
static void bitwiddle_fft(size_t h, double *re, double *im, const double *o) {
    for (size_t i = 0; i < h; i++) {
        ctw(&re[i], &im[i], &re[2 * h + i], &im[2 * h + i], o[0], o[1]);
        ctw(&re[h + i], &im[h + i], &re[3 * h + i], &im[3 * h + i], o[0], o[1]);
    }
    for (size_t i = 0; i < h; i++) {
        ctw(&re[i], &im[i], &re[h + i], &im[h + i], o[2], o[3]);
        citw(&re[2 * h + i], &im[2 * h + i], &re[3 * h + i], &im[3 * h + i], o[2], o[3]);
    }
}
/* :247-277 */
static size_t fft_bfs16(size_t m, double *re, double *im, const double *omg, size_t pos) {
    size_t log_m = log2_ceil(m), mm = m;
    if (log_m % 2 != 0) {
        size_t h = mm >> 1;
        twiddle_fft(h, re, im, omg + pos);
        pos += 2;
        mm = h;
    }
    while (mm > 16) {
        size_t h = mm >> 2;
        for (size_t off = 0; off < m; off += mm) {
            bitwiddle_fft(h, re + off, im + off, omg + pos);
            pos += 4;
        }
        mm = h;
    }
    for (size_t off = 0; off < m; off += 16) {
        fft16(re + off, im + off, omg + pos);
        pos += 16;
    }
    return pos;
}
/* :46-57 */
static size_t fft_rec16(size_t m, double *re, double *im, const double *omg, size_t pos) {
    if (m <= 2048) return fft_bfs16(m, re, im, omg, pos);
    size_t h = m >> 1;
    twiddle_fft(h, re, im, omg + pos);
    pos += 2;
    pos = fft_rec16(h, re, im, omg, pos);
    pos = fft_rec16(h, re + h, im + h, omg, pos);
    return pos;
}
/* :25-43 */
void orc_fft64_fft(const orc_fft64_module *md, double *data) {
    size_t m = md->m;
    double *re = data, *im = data + m;
    const double *o = md->omg_f;
    if (m <= 16) {
        if (m == 2) fft2(re, im, o);
        else if (m == 4) fft4(re, im, o);
        else if (m == 8) fft8(re, im, o);
        else if (m == 16) fft16(re, im, o);
    } else if (m <= 2048) {
        fft_bfs16(m, re, im, o, 0);
    } else {
        fft_rec16(m, re, im, o, 0);
    }
}

/* ---- ifft_ref.rs ---------------------------------------------------------------- */
/* :91-98 */
static inline void itw(double *ra, double *ia, double *rb, double *ib, double wr, double wi) {
    double rd = *ra - *rb, id = *ia - *ib;
    *ra = *ra + *rb;
    *ia = *ia + *ib;
    *rb = rd * wr - id * wi;
    *ib = rd * wi + id * wr;
}
/* :101-108 */
static inline void iitw(double *ra, double *ia, double *rb, double *ib, double wr, double wi) {
    double rd = *ra - *rb, id = *ia - *ib;
    *ra = *ra + *rb;
    *ia = *ia + *ib;
    *rb = rd * wi + id * wr;
    *ib = -rd * wr + id * wi;
}
#define IT(a, b, wr, wi) itw(&re[a], &im[a], &re[b], &im[b], wr, wi)
#define IIT(a, b, wr, wi) iitw(&re[a], &im[a], &re[b], &im[b], wr, wi)
static void ifft2(double *re, double *im, const double *o) { IT(0, 1, o[0], o[1]); }
static void ifft4(double *re, double *im, const double *o) {
    IT(0, 1, o[0], o[1]);
    IIT(2, 3, o[0], o[1]);
    IT(0, 2, o[2], o[3]);
    IT(1, 3, o[2], o[3]);
}
static void ifft8(double *re, double *im, const double *o) {
    IT(0, 1, o[0], o[2]);
    IIT(2, 3, o[0], o[2]);
    IT(4, 5, o[1], o[3]);
    IIT(6, 7, o[1], o[3]);
    IT(0, 2, o[4], o[5]);
    IT(1, 3, o[4], o[5]);
    IIT(4, 6, o[4], o[5]);
    IIT(5, 7, o[4], o[5]);
    for (int i = 0; i < 4; i++) IT(i, i + 4, o[6], o[7]);
}
/* :170-268 */
static void ifft16(double *re, double *im, const double *o) {
    IT(0, 1, o[0], o[4]);
    IIT(2, 3, o[0], o[4]);
    IT(4, 5, o[1], o[5]);
    IIT(6, 7, o[1], o[5]);
    IT(8, 9, o[2], o[6]);
    IIT(10, 11, o[2], o[6]);
    IT(12, 13, o[3], o[7]);
    IIT(14, 15, o[3], o[7]);
    IT(0, 2, o[8], o[9]);
    IT(1, 3, o[8], o[9]);
    IIT(4, 6, o[8], o[9]);
    IIT(5, 7, o[8], o[9]);
    IT(8, 10, o[10], o[11]);
    IT(9, 11, o[10], o[11]);
    IIT(12, 14, o[10], o[11]);
    IIT(13, 15, o[10], o[11]);
    for (int i = 0; i < 4; i++) IT(i, i + 4, o[12], o[13]);
    for (int i = 8; i < 12; i++) IIT(i, i + 4, o[12], o[13]);
    for (int i = 0; i < 8; i++) IT(i, i + 8, o[14], o[15]);
}
static void inv_twiddle_ifft(size_t h, double *re, double *im, const double *o) {
    for (size_t i = 0; i < h; i++) itw(&re[i], &im[i], &re[h + i], &im[h + i], o[0], o[1]);
}
static void inv_bitwiddle_ifft(size_t h, double *re, double *im, const double *o) {
    for (size_t i = 0; i < h; i++) {
        itw(&re[i], &im[i], &re[h + i], &im[h + i], o[0], o[1]);
        iitw(&re[2 * h + i], &im[2 * h + i], &re[3 * h + i], &im[3 * h + i], o[0], o[1]);
    }
    for (size_t i = 0; i < h; i++) {
        itw(&re[i], &im[i], &re[2 * h + i], &im[2 * h + i], o[2], o[3]);
        itw(&re[h + i], &im[h + i], &re[3 * h + i], &im[3 * h + i], o[2], o[3]);
    }
}
/* :56-88 */
static size_t ifft_bfs16(size_t m, double *re, double *im, const double *omg, size_t pos) {
    size_t log_m = log2_ceil(m);
    for (size_t off = 0; off < m; off += 16) {
        ifft16(re + off, im + off, omg + pos);
        pos += 16;
    }
    size_t h = 16, m_half = m >> 1;
    while (h < m_half) {
        size_t mm = h << 2;
        for (size_t off = 0; off < m; off += mm) {
            inv_bitwiddle_ifft(h, re + off, im + off, omg + pos);
            pos += 4;
        }
        h = mm;
    }
    if (log_m % 2 != 0) {
        inv_twiddle_ifft(h, re, im, omg + pos);
        pos += 2;
    }
    return pos;
}
/* :43-54 */
static size_t ifft_rec16(size_t m, double *re, double *im, const double *omg, size_t pos) {
    if (m <= 2048) return ifft_bfs16(m, re, im, omg, pos);
    size_t h = m >> 1;
    pos = ifft_rec16(h, re, im, omg, pos);
    pos = ifft_rec16(h, re + h, im + h, omg, pos);
    inv_twiddle_ifft(h, re, im, omg + pos);
    pos += 2;
    return pos;
}
/* :24-41 */
void orc_fft64_ifft(const orc_fft64_module *md, double *data) {
    size_t m = md->m;
    double *re = data, *im = data + m;
    const double *o = md->omg_i;
    if (m <= 16) {
        if (m == 2) ifft2(re, im, o);
        else if (m == 4) ifft4(re, im, o);
        else if (m == 8) ifft8(re, im, o);
        else if (m == 16) ifft16(re, im, o);
    } else if (m <= 2048) {
        ifft_bfs16(m, re, im, o, 0);
    } else {
        ifft_rec16(m, re, im, o, 0);
    }
}

/* ---- helpers ------------------------------------------------------------------ */
static inline double *dlimb(const orc_vec_znx_dft *v, size_t col, size_t limb) {
    assert(col < v->cols && limb < v->size);
    return (double *)v->data + v->n * (limb * v->cols + col);
}
static inline int64_t *blimb(const orc_vec_znx_big *v, size_t col, size_t limb) {
    assert(col < v->cols && limb < v->size);
    return (int64_t *)v->data + v->n * (limb * v->cols + col);
}
static inline int64_t *zlimb(const orc_vec_znx *v, size_t col, size_t limb) {
    assert(col < v->cols && limb < v->size);
    return v->data + v->n * (limb * v->cols + col);
}
static inline size_t zmin(size_t a, size_t b) { return a < b ? a : b; }
static inline size_t div_ceil(size_t a, size_t b) { return (a + b - 1) / b; }

/* conversion.rs:19-28 */
static void reim_from_znx(size_t n, double *r, const int64_t *a) {
    for (size_t i = 0; i < n; i++) r[i] = (double)a[i];
}
/* Rust `as i64` saturates and maps NaN to 0 */
static inline int64_t f64_to_i64_sat(double x) {
    if (x != x) return 0;
    if (x >= 9223372036854775808.0) return INT64_MAX;
    if (x <= -9223372036854775808.0) return INT64_MIN;
    return (int64_t)x;
}
/* conversion.rs:43-60 ; f64::round = half away from zero = C round() */
static void reim_to_znx(size_t n, int64_t *r, double divisor, const double *a) {
    double inv = 1.0 / divisor;
    for (size_t i = 0; i < n; i++) r[i] = f64_to_i64_sat(round(a[i] * inv));
}

/* ---- vec_znx_dft.rs ------------------------------------------------------------ */
/* :160-200 -- NOTE: a limb with source index >= a.size inside min_steps is left untouched */
void orc_fft64_vec_znx_dft_apply(const orc_fft64_module *m, size_t step, size_t offset, orc_vec_znx_dft *res,
                                 size_t res_col, const orc_vec_znx *a, size_t a_col) {
    size_t n = res->n;
    size_t steps = div_ceil(a->size, step);
    size_t min_steps = zmin(res->size, steps);
    for (size_t j = 0; j < min_steps; j++) {
        size_t limb = offset + j * step;
        if (limb < a->size) {
            reim_from_znx(n, dlimb(res, res_col, j), zlimb(a, a_col, limb));
            orc_fft64_fft(m, dlimb(res, res_col, j));
        }
    }
    for (size_t j = min_steps; j < res->size; j++) memset(dlimb(res, res_col, j), 0, 8 * n);
}
/* :202-232 */
void orc_fft64_vec_znx_idft_apply(const orc_fft64_module *m, orc_vec_znx_big *res, size_t res_col,
                                  const orc_vec_znx_dft *a, size_t a_col) {
    size_t n = res->n, min_size = zmin(res->size, a->size);
    double *tmp = (double *)malloc(8 * n);
    for (size_t j = 0; j < min_size; j++) {
        memcpy(tmp, dlimb(a, a_col, j), 8 * n);
        orc_fft64_ifft(m, tmp);
        reim_to_znx(n, blimb(res, res_col, j), (double)m->m, tmp);
    }
    for (size_t j = min_size; j < res->size; j++) memset(blimb(res, res_col, j), 0, 8 * n);
    free(tmp);
}
/* :234-262 */
void orc_fft64_vec_znx_idft_apply_tmpa(const orc_fft64_module *m, orc_vec_znx_big *res, size_t res_col,
                                       orc_vec_znx_dft *a, size_t a_col) {
    size_t n = res->n, min_size = zmin(res->size, a->size);
    for (size_t j = 0; j < min_size; j++) {
        orc_fft64_ifft(m, dlimb(a, a_col, j));
        reim_to_znx(n, blimb(res, res_col, j), (double)m->m, dlimb(a, a_col, j));
    }
    for (size_t j = min_size; j < res->size; j++) memset(blimb(res, res_col, j), 0, 8 * n);
}
/* :264-288 */
void orc_fft64_vec_znx_idft_apply_consume(const orc_fft64_module *m, orc_vec_znx_dft *a) {
    size_t n = a->n;
    for (size_t i = 0; i < a->cols; i++)
        for (size_t j = 0; j < a->size; j++) {
            double *d = dlimb(a, i, j);
            orc_fft64_ifft(m, d);
            reim_to_znx(n, (int64_t *)d, (double)m->m, d);
        }
}

/* fft_vec.rs leaf loops */
static void v_add(size_t n, double *r, const double *a, const double *b) { for (size_t i = 0; i < n; i++) r[i] = a[i] + b[i]; }
static void v_sub(size_t n, double *r, const double *a, const double *b) { for (size_t i = 0; i < n; i++) r[i] = a[i] - b[i]; }
static void v_neg(size_t n, double *r, const double *a) { for (size_t i = 0; i < n; i++) r[i] = -a[i]; }

/* :13-66 */
void orc_fft64_vec_znx_dft_add_into(orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx_dft *a, size_t a_col,
                                    const orc_vec_znx_dft *b, size_t b_col) {
    size_t n = res->n, rs = res->size;
    const orc_vec_znx_dft *lo = a->size <= b->size ? a : b, *hi = a->size <= b->size ? b : a;
    size_t hi_col = a->size <= b->size ? b_col : a_col;
    size_t sum = zmin(lo->size, rs), cpy = zmin(hi->size, rs);
    for (size_t j = 0; j < sum; j++) v_add(n, dlimb(res, res_col, j), dlimb(a, a_col, j), dlimb(b, b_col, j));
    for (size_t j = sum; j < cpy; j++) memcpy(dlimb(res, res_col, j), dlimb(hi, hi_col, j), 8 * n);
    for (size_t j = cpy; j < rs; j++) memset(dlimb(res, res_col, j), 0, 8 * n);
}
/* :68-90 */
void orc_fft64_vec_znx_dft_add_assign(orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx_dft *a, size_t a_col) {
    size_t sum = zmin(res->size, a->size);
    for (size_t j = 0; j < sum; j++) v_add(res->n, dlimb(res, res_col, j), dlimb(res, res_col, j), dlimb(a, a_col, j));
}
/* :290-340 */
void orc_fft64_vec_znx_dft_sub(orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx_dft *a, size_t a_col,
                               const orc_vec_znx_dft *b, size_t b_col) {
    size_t n = res->n, rs = res->size;
    if (a->size <= b->size) {
        size_t sum = zmin(a->size, rs), cpy = zmin(b->size, rs);
        for (size_t j = 0; j < sum; j++) v_sub(n, dlimb(res, res_col, j), dlimb(a, a_col, j), dlimb(b, b_col, j));
        for (size_t j = sum; j < cpy; j++) v_neg(n, dlimb(res, res_col, j), dlimb(b, b_col, j));
        for (size_t j = cpy; j < rs; j++) memset(dlimb(res, res_col, j), 0, 8 * n);
    } else {
        size_t sum = zmin(b->size, rs), cpy = zmin(a->size, rs);
        for (size_t j = 0; j < sum; j++) v_sub(n, dlimb(res, res_col, j), dlimb(a, a_col, j), dlimb(b, b_col, j));
        for (size_t j = sum; j < cpy; j++) memcpy(dlimb(res, res_col, j), dlimb(a, a_col, j), 8 * n);
        for (size_t j = cpy; j < rs; j++) memset(dlimb(res, res_col, j), 0, 8 * n);
    }
}
/* :342-364 */
void orc_fft64_vec_znx_dft_sub_assign(orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx_dft *a, size_t a_col) {
    size_t sum = zmin(res->size, a->size);
    for (size_t j = 0; j < sum; j++) v_sub(res->n, dlimb(res, res_col, j), dlimb(res, res_col, j), dlimb(a, a_col, j));
}
/* :366-392 */
void orc_fft64_vec_znx_dft_sub_negate_assign(orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx_dft *a,
                                             size_t a_col) {
    size_t rs = res->size, sum = zmin(rs, a->size);
    for (size_t j = 0; j < sum; j++) v_sub(res->n, dlimb(res, res_col, j), dlimb(a, a_col, j), dlimb(res, res_col, j));
    for (size_t j = sum; j < rs; j++) v_neg(res->n, dlimb(res, res_col, j), dlimb(res, res_col, j));
}
/* :126-157 */
void orc_fft64_vec_znx_dft_copy(size_t step, size_t offset, orc_vec_znx_dft *res, size_t res_col,
                                const orc_vec_znx_dft *a, size_t a_col) {
    size_t n = res->n;
    size_t steps = div_ceil(a->size, step), min_steps = zmin(res->size, steps);
    for (size_t j = 0; j < min_steps; j++) {
        size_t limb = offset + j * step;
        if (limb < a->size)
            memcpy(dlimb(res, res_col, j), dlimb(a, a_col, limb), 8 * n);
        else
            memset(dlimb(res, res_col, j), 0, 8 * n);
    }
    for (size_t j = min_steps; j < res->size; j++) memset(dlimb(res, res_col, j), 0, 8 * n);
}
/* :394-405 */
void orc_fft64_vec_znx_dft_zero(orc_vec_znx_dft *res, size_t res_col) {
    for (size_t j = 0; j < res->size; j++) memset(dlimb(res, res_col, j), 0, 8 * res->n);
}

/* ---- svp.rs -------------------------------------------------------------------- */
/* :9-20 */
void orc_fft64_svp_prepare(const orc_fft64_module *m, orc_svp_ppol *res, size_t res_col, const orc_scalar_znx *a,
                           size_t a_col) {
    size_t n = res->n;
    double *r = (double *)res->data + n * res_col;
    reim_from_znx(n, r, a->data + n * a_col);
    orc_fft64_fft(m, r);
}
/* fft_vec.rs:150-173 (reim_mul_ref): res = a * b, complex, [re | im] halves */
static void reim_mul(size_t n, double *r, const double *a, const double *b) {
    size_t m = n / 2;
    for (size_t i = 0; i < m; i++) {
        double ar = a[i], ai = a[i + m], br = b[i], bi = b[i + m];
        double rr = ar * br - ai * bi;
        double ri = ar * bi + ai * br;
        r[i] = rr;
        r[i + m] = ri;
    }
}
/* :57-79 */
void orc_fft64_svp_apply_dft_to_dft(const orc_fft64_module *md, orc_vec_znx_dft *res, size_t res_col,
                                    const orc_svp_ppol *a, size_t a_col, const orc_vec_znx_dft *b, size_t b_col) {
    (void)md;
    size_t n = res->n, min_size = zmin(res->size, b->size);
    const double *pp = (const double *)a->data + n * a_col;
    for (size_t j = 0; j < min_size; j++) reim_mul(n, dlimb(res, res_col, j), pp, dlimb(b, b_col, j));
    for (size_t j = min_size; j < res->size; j++) memset(dlimb(res, res_col, j), 0, 8 * n);
}
/* :81-94 (reim_mul_assign(res, ppol): res = res * ppol) */
void orc_fft64_svp_apply_dft_to_dft_assign(const orc_fft64_module *md, orc_vec_znx_dft *res, size_t res_col,
                                           const orc_svp_ppol *a, size_t a_col) {
    (void)md;
    size_t n = res->n;
    const double *pp = (const double *)a->data + n * a_col;
    for (size_t j = 0; j < res->size; j++) reim_mul(n, dlimb(res, res_col, j), pp, dlimb(res, res_col, j));
}

/* ---- vmp.rs -------------------------------------------------------------------- */
/* :52-93 */
void orc_fft64_vmp_prepare(const orc_fft64_module *md, orc_vmp_pmat *res, const orc_mat_znx *a) {
    size_t n = res->n, m = md->m;
    assert(n >= 8 && a->n == n);
    size_t nrows = a->cols_in * a->rows, ncols = a->cols_out * a->size;
    size_t offset = nrows * ncols * 8;
    double *pm = (double *)res->data;
    double *tmp = (double *)malloc(8 * n);
    for (size_t row_i = 0; row_i < nrows; row_i++)
        for (size_t col_i = 0; col_i < ncols; col_i++) {
            size_t pos = n * (row_i * ncols + col_i);
            reim_from_znx(n, tmp, a->data + pos);
            orc_fft64_fft(md, tmp);
            double *dst = (col_i == ncols - 1 && ncols % 2 != 0)
                              ? pm + col_i * nrows * 8 + row_i * 8
                              : pm + (col_i / 2) * (nrows * 16) + row_i * 16 + (col_i % 2) * 8;
            for (size_t blk = 0; blk < (m >> 2); blk++) { /* reim4_extract_1blk (rows=1): re chunk then im chunk */
                memcpy(dst + blk * offset, tmp + 4 * blk, 32);
                memcpy(dst + blk * offset + 4, tmp + m + 4 * blk, 32);
            }
        }
    free(tmp);
}

/* reim4/arithmetic_ref.rs:223-232 */
static inline void reim4_add_mul(double *dst, const double *a, const double *b) {
    for (int k = 0; k < 4; k++) {
        double ar = a[k], br = b[k], ai = a[k + 4], bi = b[k + 4];
        dst[k] += ar * br - ai * bi;
        dst[k + 4] += ar * bi + ai * br;
    }
}
/* :138-159 */
static void mat1col(size_t nrows, double *dst, const double *u, const double *v) {
    double acc[8] = {0};
    for (size_t i = 0; i < nrows; i++) reim4_add_mul(acc, u + 8 * i, v + 8 * i);
    memcpy(dst, acc, 64);
}
/* :161-186 */
static void mat2cols(size_t nrows, double *dst, const double *u, const double *v) {
    double a0[8] = {0}, a1[8] = {0};
    for (size_t i = 0; i < nrows; i++) {
        reim4_add_mul(a0, u + 8 * i, v + 16 * i);
        reim4_add_mul(a1, u + 8 * i, v + 16 * i + 8);
    }
    memcpy(dst, a0, 64);
    memcpy(dst + 8, a1, 64);
}
/* :188-221 */
static void mat2cols_2nd(size_t nrows, double *dst, const double *u, const double *v) {
    double acc[8] = {0};
    for (size_t i = 0; i < nrows; i++) reim4_add_mul(acc, u + 8 * i, v + 16 * i + 8);
    memcpy(dst, acc, 64);
}
/* :53-72 (OVERWRITE) : one output poly, re chunk at blk*4, im chunk at m + blk*4 */
static void save_1blk(size_t m, size_t blk, double *dst, const double *src) {
    memcpy(dst + 4 * blk, src, 32);
    memcpy(dst + m + 4 * blk, src + 4, 32);
}
/* :74-135 : two consecutive output polys */
static void save_2blk(size_t m, size_t blk, double *dst, const double *src) {
    save_1blk(m, blk, dst, src);
    save_1blk(m, blk, dst + 2 * m, src + 8);
}

/* :186-264 (OVERWRITE = true) */
static void vmp_core(size_t n, double *res, size_t res_size, const double *a, size_t a_size, const double *pmat,
                     size_t limb_offset, size_t nrows, size_t ncols) {
    size_t m = n >> 1;
    size_t row_max = zmin(nrows, a_size), col_max = zmin(ncols, res_size);
    if (limb_offset >= col_max) {
        memset(res, 0, 8 * n * res_size);
        return;
    }
    double out[16];
    double *ext = (double *)malloc(64 * (row_max ? row_max : 1));
    for (size_t blk = 0; blk < (m >> 2); blk++) {
        const double *mb = pmat + blk * (8 * nrows * ncols);
        /* reim4_extract_1blk_from_reim_contiguous: 2*rows chunks of 4 spaced by m */
        for (size_t c = 0; c < 2 * row_max; c++) memcpy(ext + 4 * c, a + c * m + 4 * blk, 32);
        if (limb_offset % 2 == 0) {
            size_t col_res = 0;
            for (size_t col_pmat = limb_offset; col_pmat + 1 < col_max; col_pmat += 2, col_res += 2) {
                mat2cols(row_max, out, ext, mb + col_pmat * (8 * nrows));
                save_2blk(m, blk, res + col_res * n, out);
            }
        } else {
            mat2cols_2nd(row_max, out, ext, mb + (limb_offset - 1) * (8 * nrows));
            save_1blk(m, blk, res, out);
            size_t col_res = 1;
            for (size_t col_pmat = limb_offset + 1; col_pmat + 1 < col_max; col_pmat += 2, col_res += 2) {
                mat2cols(row_max, out, ext, mb + col_pmat * (8 * nrows));
                save_2blk(m, blk, res + col_res * n, out);
            }
        }
        if (col_max % 2 != 0) {
            size_t last = col_max - 1;
            if (last >= limb_offset) {
                if (ncols == col_max)
                    mat1col(row_max, out, ext, mb + last * (8 * nrows));
                else
                    mat2cols(row_max, out, ext, mb + last * (8 * nrows));
                save_1blk(m, blk, res + (last - limb_offset) * n, out);
            }
        }
    }
    memset(res + col_max * n, 0, 8 * n * (res_size - col_max));
    free(ext);
}
/* :144-183 */
void orc_fft64_vmp_apply_dft_to_dft(const orc_fft64_module *md, orc_vec_znx_dft *res, const orc_vec_znx_dft *a,
                                    const orc_vmp_pmat *pmat, size_t limb_offset) {
    (void)md;
    size_t n = res->n;
    size_t nrows = pmat->cols_in * pmat->rows, ncols = pmat->cols_out * pmat->size;
    vmp_core(n, (double *)res->data, res->cols * res->size, (const double *)a->data, a->cols * a->size,
             (const double *)pmat->data, limb_offset * pmat->cols_out, nrows, ncols);
}

/* ---- vec_znx_big.rs (i64 big == VecZnx) ------------------------------------------ */
void orc_fft64_vec_znx_big_add_small_assign(orc_vec_znx_big *res, size_t res_col, const orc_vec_znx *a, size_t a_col) {
    size_t sum = zmin(res->size, a->size);
    for (size_t j = 0; j < sum; j++) {
        int64_t *r = blimb(res, res_col, j);
        const int64_t *x = zlimb(a, a_col, j);
        for (size_t i = 0; i < res->n; i++) r[i] = (int64_t)((uint64_t)r[i] + (uint64_t)x[i]);
    }
}
/* :241-278 */
void orc_fft64_vec_znx_big_normalize(orc_vec_znx *res, size_t res_base2k, int64_t res_offset, size_t res_col,
                                     const orc_vec_znx_big *a, size_t a_base2k, size_t a_col, int op) {
    orc_vec_znx av = {(int64_t *)a->data, a->n, a->cols, a->size};
    orc_vec_znx_normalize(res, res_base2k, res_offset, res_col, &av, a_base2k, a_col, op);
}

/* ---- reference/fft64/convolution.rs (bivariate convolution, FFT64) -----------------------------------------------------
 * CnvPVecL / CnvPVecR are opaque prepared layouts (the reference interleaves reim4 blocks, convolution.rs:61-70); this
 * restatement keeps them in the VecZnxDft layout.  Per frequency the arithmetic and its order are the reference's:
 * reim4_convolution_1coeff_ref (reim4/arithmetic_ref.rs:235-247) = sum over j ascending of reim4_add_mul. */
static void reim_from_znx_masked(size_t n, double *r, const int64_t *a, int64_t mask) {
    for (size_t i = 0; i < n; i++) r[i] = (double)(a[i] & mask);
}
/* convolution.rs:33-73 */
void orc_fft64_cnv_prepare(const orc_fft64_module *m, orc_vec_znx_dft *res, const orc_vec_znx *a, int64_t mask) {
    size_t n = res->n, min_size = zmin(res->size, a->size);
    for (size_t col = 0; col < res->cols; col++) {
        for (size_t j = 0; j < min_size; j++) {
            double *r = dlimb(res, col, j);
            if (j + 1 == min_size) reim_from_znx_masked(n, r, zlimb(a, col, j), mask);
            else reim_from_znx(n, r, zlimb(a, col, j));
            orc_fft64_fft(m, r);
        }
        for (size_t j = min_size; j < res->size; j++) memset(dlimb(res, col, j), 0, 8 * n);
    }
}
/* convolution.rs:199-249 (apply_dft), :256-334 (pairwise: operands summed first with reim_add) */
static void fcnv_apply_core(size_t cnv_offset, orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx_dft *a, size_t a_i, size_t a_j,
                            const orc_vec_znx_dft *b, size_t b_i, size_t b_j) {
    size_t n = res->n, mh = n / 2, res_size = res->size, a_size = a->size, b_size = b->size;
    if (a_size == 0 || b_size == 0) {
        for (size_t j = 0; j < res_size; j++) memset(dlimb(res, res_col, j), 0, 8 * n);
        return;
    }
    size_t bound = a_size + b_size - 1, min_size = zmin(res_size, bound), offset = zmin(cnv_offset, bound);
    double *as = (double *)malloc(8 * n * a_size), *bs = (double *)malloc(8 * n * b_size);
    for (size_t j = 0; j < a_size; j++)
        for (size_t i = 0; i < n; i++) as[j * n + i] = a_i == a_j ? dlimb(a, a_i, j)[i] : dlimb(a, a_i, j)[i] + dlimb(a, a_j, j)[i];
    for (size_t j = 0; j < b_size; j++)
        for (size_t i = 0; i < n; i++) bs[j * n + i] = b_i == b_j ? dlimb(b, b_i, j)[i] : dlimb(b, b_i, j)[i] + dlimb(b, b_j, j)[i];
    for (size_t k = 0; k < min_size; k++) {
        size_t k_abs = k + offset;
        double *r = dlimb(res, res_col, k);
        memset(r, 0, 8 * n);
        if (k_abs >= a_size + b_size) continue;
        size_t j_min = k_abs > a_size - 1 ? k_abs - (a_size - 1) : 0, j_max = zmin(k_abs + 1, b_size);
        for (size_t j = j_min; j < j_max; j++) {
            const double *x = as + (k_abs - j) * n, *y = bs + j * n;
            for (size_t f = 0; f < mh; f++) { /* reim4_add_mul, reim4/arithmetic_ref.rs:223-232 */
                double ar = x[f], ai = x[f + mh], br = y[f], bi = y[f + mh];
                r[f] += ar * br - ai * bi;
                r[f + mh] += ar * bi + ai * br;
            }
        }
    }
    free(as);
    free(bs);
    for (size_t j = min_size; j < res_size; j++) memset(dlimb(res, res_col, j), 0, 8 * n);
}
void orc_fft64_cnv_apply_dft(size_t cnv_offset, orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx_dft *a, size_t a_col,
                             const orc_vec_znx_dft *b, size_t b_col) {
    fcnv_apply_core(cnv_offset, res, res_col, a, a_col, a_col, b, b_col, b_col);
}
void orc_fft64_cnv_pairwise_apply_dft(size_t cnv_offset, orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx_dft *a,
                                      const orc_vec_znx_dft *b, size_t col_i, size_t col_j) {
    fcnv_apply_core(cnv_offset, res, res_col, a, col_i, col_j, b, col_i, col_j);
}
/* convolution.rs:144-191 with i64_convolution_by_const_1coeff_ref (:401-426): wrapping i64 */
void orc_fft64_cnv_by_const_apply(size_t cnv_offset, orc_vec_znx_big *res, size_t res_col, const orc_vec_znx *a, size_t a_col,
                                  const int64_t *b, size_t b_size) {
    size_t n = res->n, res_size = res->size, a_size = a->size;
    size_t bound = a_size + b_size - 1, min_size = zmin(res_size, bound), offset = zmin(cnv_offset, bound);
    for (size_t k = 0; k < min_size; k++) {
        size_t k_abs = k + offset;
        int64_t *r = blimb(res, res_col, k);
        memset(r, 0, 8 * n);
        if (k_abs >= a_size + b_size) continue;
        size_t j_min = k_abs > a_size - 1 ? k_abs - (a_size - 1) : 0, j_max = zmin(k_abs + 1, b_size);
        for (size_t j = j_min; j < j_max; j++) {
            const int64_t *x = zlimb(a, a_col, k_abs - j);
            for (size_t i = 0; i < n; i++) r[i] = (int64_t)((uint64_t)r[i] + (uint64_t)x[i] * (uint64_t)b[j]);
        }
    }
    for (size_t j = min_size; j < res_size; j++) memset(blimb(res, res_col, j), 0, 8 * n);
}
