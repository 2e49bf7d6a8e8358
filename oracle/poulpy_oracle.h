/*
 * poulpy_oracle.h -- CPU restatement of the poulpy-cpu-ref hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is the parity checker for the B200 backend
 * and the "port" CPU baseline of bench.py.  Nothing under poulpy_b200/ may
 * include, link or call it; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py do.
 *
 * The reference (phantomzone-org/poulpy v0.5.0) is pure Rust on a pinned
 * nightly; no cargo/rustc exists in this image, so the reference itself cannot
 * be compiled (no oracle/_ref).  Every function below cites the reference
 * file:line it restates (paths relative to the reference root).  Parity is
 * pinned by the reference's own known-answer tests restated in
 * tests/test_oracle_kat.py (ntt_convolution, ntt_intt_identity, Primes30
 * constants, normalize torus-value property, FFT64<->NTT120 cross-backend
 * equality) -- see DESIGN.md "Oracle".
 *
 * Layouts are the reference's: limb-major, column-minor; limb j of column i
 * starts at scalar offset n*(j*cols+i)  (poulpy-hal/src/layouts/znx_base.rs:57-84).
 */
#ifndef POULPY_ORACLE_H
#define POULPY_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef __int128 i128;
typedef unsigned __int128 u128;

/* poulpy-hal/src/layouts/vec_znx.rs:33-41 (i64 scalars) */
typedef struct { int64_t *data; size_t n, cols, size; } orc_vec_znx;
/* poulpy-hal/src/layouts/vec_znx_dft.rs:25-34 ; ScalarPrep = Q120bScalar (4 x u64) or f64 */
typedef struct { void *data; size_t n, cols, size; } orc_vec_znx_dft;
/* poulpy-hal/src/layouts/vec_znx_big.rs:25 ; ScalarBig = i128 or i64 */
typedef struct { void *data; size_t n, cols, size; } orc_vec_znx_big;
/* poulpy-hal/src/layouts/svp_ppol.rs:23 */
typedef struct { void *data; size_t n, cols; } orc_svp_ppol;
/* poulpy-hal/src/layouts/scalar_znx.rs */
typedef struct { int64_t *data; size_t n, cols; } orc_scalar_znx;
/* poulpy-hal/src/layouts/vmp_pmat.rs:25 */
typedef struct { void *data; size_t n, rows, cols_in, cols_out, size; } orc_vmp_pmat;
/* poulpy-hal/src/layouts/mat_znx.rs:28 */
typedef struct { int64_t *data; size_t n, rows, cols_in, cols_out, size; } orc_mat_znx;

/* ------------------------------------------------------------------ NTT120 */
/* poulpy-cpu-ref/src/reference/ntt120/primes.rs:80-90 */
extern const uint32_t ORC_Q[4];
extern const uint32_t ORC_OMEGA[4];
extern const uint32_t ORC_CRT_CST[4];

typedef struct orc_ntt120_module orc_ntt120_module;
orc_ntt120_module *orc_ntt120_new(size_t n);       /* hal_defaults/module.rs:25-33 */
void orc_ntt120_free(orc_ntt120_module *m);
size_t orc_ntt120_n(const orc_ntt120_module *m);
/* table introspection for the KATs (level bit sizes etc.) */
size_t orc_ntt120_fwd_levels(const orc_ntt120_module *m, uint64_t *bs, int *reduce, size_t cap);
size_t orc_ntt120_inv_levels(const orc_ntt120_module *m, uint64_t *bs, int *reduce, size_t cap);
uint64_t orc_ntt120_reduc_h(const orc_ntt120_module *m);
uint64_t orc_ntt120_bbc_h(const orc_ntt120_module *m);

void orc_ntt120_b_from_znx64(size_t nn, uint64_t *res, const int64_t *x);   /* arithmetic.rs:39-60 */
void orc_ntt120_c_from_b(size_t nn, uint32_t *res, const uint64_t *x);      /* arithmetic.rs:202-214 */
void orc_ntt120_b_to_znx128(size_t nn, i128 *res, const uint64_t *x);       /* arithmetic.rs:119-140 */
void orc_ntt120_ntt(const orc_ntt120_module *m, uint64_t *data);            /* ntt.rs:558-602 */
void orc_ntt120_intt(const orc_ntt120_module *m, uint64_t *data);           /* ntt.rs:617-684 */

/* vec_znx_dft.rs */
void orc_ntt120_vec_znx_dft_apply(const orc_ntt120_module *m, size_t step, size_t offset,
                                  orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx *a, size_t a_col);
void orc_ntt120_vec_znx_idft_apply(const orc_ntt120_module *m, orc_vec_znx_big *res, size_t res_col,
                                   const orc_vec_znx_dft *a, size_t a_col);
void orc_ntt120_vec_znx_idft_apply_tmpa(const orc_ntt120_module *m, orc_vec_znx_big *res, size_t res_col,
                                        orc_vec_znx_dft *a, size_t a_col);
/* in place: on return a->data holds a VecZnxBig(n, cols, size) of i128 in its first half */
void orc_ntt120_vec_znx_idft_apply_consume(const orc_ntt120_module *m, orc_vec_znx_dft *a);
void orc_ntt120_vec_znx_dft_add_into(orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx_dft *a, size_t a_col,
                                     const orc_vec_znx_dft *b, size_t b_col);
void orc_ntt120_vec_znx_dft_add_assign(orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx_dft *a, size_t a_col);
void orc_ntt120_vec_znx_dft_add_scaled_assign(orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx_dft *a, size_t a_col, int64_t a_scale);
void orc_ntt120_vec_znx_dft_sub(orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx_dft *a, size_t a_col,
                                const orc_vec_znx_dft *b, size_t b_col);
void orc_ntt120_vec_znx_dft_sub_assign(orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx_dft *a, size_t a_col);
void orc_ntt120_vec_znx_dft_sub_negate_assign(orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx_dft *a, size_t a_col);
void orc_ntt120_vec_znx_dft_copy(size_t step, size_t offset, orc_vec_znx_dft *res, size_t res_col,
                                 const orc_vec_znx_dft *a, size_t a_col);
void orc_ntt120_vec_znx_dft_zero(orc_vec_znx_dft *res, size_t res_col);

/* svp.rs */
void orc_ntt120_svp_prepare(const orc_ntt120_module *m, orc_svp_ppol *res, size_t res_col,
                            const orc_scalar_znx *a, size_t a_col);
void orc_ntt120_svp_apply_dft_to_dft(const orc_ntt120_module *m, orc_vec_znx_dft *res, size_t res_col,
                                     const orc_svp_ppol *a, size_t a_col, const orc_vec_znx_dft *b, size_t b_col);
void orc_ntt120_svp_apply_dft_to_dft_assign(const orc_ntt120_module *m, orc_vec_znx_dft *res, size_t res_col,
                                            const orc_svp_ppol *a, size_t a_col);

/* vmp.rs */
void orc_ntt120_vmp_prepare(const orc_ntt120_module *m, orc_vmp_pmat *res, const orc_mat_znx *a);
void orc_ntt120_vmp_apply_dft_to_dft(const orc_ntt120_module *m, orc_vec_znx_dft *res, const orc_vec_znx_dft *a,
                                     const orc_vmp_pmat *pmat, size_t limb_offset);

/* vec_znx_big.rs (i128) */
void orc_ntt120_vec_znx_big_add_small_assign(orc_vec_znx_big *res, size_t res_col, const orc_vec_znx *a, size_t a_col);
void orc_ntt120_vec_znx_big_from_small(orc_vec_znx_big *res, size_t res_col, const orc_vec_znx *a, size_t a_col);
void orc_ntt120_vec_znx_big_add_assign(orc_vec_znx_big *res, size_t res_col, const orc_vec_znx_big *a, size_t a_col);
void orc_ntt120_vec_znx_big_sub_assign(orc_vec_znx_big *res, size_t res_col, const orc_vec_znx_big *a, size_t a_col);
void orc_ntt120_vec_znx_big_negate_assign(orc_vec_znx_big *res, size_t res_col);
/* op: 0 = overwrite, +1 = add-assign, -1 = sub-assign */
void orc_ntt120_vec_znx_big_normalize(orc_vec_znx *res, size_t res_base2k, int64_t res_offset, size_t res_col,
                                      const orc_vec_znx_big *a, size_t a_base2k, size_t a_col, int op);

/* ------------------------------------------------------------------- FFT64 */
typedef struct orc_fft64_module orc_fft64_module;
orc_fft64_module *orc_fft64_new(size_t n);          /* poulpy-cpu-ref/src/fft64/module.rs:62-69 */
void orc_fft64_free(orc_fft64_module *m);
const double *orc_fft64_omg(const orc_fft64_module *m, int inverse, size_t *len);
void orc_fft64_fft(const orc_fft64_module *m, double *data);    /* reim/fft_ref.rs:25-43 */
void orc_fft64_ifft(const orc_fft64_module *m, double *data);   /* reim/ifft_ref.rs:24- */

void orc_fft64_vec_znx_dft_apply(const orc_fft64_module *m, size_t step, size_t offset,
                                 orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx *a, size_t a_col);
void orc_fft64_vec_znx_idft_apply(const orc_fft64_module *m, orc_vec_znx_big *res, size_t res_col,
                                  const orc_vec_znx_dft *a, size_t a_col);
void orc_fft64_vec_znx_idft_apply_tmpa(const orc_fft64_module *m, orc_vec_znx_big *res, size_t res_col,
                                       orc_vec_znx_dft *a, size_t a_col);
void orc_fft64_vec_znx_idft_apply_consume(const orc_fft64_module *m, orc_vec_znx_dft *a);
void orc_fft64_vec_znx_dft_add_into(orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx_dft *a, size_t a_col,
                                    const orc_vec_znx_dft *b, size_t b_col);
void orc_fft64_vec_znx_dft_add_assign(orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx_dft *a, size_t a_col);
void orc_fft64_vec_znx_dft_sub(orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx_dft *a, size_t a_col,
                               const orc_vec_znx_dft *b, size_t b_col);
void orc_fft64_vec_znx_dft_sub_assign(orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx_dft *a, size_t a_col);
void orc_fft64_vec_znx_dft_sub_negate_assign(orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx_dft *a, size_t a_col);
void orc_fft64_vec_znx_dft_copy(size_t step, size_t offset, orc_vec_znx_dft *res, size_t res_col,
                                const orc_vec_znx_dft *a, size_t a_col);
void orc_fft64_vec_znx_dft_zero(orc_vec_znx_dft *res, size_t res_col);
void orc_fft64_svp_prepare(const orc_fft64_module *m, orc_svp_ppol *res, size_t res_col,
                           const orc_scalar_znx *a, size_t a_col);
void orc_fft64_svp_apply_dft_to_dft(const orc_fft64_module *m, orc_vec_znx_dft *res, size_t res_col,
                                    const orc_svp_ppol *a, size_t a_col, const orc_vec_znx_dft *b, size_t b_col);
void orc_fft64_svp_apply_dft_to_dft_assign(const orc_fft64_module *m, orc_vec_znx_dft *res, size_t res_col,
                                           const orc_svp_ppol *a, size_t a_col);
void orc_fft64_vmp_prepare(const orc_fft64_module *m, orc_vmp_pmat *res, const orc_mat_znx *a);
void orc_fft64_vmp_apply_dft_to_dft(const orc_fft64_module *m, orc_vec_znx_dft *res, const orc_vec_znx_dft *a,
                                    const orc_vmp_pmat *pmat, size_t limb_offset);
void orc_fft64_vec_znx_big_add_small_assign(orc_vec_znx_big *res, size_t res_col, const orc_vec_znx *a, size_t a_col);
void orc_fft64_vec_znx_big_normalize(orc_vec_znx *res, size_t res_base2k, int64_t res_offset, size_t res_col,
                                     const orc_vec_znx_big *a, size_t a_base2k, size_t a_col, int op);

/* ------------------------------------------------- coefficient-domain (znx) */
/* reference/vec_znx/normalize.rs:18-50 ; op as above */
void orc_vec_znx_normalize(orc_vec_znx *res, size_t res_base2k, int64_t res_offset, size_t res_col,
                           const orc_vec_znx *a, size_t a_base2k, size_t a_col, int op);
/* reference/vec_znx/rotate.rs, znx/rotate.rs:3-26 */
void orc_vec_znx_rotate(int64_t p, orc_vec_znx *res, size_t res_col, const orc_vec_znx *a, size_t a_col);
/* reference/vec_znx/add.rs:60-82, mul_xp_minus_one.rs:24-38, normalize.rs:403-425 */
void orc_vec_znx_add_assign(orc_vec_znx *res, size_t res_col, const orc_vec_znx *a, size_t a_col);
void orc_vec_znx_mul_xp_minus_one_assign(int64_t p, orc_vec_znx *res, size_t res_col);
void orc_vec_znx_normalize_assign(size_t base2k, orc_vec_znx *res, size_t res_col);
/* reference/znx/automorphism.rs:1-17, reference/vec_znx/automorphism.rs:9-38 */
void orc_znx_automorphism(int64_t p, int64_t *res, const int64_t *a, size_t n);
void orc_vec_znx_automorphism(int64_t p, orc_vec_znx *res, size_t res_col, const orc_vec_znx *a, size_t a_col);
void orc_znx_rotate(int64_t p, int64_t *res, const int64_t *a, size_t n);
/* reference/vec_znx/shift.rs:186-243 */
void orc_vec_znx_rsh_assign(size_t base2k, size_t k, orc_vec_znx *res, size_t res_col);

/* ------------------------------------------------------------ compositions */
/* flavour: 0 = NTT120, 1 = FFT64.  `mod` is the matching module pointer. */
/* poulpy-core/src/keyswitching/glwe.rs:53-109 (dsize = 1 and dsize > 1) */
void orc_glwe_keyswitch(int flavour, const void *mod, orc_vec_znx *res, size_t res_base2k,
                        const orc_vec_znx *a, size_t a_base2k, const orc_vmp_pmat *key, size_t key_base2k,
                        size_t dsize);
/* poulpy-core/src/external_product/glwe.rs:99-141 */
void orc_glwe_external_product(int flavour, const void *mod, orc_vec_znx *res, size_t res_base2k,
                               const orc_vec_znx *a, size_t a_base2k, const orc_vmp_pmat *ggsw, size_t ggsw_base2k,
                               size_t dsize);

/* poulpy-core/src/automorphism/glwe_ct.rs:51-72 */
void orc_glwe_automorphism(int flavour, const void *mod, orc_vec_znx *res, size_t res_base2k, const orc_vec_znx *a, size_t a_base2k,
                           const orc_vmp_pmat *key, size_t key_base2k, int64_t p, size_t dsize);

/* poulpy-core/src/automorphism/glwe_ct.rs:142-183; poulpy-core/src/glwe_trace.rs:34-44, :129-175 */
void orc_glwe_automorphism_add_assign(int flavour, const void *mod, orc_vec_znx *res, size_t res_base2k, const orc_vmp_pmat *key,
                                      size_t key_base2k, int64_t p, size_t dsize);
/* automorphism/glwe_ct.rs:95-275: op 0 add, 1 sub, 2 sub_negate; res may alias a */
void orc_glwe_automorphism_op(int flavour, const void *mod, int op, orc_vec_znx *res, size_t res_base2k, const orc_vec_znx *a,
                              const orc_vmp_pmat *key, size_t key_base2k, int64_t p, size_t dsize);
int64_t orc_trace_galois_element(size_t i, size_t n);
void orc_glwe_trace_assign(int flavour, const void *mod, orc_vec_znx *res, size_t res_base2k, size_t skip, const orc_vmp_pmat *const *keys,
                           size_t key_base2k, size_t dsize);

/* poulpy-core/src/conversion/gglwe_to_ggsw.rs:116-268 */
void orc_ggsw_expand_row(int flavour, const void *mod, int64_t *ggsw, size_t n, size_t dnum, size_t rank, size_t size, size_t res_base2k,
                         const orc_vmp_pmat *const *tsk, size_t tsk_base2k, size_t dsize);

/* ------------------------------------------------------------ bivariate convolution (HalImpl::cnv_*, hal_impl.rs:670-754) */
/* CnvPVecL / CnvPVecR are opaque prepared layouts: this restatement keeps both in the VecZnxDft layout.
 * reference/ntt120/convolution.rs:66-236 (prepare_left / right / self), :256-335 (apply_dft), :441-557 (pairwise), :361-410 (by_const);
 * reference/fft64/convolution.rs:13-137, :199-249, :256-334, :144-191 */
void orc_ntt120_cnv_prepare(const orc_ntt120_module *m, orc_vec_znx_dft *res, const orc_vec_znx *a, int64_t mask);
void orc_ntt120_cnv_apply_dft(const orc_ntt120_module *m, size_t cnv_offset, orc_vec_znx_dft *res, size_t res_col,
                              const orc_vec_znx_dft *a, size_t a_col, const orc_vec_znx_dft *b, size_t b_col);
void orc_ntt120_cnv_pairwise_apply_dft(const orc_ntt120_module *m, size_t cnv_offset, orc_vec_znx_dft *res, size_t res_col,
                                       const orc_vec_znx_dft *a, const orc_vec_znx_dft *b, size_t col_i, size_t col_j);
void orc_ntt120_cnv_by_const_apply(size_t cnv_offset, orc_vec_znx_big *res, size_t res_col, const orc_vec_znx *a, size_t a_col,
                                   const int64_t *b, size_t b_size);
void orc_fft64_cnv_prepare(const orc_fft64_module *m, orc_vec_znx_dft *res, const orc_vec_znx *a, int64_t mask);
void orc_fft64_cnv_apply_dft(size_t cnv_offset, orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx_dft *a, size_t a_col,
                             const orc_vec_znx_dft *b, size_t b_col);
void orc_fft64_cnv_pairwise_apply_dft(size_t cnv_offset, orc_vec_znx_dft *res, size_t res_col, const orc_vec_znx_dft *a,
                                      const orc_vec_znx_dft *b, size_t col_i, size_t col_j);
void orc_fft64_cnv_by_const_apply(size_t cnv_offset, orc_vec_znx_big *res, size_t res_col, const orc_vec_znx *a, size_t a_col,
                                  const int64_t *b, size_t b_size);
/* poulpy-core/src/operations/glwe.rs:699-818 (glwe_tensor_apply) and :545-610 (glwe_tensor_relinearize): the two halves of
 * poulpy-ckks ckks_mul_into (poulpy-ckks/src/leveled/default/mul.rs:49-86).  res of tensor_apply / a of relinearize is the
 * GLWETensor VecZnx with (rank+1)(rank+2)/2 columns. */
void orc_glwe_tensor_apply(int flavour, const void *mod, size_t cnv_offset, orc_vec_znx *res, size_t res_base2k,
                           const orc_vec_znx *a, size_t a_effective_k, const orc_vec_znx *b, size_t b_effective_k, size_t ab_base2k);
void orc_glwe_tensor_relinearize(int flavour, const void *mod, orc_vec_znx *res, size_t res_base2k, const orc_vec_znx *a,
                                 size_t a_base2k, const orc_vmp_pmat *tsk, size_t key_base2k, size_t dsize);

/* ------------------------------------------------------- CGGI blind rotation */
/* poulpy-bin-fhe/src/blind_rotation/algorithms/mod.rs:136-176 ; rot_left != 0 negates (LookUpTableRotationDirection::Left).
 * lwe = VecZnx(n = n_lwe + 1, cols = 1, size); res has n_lwe + 1 entries (b, a_0, ...). */
void orc_mod_switch_2n(size_t two_n_domain, int64_t *res, const orc_vec_znx *lwe, size_t lwe_base2k, int rot_left);
/* poulpy-bin-fhe/src/blind_rotation/algorithms/cggi/key_prepared.rs:66-75 + utils.rs:6-41:
 * x_pow_a[i] = svp_prepare(X^i) for i in [0, 2n) (X^i = -X^(i-n) for i >= n); res = SvpPPol with 2n columns. */
void orc_cggi_x_pow_a(int flavour, const void *mod, orc_svp_ppol *res);
/* poulpy-bin-fhe/src/blind_rotation/algorithms/cggi/algorithm.rs:275-368 (execute_block_binary).
 * lwe_2n: mod-switched LWE (b, a_0..a_{n_lwe-1}); lut: VecZnx(1 col); brk: n_lwe prepared GGSW
 * (rows = dnum, cols_in = cols_out = rank + 1); x_pow_a as above; res = GLWE VecZnx(rank + 1, size). */
void orc_cggi_blind_rotate_block_binary(int flavour, const void *mod, orc_vec_znx *res, const int64_t *lwe_2n,
                                        size_t n_lwe, const orc_vec_znx *lut, const orc_vmp_pmat *brk,
                                        const orc_svp_ppol *x_pow_a, size_t block_size, size_t base2k);
/* algorithm.rs:121-273: lut = `ext` VecZnx(1 col) (LookupTable.data), lwe_2n mod-switched to 2 * n * ext */
void orc_cggi_blind_rotate_block_binary_extended(int flavour, const void *mod, orc_vec_znx *res, const int64_t *lwe_2n, size_t n_lwe,
                                                 const orc_vec_znx *lut, size_t ext, const orc_vmp_pmat *brk, const orc_svp_ppol *x_pow_a,
                                                 size_t block_size, size_t base2k);
/* algorithm.rs:370-443 (execute_standard): brk = n_lwe prepared GGSWs (block_size == 1 keys) */
void orc_cggi_blind_rotate_standard(int flavour, const void *mod, orc_vec_znx *res, size_t res_base2k, const int64_t *lwe_2n,
                                    size_t n_lwe, const orc_vec_znx *lut, const orc_vmp_pmat *brk, size_t brk_base2k);

/* ------------------------------------------------------------ batched drivers */
/* CPU-baseline helpers: run `batch` independent key-switches / external products with OpenMP over
 * ciphertexts (threads <= 0: all cores).  Inputs/outputs are `batch` consecutive VecZnx buffers. */
void orc_glwe_keyswitch_batch(int flavour, const void *mod, int64_t *res, size_t res_size, size_t res_base2k,
                              const int64_t *a, size_t a_size, size_t a_base2k, size_t n, size_t rank_in,
                              size_t rank_out, const orc_vmp_pmat *key, size_t key_base2k, size_t dsize,
                              size_t batch, int threads);
void orc_glwe_external_product_batch(int flavour, const void *mod, int64_t *res, size_t res_size, size_t res_base2k,
                                     const int64_t *a, size_t a_size, size_t a_base2k, size_t n, size_t rank,
                                     const orc_vmp_pmat *ggsw, size_t ggsw_base2k, size_t dsize, size_t batch,
                                     int threads);
int orc_num_threads(void);

#ifdef __cplusplus
}
#endif
/* batch.c: the block-binary blind rotation over `batch` mod-switched LWEs [batch][n_lwe + 1], ciphertexts spread over host threads */
void orc_cggi_blind_rotate_block_binary_batch(int flavour, const void *mod, int64_t *res, size_t n, size_t cols, size_t res_size,
                                              const int64_t *lwe_2n, size_t n_lwe, const orc_vec_znx *lut, const orc_vmp_pmat *brk,
                                              const orc_svp_ppol *x_pow_a, size_t block_size, size_t base2k, size_t batch, int threads);
/* ntt120.c: 1 = "cpu-avx-style" data path (four primes per __m256i, poulpy-cpu-avx/src/ntt120/ntt.rs:81-110, mat_vec_avx.rs) for the NTT
 * butterflies and the bbc products; 0 (default) = scalar restatement of poulpy-cpu-ref.  Identical lazy values either way. */
void orc_ntt120_set_simd(int on);
int orc_ntt120_get_simd(void);
#endif
