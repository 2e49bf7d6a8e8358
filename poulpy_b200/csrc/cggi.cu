// cggi.cu -- CGGI blind rotation (block-binary), batched and device resident (C3).
// Restates poulpy-bin-fhe/src/blind_rotation/algorithms/cggi/algorithm.rs:275-368 over the batched HAL kernels.
#include "internal.h"

static const uint64_t ALIGN = 256;
static inline uint64_t align_up(uint64_t x) { return (x + ALIGN - 1) / ALIGN * ALIGN; }

// x_pow_a[i] = svp_prepare(X^i), i in [0, 2n)  (cggi/key_prepared.rs:66-75, utils.rs:6-41 with y = 0)
__global__ void xpow_fill_kernel(long long *buf, uint32_t n) {
    const uint32_t ai = blockIdx.x; // [0, 2n)
    long long *p = buf + (size_t)ai * n;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        long long v = 0;
        if (ai < n) v = (i == ai) ? 1 : 0;
        else v = (i == ((ai - n) & (n - 1))) ? -1 : 0;
        p[i] = v;
    }
}
extern "C" int pgb_cggi_x_pow_a(pgb_module *m, pgb_svp_ppol *res) {
    PGB_REQUIRE(res->n == m->n && res->cols == 2 * m->n, "cggi_x_pow_a: res must be an SvpPPol with 2n columns");
    const uint64_t n = m->n, pb = prep_bytes(m);
    long long *buf = nullptr;
    PGB_CHECK_CUDA(cudaMalloc(&buf, 2 * n * n * 8));
    xpow_fill_kernel<<<(unsigned)(2 * n), 256, 0, m->stream>>>(buf, (uint32_t)n);
    m->launches++;
    LimbSet in = {(char *)buf, n * 8, 0}, out = {(char *)res->data, n * pb, 0};
    int s = m->flavour == PGB_NTT120 ? ntt120_forward(m, in, out, (int)(2 * n), 1) : fft64_forward(m, in, out, (int)(2 * n), 1);
    cudaStreamSynchronize(m->stream);
    cudaFree(buf);
    return s;
}

extern "C" size_t pgb_cggi_blind_rotate_tmp_bytes(const pgb_module *m, uint64_t rank, uint64_t res_size, uint64_t dnum, uint64_t brk_size,
                                                  uint64_t batch) {
    (void)res_size;
    const uint64_t n = m->n, pb = prep_bytes(m), bb = big_bytes(m), cols = rank + 1;
    uint64_t t = 0;
    t += align_up(batch * n * cols * dnum * pb);     // acc_dft
    t += align_up(batch * n * cols * brk_size * pb); // vmp_res
    t += align_up(batch * n * cols * brk_size * pb); // acc_add_dft
    t += align_up(batch * n * brk_size * pb);        // vmp_xai
    t += align_up(batch * n * brk_size * bb);        // acc_add_big
    return t + ALIGN;
}

// res[b][j][k][f] (+)= ppol[idx[b]][k][f] * v[b][j][k][f] : svp with a per-item gathered SvpPPol column
int cggi_blind_rotate_impl(pgb_module *m, pgb_vec_znx *res, const int64_t *lwe_2n, uint64_t n_lwe, const pgb_vec_znx *lut,
                           const pgb_vmp_pmat *brk, const pgb_svp_ppol *x_pow_a, uint64_t block_size, uint64_t base2k, const pgb_batch *bt,
                           void *scratch, size_t scratch_len);

extern "C" int pgb_cggi_blind_rotate_batched(pgb_module *m, pgb_vec_znx *res, const int64_t *lwe_2n, uint64_t n_lwe, const pgb_vec_znx *lut,
                                             const pgb_vmp_pmat *brk, const pgb_svp_ppol *x_pow_a, uint64_t block_size, uint64_t base2k,
                                             const pgb_batch *bt, void *scratch, size_t scratch_len) {
    return cggi_blind_rotate_impl(m, res, lwe_2n, n_lwe, lut, brk, x_pow_a, block_size, base2k, bt, scratch, scratch_len);
}

int cggi_blind_rotate_impl(pgb_module *, pgb_vec_znx *, const int64_t *, uint64_t, const pgb_vec_znx *, const pgb_vmp_pmat *,
                           const pgb_svp_ppol *, uint64_t, uint64_t, const pgb_batch *, void *, size_t) {
    pgb_set_error("cggi_blind_rotate: not implemented yet");
    return PGB_ERR_UNSUPPORTED;
}
