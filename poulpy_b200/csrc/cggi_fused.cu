// cggi_fused.cu -- CGGI blind rotation (block-binary), FFT64 flavour, as ONE persistent kernel per batch: the accumulator of
// G ciphertexts lives in shared memory for the whole bootstrap, the bootstrapping key streams from L2 and every loaded key value
// is reused for the G ciphertexts of the CTA.  Per block of `block_size` LWE coefficients (algorithm.rs:338-367):
//   acc_dft = FFT(acc)                                   (vec_znx_dft_apply, 4 limbs)
//   acc_add = sum_t (X^{a_t} - 1) * (acc_dft x BRK_t)    (vmp_apply_dft_to_dft + svp_apply_dft_to_dft + dft add/sub, fused)
//   acc     = normalize(round(IFFT(acc_add)) + acc)      (vec_znx_idft_apply + big_add_small_assign + big_normalize)
// Nothing but the final accumulator, the LWE coefficients and the key stream touches global memory: the unfused HAL sequence
// moves ~1.2 MB per block and ciphertext through HBM, this kernel moves 32 KB (the i64 accumulator, L2 resident).
#include "internal.h"
#include "fft64.cuh"

struct CggiFusedArgs {
    long long *res;          uint64_t res_stride;   // GLWE VecZnx(cols, out_size), i64 words between ciphertexts
    const long long *lwe;    uint64_t lwe_stride;   // mod-switched LWE (b, a_0 .. a_{n_lwe-1}) per ciphertext
    const double *brk;       uint64_t brk_doubles;  // prepared GGSW i: brk + i * brk_doubles, layout [r][p][re(m) | im(m)]
    const double *xpa;                               // x_pow_a table: 2n polys of n doubles
    int n_lwe, block_size, base2k, cols, dnum, brk_size, out_size, batch;
};

template <int L, int L0> struct SmFwd {
    static __device__ __forceinline__ void run(double2 *buf, const double2 *tw, int t, bool valid) {
        constexpr int SL = L - L0 - 3;
        if (valid) {
            const int a = t >> SL, b = t & ((1 << SL) - 1), base = (a << (SL + 3)) | b;
            double2 x[8];
#pragma unroll
            for (int j = 0; j < 8; j++) x[j] = buf[FPAD(base + (j << SL))];
            fct_radix8<3>(x, tw, (1u << L0) | (uint32_t)a);
#pragma unroll
            for (int j = 0; j < 8; j++) buf[FPAD(base + (j << SL))] = x[j];
        }
        __syncthreads();
        SmFwd<L, (L0 + 3 < L) ? L0 + 3 : L>::run(buf, tw, t, valid);
    }
};
template <int L> struct SmFwd<L, L> {
    static __device__ __forceinline__ void run(double2 *, const double2 *, int, bool) {}
};
template <int L, int L0> struct SmInv {
    static __device__ __forceinline__ void run(double2 *buf, const double2 *tw, int t, bool valid) {
        typedef FGeo<L> G;
        constexpr int SL = L - L0 - 3;
        if (valid) {
            const int a = t >> SL, b = t & ((1 << SL) - 1), base = (a << (SL + 3)) | b;
            double2 x[8];
#pragma unroll
            for (int j = 0; j < 8; j++) x[j] = buf[FPAD(base + (j << SL))];
            fgs_radix8<3>(x, tw, (1u << L0) | (uint32_t)a);
#pragma unroll
            for (int j = 0; j < 8; j++) buf[FPAD(base + (j << SL))] = x[j];
        }
        __syncthreads();
        SmInv<L, (L0 - 3 >= G::R0) ? L0 - 3 : -1>::run(buf, tw, t, valid);
    }
};
template <int L> struct SmInv<L, -1> {
    static __device__ __forceinline__ void run(double2 *, const double2 *, int, bool) {}
};

// res_add[g][p][f] = (res_add + xpa[pos_g][f] * v) - v,  v = sum_r acc_dft[g][r][f] * brk[r][p][f]   for PC polys p0..p0+PC-1
template <int RT, int PC, int G, int M, int PL>
__device__ __forceinline__ void vmp_xai_chunk(const double2 *acc_dft, double2 *acc_add, const double *bk, const double *xpa, const int *s_pos,
                                              int C, int p0, int f, int npoly) {
    constexpr int N = 2 * M;
    double br[RT][PC], bi[RT][PC];
#pragma unroll
    for (int r = 0; r < RT; r++)
#pragma unroll
        for (int q = 0; q < PC; q++) {
            const bool ok = q < npoly;
            const double *pp = bk + ((size_t)r * C + p0 + (ok ? q : 0)) * N + f;
            br[r][q] = __ldg(pp);
            bi[r][q] = __ldg(pp + M);
        }
    double wr[G], wi[G];
#pragma unroll
    for (int g = 0; g < G; g++) { // the G table look-ups are independent: issue them together
        const double *w = xpa + (size_t)s_pos[g] * N + f;
        wr[g] = __ldg(w);
        wi[g] = __ldg(w + M);
    }
#pragma unroll
    for (int g = 0; g < G; g++) {
#pragma unroll
        for (int q = 0; q < PC; q++) {
            if (q < npoly) {
                double vr = 0.0, vi = 0.0;
#pragma unroll
                for (int r = 0; r < RT; r++) { // row order of reim4_add_mul (reim4/arithmetic_ref.rs:223-232), FMA-contracted
                    const double2 a = acc_dft[(g * RT + r) * PL + FPAD(f)];
                    vr = fma(a.x, br[r][q], vr);
                    vr = fma(-a.y, bi[r][q], vr);
                    vi = fma(a.x, bi[r][q], vi);
                    vi = fma(a.y, br[r][q], vi);
                }
                const double pr = fma(wr[g], vr, -(wi[g] * vi)), pi = fma(wr[g], vi, wi[g] * vr); // svp: reim_mul(ppol, v)
                double2 *ap = acc_add + (g * C + p0 + q) * PL + FPAD(f);
                double2 acc = *ap;
                acc.x = (acc.x + pr) - vr; // dft_add_assign then dft_sub_assign
                acc.y = (acc.y + pi) - vi;
                *ap = acc;
            }
        }
    }
}

template <int LM, int G, int RT> __global__ void __launch_bounds__(G << LM) cggi_fused_fft64_kernel(CggiFusedArgs p, const double2 *__restrict__ twf,
                                                                                                    const double2 *__restrict__ twi, double inv_m) {
    typedef FGeo<LM> FG;
    constexpr int M = 1 << LM, N = 2 * M, T = FG::T, NT = G * M, PL = FG::PLANE;
    static_assert(LM > FG::R0, "needs at least two passes");
    extern __shared__ __align__(16) double2 csm[];
    __shared__ int s_pos[G];
    const int cols = p.cols, C = cols * p.brk_size, K = p.base2k;
    double2 *acc_dft = csm;                // [G][RT][PL]
    double2 *acc_add = csm + G * RT * PL;  // [G][C][PL]
    const int tid = threadIdx.x, slot = tid / T, t = tid % T;
    const int ct0 = blockIdx.x * G;
    const int mn_small = min(p.brk_size, p.out_size);
    const int a_start = min(p.out_size, p.brk_size); // same-base2k plan with offset 0: limbs >= a_start only feed the carry

    for (int blk = 0; blk + p.block_size <= p.n_lwe; blk += p.block_size) {
        // ---- acc_dft = FFT(acc) ------------------------------------------------------------------------------
        {
            const bool valid = slot < G * RT;
            const int g = valid ? slot / RT : 0, r = valid ? slot % RT : 0, limb = r / cols, col = r % cols;
            double2 *buf = acc_dft + (g * RT + r) * PL;
            if (valid) {
                const int ct = ct0 + g;
                const bool live = ct < p.batch && limb < p.out_size;
                const long long *src = p.res + (size_t)(live ? ct : 0) * p.res_stride + (size_t)(limb * cols + col) * N;
#pragma unroll
                for (int jj = 0; jj < 8; jj++) {
                    const int idx = t + jj * T;
                    buf[FPAD(idx)] = live ? make_double2((double)src[idx], (double)src[idx + M]) : make_double2(0.0, 0.0);
                }
                double2 x[8];
#pragma unroll
                for (int jj = 0; jj < 8; jj++) x[jj] = buf[FPAD(t + jj * T)];
                fct_radix8<FG::R0>(x, twf, 1u);
#pragma unroll
                for (int jj = 0; jj < 8; jj++) buf[FPAD(t + jj * T)] = x[jj];
            }
            // zero the accumulator of the block while the transforms run
            for (int i = tid; i < G * C * PL; i += NT) acc_add[i] = make_double2(0.0, 0.0);
            __syncthreads();
            SmFwd<LM, FG::R0>::run(buf, twf, t, valid);
        }
        // ---- acc_add += (X^{a_t} - 1) * (acc_dft x BRK_t) for the keys of the block -----------------------------------
        for (int tt = 0; tt < p.block_size; tt++) {
            if (tid < G) {
                const int ct = ct0 + tid;
                const long long ai = ct < p.batch ? p.lwe[(size_t)ct * p.lwe_stride + 1 + blk + tt] : 0;
                s_pos[tid] = (int)((ai + (long long)(2 * N)) & (long long)(2 * N - 1));
            }
            __syncthreads();
            const double *bk = p.brk + (size_t)(blk + tt) * p.brk_doubles;
            const int f = tid % M, pg = tid / M;
            constexpr int PC = 1;
            for (int p0 = pg * PC; p0 < C; p0 += PC * G)
                vmp_xai_chunk<RT, PC, G, M, PL>(acc_dft, acc_add, bk, p.xpa, s_pos, C, p0, f, min(PC, C - p0));
            __syncthreads();
        }
        // ---- acc = normalize(round(IFFT(acc_add) / m) + acc) -------------------------------------------------------------
        {
            const bool valid = slot < G * C;
            double2 *buf = acc_add + (valid ? slot : 0) * PL;
            if (valid) {
                double2 x[8];
#pragma unroll
                for (int jj = 0; jj < 8; jj++) x[jj] = buf[FPAD(8 * t + jj)];
                fgs_radix8<3>(x, twi, (1u << (LM - 3)) | (uint32_t)t);
#pragma unroll
                for (int jj = 0; jj < 8; jj++) buf[FPAD(8 * t + jj)] = x[jj];
            }
            __syncthreads();
            SmInv<LM, (LM - 6 >= FG::R0) ? LM - 6 : -1>::run(buf, twi, t, valid);
            double2 x[8];
            if (valid) {
#pragma unroll
                for (int jj = 0; jj < 8; jj++) x[jj] = buf[FPAD(t + jj * T)];
                fgs_radix8<FG::R0>(x, twi, 1u);
            }
            __syncthreads(); // every transform has read its inputs: the buffers are now reused for the rounded i64 coefficients
            if (valid) {
                long long *big = reinterpret_cast<long long *>(buf);
#pragma unroll
                for (int jj = 0; jj < 8; jj++) {
                    const int idx = t + jj * T;
                    big[idx] = (long long)round(x[jj].x * inv_m); // reim_to_znx_i64 (conversion.rs:43-52)
                    big[idx + M] = (long long)round(x[jj].y * inv_m);
                }
            }
            __syncthreads();
            for (int item = tid; item < G * cols * N; item += NT) {
                const int g = item / (cols * N), col = (item / N) % cols, i = item % N;
                const int ct = ct0 + g;
                if (ct >= p.batch) continue;
                long long *acc = p.res + (size_t)ct * p.res_stride + (size_t)col * N + i; // limb j at + j*cols*N
                long long c = 0;
                for (int j = p.brk_size - 1; j >= 0; j--) {
                    long long v = reinterpret_cast<const long long *>(acc_add + (g * C + j * cols + col) * PL)[i];
                    if (j < mn_small) v = (long long)((unsigned long long)v + (unsigned long long)acc[(size_t)j * cols * N]);
                    const long long tsum = (long long)((unsigned long long)v + (unsigned long long)c);
                    const long long out = (long long)((unsigned long long)tsum << (64 - K)) >> (64 - K);
                    c = (long long)((unsigned long long)tsum - (unsigned long long)out) >> K;
                    if (j < a_start) acc[(size_t)j * cols * N] = out;
                }
                for (int j = a_start; j < p.out_size; j++) acc[(size_t)j * cols * N] = 0;
            }
            __syncthreads();
        }
    }
}

template <int LM, int G, int RT> static int launch_cggi(pgb_module *m, const CggiFusedArgs &p, int C) {
    typedef FGeo<LM> FG;
    const size_t smem = (size_t)G * (RT + C) * FG::PLANE * sizeof(double2);
    PGB_CHECK_CUDA(cudaFuncSetAttribute(cggi_fused_fft64_kernel<LM, G, RT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (p.batch + G - 1) / G;
    { ProfScope _ps(m, PROF_OTHER);
    cggi_fused_fft64_kernel<LM, G, RT><<<grid, G << LM, smem, m->stream>>>(p, m->fft_fwd, m->fft_inv, 1.0 / (double)(1 << LM));
    }
    PGB_CHECK_CUDA(cudaGetLastError());
    return PGB_OK;
}

bool cggi_fused_supported(const pgb_module *m, uint64_t cols, uint64_t dnum, uint64_t brk_size) {
    if (m->flavour != PGB_FFT64 || m->log_n < 8 || m->log_n > 11) return false;
    const uint64_t R = cols * dnum, C = cols * brk_size;
    if (R < 1 || R > 8 || C > 8) return false;
    const uint64_t M = m->n / 2, G = 1024 / M, PL = M + (M >> 3) + 2;
    return G * (R + C) * PL * 16 + 64 <= 227 * 1024;
}

template <int LM, int G> static int launch_cggi_rt(pgb_module *m, const CggiFusedArgs &p, int R, int C) {
    switch (R) {
    case 1: return launch_cggi<LM, G, 1>(m, p, C);
    case 2: return launch_cggi<LM, G, 2>(m, p, C);
    case 3: return launch_cggi<LM, G, 3>(m, p, C);
    case 4: return launch_cggi<LM, G, 4>(m, p, C);
    case 6: return launch_cggi<LM, G, 6>(m, p, C);
    case 8: return launch_cggi<LM, G, 8>(m, p, C);
    default: pgb_set_error("cggi fused: unsupported row count %d", R); return PGB_ERR_UNSUPPORTED;
    }
}

// res must already hold X^b * LUT in column 0 (and zeros elsewhere)
int cggi_fused_fft64(pgb_module *m, long long *res, uint64_t res_stride_words, const long long *lwe, uint64_t lwe_stride, const double *brk,
                     uint64_t brk_doubles, const double *xpa, int n_lwe, int block_size, int base2k, int cols, int dnum, int brk_size,
                     int out_size, int batch) {
    CggiFusedArgs p = {res, res_stride_words, lwe, lwe_stride, brk, brk_doubles, xpa, n_lwe, block_size, base2k, cols, dnum, brk_size, out_size, batch};
    const int R = cols * dnum, C = cols * brk_size;
    switch (m->log_n) {
    case 8: return launch_cggi_rt<7, 8>(m, p, R, C);
    case 9: return launch_cggi_rt<8, 4>(m, p, R, C);
    case 10: return launch_cggi_rt<9, 2>(m, p, R, C);
    case 11: return launch_cggi_rt<10, 1>(m, p, R, C);
    default: pgb_set_error("cggi fused: unsupported n"); return PGB_ERR_UNSUPPORTED;
    }
}
