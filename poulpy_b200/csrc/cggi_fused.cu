// cggi_fused.cu -- CGGI blind rotation (block-binary), FFT64 flavour, as ONE persistent kernel per batch: the accumulator of
// G ciphertexts lives in shared memory for the whole bootstrap, the bootstrapping key streams from L2 and every loaded key value
// is reused for the G ciphertexts of the CTA.  Per block of `block_size` LWE coefficients (algorithm.rs:338-367):
//   acc_dft = FFT(acc)                                   (vec_znx_dft_apply, 4 limbs)
//   acc_add = sum_t (X^{a_t} - 1) * (acc_dft x BRK_t)    (vmp_apply_dft_to_dft + svp_apply_dft_to_dft + dft add/sub, fused)
//   acc     = normalize(round(IFFT(acc_add)) + acc)      (vec_znx_idft_apply + big_add_small_assign + big_normalize)
// Nothing but the final accumulator, the LWE coefficients and the key stream touches global memory: the unfused HAL sequence
// moves ~1.2 MB per block and ciphertext through HBM, this kernel moves 32 KB (the i64 accumulator, L2 resident).
#include <stdlib.h>

#include "internal.h"
#include "fft64.cuh"
#include "tma.cuh"

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

// ---- version 2: register-blocked key products ----------------------------------------------------------------------
// ncu on the kernel above (profiles/r1_ncu_fused_v2.md): mio_throttle / long_scoreboard stalls, FP64 pipe 22 % busy.  Per key and
// frequency it re-reads the 16 accumulator values and read-modify-writes the 32 partial sums through shared memory (288 complex
// LDS/STS per frequency and block) -- at 128 B/clk of shared-memory bandwidth that, not FP64, is the limit.  Here one thread owns a
// (ciphertext, frequency) pair for the whole block: its RT accumulator values and its C partial sums stay in registers across the
// block_size keys (48 complex shared-memory accesses per frequency and block), the lanes of a warp cover 32/G consecutive
// frequencies x G ciphertexts so that a key value is fetched once per warp and broadcast, and the products are written over the
// transform buffers they were read from (out[g][p] aliases acc_dft[g][p]; every thread touches only its own frequency).
template <int LM, int G, int RT, int CT> __global__ void __launch_bounds__(512, 1)
cggi_fused2_fft64_kernel(CggiFusedArgs p, const double2 *__restrict__ twf, const double2 *__restrict__ twi, double inv_m) {
    typedef FGeo<LM> FG;
    constexpr int M = 1 << LM, N = 2 * M, T = FG::T, NT = 512, PL = FG::PLANE, NSLOT = NT / T;
    constexpr int PMAX = RT > CT ? RT : CT; // polys per ciphertext held in shared memory
    static_assert(G * M == 1024 && LM > FG::R0, "geometry");
    extern __shared__ __align__(16) double2 csm[];
    __shared__ int s_pos[G * 8];
    const int cols = p.cols, C = cols * p.brk_size, K = p.base2k, bs = p.block_size;
    const int tid = threadIdx.x, slot = tid / T, t = tid % T;
    const int ct0 = blockIdx.x * G;
    const int mn_small = min(p.brk_size, p.out_size);
    const int a_start = min(p.out_size, p.brk_size);

    for (int blk = 0; blk + bs <= p.n_lwe; blk += bs) {
        // rotation amounts of the block (at most 8 keys per block, checked on the host)
        if (tid < G * bs) {
            const int g = tid / bs, tt = tid % bs, ct = ct0 + g;
            const long long ai = ct < p.batch ? p.lwe[(size_t)ct * p.lwe_stride + 1 + blk + tt] : 0;
            s_pos[g * 8 + tt] = (int)((ai + (long long)(2 * N)) & (long long)(2 * N - 1));
        }
        // ---- acc_dft = FFT(acc): G * RT transforms, NSLOT at a time -------------------------------------------------------
        for (int base = 0; base < G * RT; base += NSLOT) {
            const int job = base + slot;
            const bool valid = job < G * RT;
            const int g = valid ? job / RT : 0, r = valid ? job % RT : 0, limb = r / cols, col = r % cols;
            double2 *buf = csm + (g * PMAX + r) * PL;
            if (valid) {
                const int ct = ct0 + g;
                const bool live = ct < p.batch && limb < p.out_size;
                const long long *src = p.res + (size_t)(live ? ct : 0) * p.res_stride + (size_t)(limb * cols + col) * N;
                double2 x[8];
#pragma unroll
                for (int jj = 0; jj < 8; jj++) {
                    const int idx = t + jj * T;
                    x[jj] = live ? make_double2((double)src[idx], (double)src[idx + M]) : make_double2(0.0, 0.0);
                }
                fct_radix8<FG::R0>(x, twf, 1u);
#pragma unroll
                for (int jj = 0; jj < 8; jj++) buf[FPAD(t + jj * T)] = x[jj];
            }
            __syncthreads();
            SmFwd<LM, FG::R0>::run(buf, twf, t, valid);
        }
        // ---- out[g][p] = sum_t (X^{a_t} - 1) * (acc_dft[g] x BRK_t)[p], one (g, f) pair per thread and round ------------------
#pragma unroll 1
        for (int it = tid; it < G * M; it += NT) {
            const int g = it % G, f = it / G;
            double2 *mine = csm + (size_t)g * PMAX * PL + FPAD(f);
            double ar[RT], ai[RT];
#pragma unroll
            for (int r = 0; r < RT; r++) {
                const double2 a = mine[r * PL];
                ar[r] = a.x;
                ai[r] = a.y;
            }
            double sr[CT], si[CT];
#pragma unroll
            for (int q = 0; q < CT; q++) sr[q] = si[q] = 0.0;
#pragma unroll 1
            for (int tt = 0; tt < bs; tt++) {
                const double *bk = p.brk + (size_t)(blk + tt) * p.brk_doubles + f;
                const double *w = p.xpa + (size_t)s_pos[g * 8 + tt] * N + f;
                const double wr = __ldg(w), wi = __ldg(w + M);
#pragma unroll
                for (int q = 0; q < CT; q++) {
                    if (q < C) {
                        double vr = 0.0, vi = 0.0;
#pragma unroll
                        for (int r = 0; r < RT; r++) { // row order of reim4_add_mul (reim4/arithmetic_ref.rs:223-232), FMA-contracted
                            const double *pp = bk + ((size_t)r * C + q) * N;
                            const double br = __ldg(pp), bi = __ldg(pp + M);
                            vr = fma(ar[r], br, vr);
                            vr = fma(-ai[r], bi, vr);
                            vi = fma(ar[r], bi, vi);
                            vi = fma(ai[r], br, vi);
                        }
                        const double pr = fma(wr, vr, -(wi * vi)), pi = fma(wr, vi, wi * vr); // svp: reim_mul(ppol, v)
                        sr[q] = (sr[q] + pr) - vr; // dft_add_assign then dft_sub_assign
                        si[q] = (si[q] + pi) - vi;
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < CT; q++)
                if (q < C) mine[q * PL] = make_double2(sr[q], si[q]);
        }
        __syncthreads();
        // ---- acc = normalize(round(IFFT(out) / m) + acc): G * C transforms, NSLOT at a time --------------------------------
        for (int base = 0; base < G * C; base += NSLOT) {
            const int job = base + slot;
            const bool valid = job < G * C;
            const int g = valid ? job / C : 0, q = valid ? job % C : 0;
            double2 *buf = csm + (g * PMAX + q) * PL;
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
            __syncthreads(); // the transform has read its inputs: the buffer is reused for the rounded i64 coefficients
            if (valid) {
                long long *big = reinterpret_cast<long long *>(buf);
#pragma unroll
                for (int jj = 0; jj < 8; jj++) {
                    const int idx = t + jj * T;
                    big[idx] = (long long)round(x[jj].x * inv_m); // reim_to_znx_i64 (conversion.rs:43-52)
                    big[idx + M] = (long long)round(x[jj].y * inv_m);
                }
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
                long long v = reinterpret_cast<const long long *>(csm + (g * PMAX + j * cols + col) * PL)[i];
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

// Synchronisation among the T = m/8 threads that own one transform: a warp (or less) needs no CTA barrier at all, larger groups
// use one named barrier per transform slot.  The CTA-wide __syncthreads of versions 1/2 made every radix-8 pass wait for all warps.
template <int T> __device__ __forceinline__ void poly_sync(int slot) {
    if (T <= 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "r"(T) : "memory");
}
// tw: shared-memory copy of the first T = m/8 block twiddles (all a pass other than the last one touches); twl: shared-memory copy of
// the last-pass table [7][T] (fft64.cuh: load_tw7) -- conflict-free where the strided reads of tw[4 hi + j] were 4-way conflicted
template <int L, int L0> struct SmFwdP {
    static __device__ __forceinline__ void run(double2 *buf, const double2 *tw, const double2 *twl, int t, int slot, bool valid) {
        constexpr int SL = L - L0 - 3;
        if (valid) {
            const int a = t >> SL, b = t & ((1 << SL) - 1), base = (a << (SL + 3)) | b;
            double2 x[8];
#pragma unroll
            for (int j = 0; j < 8; j++) x[j] = buf[FPAD(base + (j << SL))];
            if (SL == 0) {
                double2 w[7];
                load_tw7<true>(w, twl, FGeo<L>::T, t);
                fct_radix8_w(x, w);
            } else {
                fct_radix8<3, true>(x, tw, (1u << L0) | (uint32_t)a);
            }
#pragma unroll
            for (int j = 0; j < 8; j++) buf[FPAD(base + (j << SL))] = x[j];
        }
        poly_sync<FGeo<L>::T>(slot);
        SmFwdP<L, (L0 + 3 < L) ? L0 + 3 : L>::run(buf, tw, twl, t, slot, valid);
    }
};
template <int L> struct SmFwdP<L, L> {
    static __device__ __forceinline__ void run(double2 *, const double2 *, const double2 *, int, int, bool) {}
};
template <int L, int L0> struct SmInvP {
    static __device__ __forceinline__ void run(double2 *buf, const double2 *tw, int t, int slot, bool valid) {
        typedef FGeo<L> G;
        constexpr int SL = L - L0 - 3;
        if (valid) {
            const int a = t >> SL, b = t & ((1 << SL) - 1), base = (a << (SL + 3)) | b;
            double2 x[8];
#pragma unroll
            for (int j = 0; j < 8; j++) x[j] = buf[FPAD(base + (j << SL))];
            fgs_radix8<3, true>(x, tw, (1u << L0) | (uint32_t)a);
#pragma unroll
            for (int j = 0; j < 8; j++) buf[FPAD(base + (j << SL))] = x[j];
        }
        poly_sync<G::T>(slot);
        SmInvP<L, (L0 - 3 >= G::R0) ? L0 - 3 : -1>::run(buf, tw, t, slot, valid);
    }
};
template <int L> struct SmInvP<L, -1> {
    static __device__ __forceinline__ void run(double2 *, const double2 *, int, int, bool) {}
};

// ---- version 3: key stream through TMA bulk copies -------------------------------------------------------------------
// Version 2 stalls on the L2 latency of the key values (long_scoreboard 7.8 warps per issue, FP64 pipe 15 %).  The key stream is
// perfectly predictable, so here thread 0 keeps a ring of NSTAGE shared-memory tiles filled with cp.async.bulk (one tile = the RT
// rows of one output poly of one key = RT contiguous chunks of 8n bytes, completion on an mbarrier), running ahead of the compute
// threads across the transform phases.  A thread owns one frequency of TWO ciphertexts (g and g + G/2): a key value read from the
// tile feeds both, and the 2 x (RT + C) complex values it needs stay in registers for the whole block.
template <int LM, int G, int RT, int CT, int NSTAGE> __global__ void __launch_bounds__(512, 1)
cggi_fused3_fft64_kernel(CggiFusedArgs p, const double2 *__restrict__ twf_g, const double2 *__restrict__ twi_g, const double2 *__restrict__ twlf_g,
                         const double2 *__restrict__ twli_g, double inv_m) {
    typedef FGeo<LM> FG;
    constexpr int M = 1 << LM, N = 2 * M, T = FG::T, NT = 512, PL = FG::PLANE, NSLOT = NT / T, GH = G / 2;
    constexpr int PMAX = RT > CT ? RT : CT;
    // planes of consecutive ciphertexts start 64 bytes (mod 128) apart: in the key products neighbouring threads hold the same frequency of
    // two different ciphertexts, and a plane stride that is a multiple of 128 bytes put both on the same banks (2-way conflicts)
    constexpr int GS = PMAX * PL + 4;
    constexpr uint32_t CHUNK = N * 8, TILE = RT * CHUNK; // bytes
    static_assert(GH * M == NT && GH >= 1 && LM > FG::R0, "geometry");
    extern __shared__ __align__(128) double2 csm[];
    __shared__ int s_pos[G * 8];
    __shared__ __align__(8) unsigned long long s_bar[NSTAGE], s_empty[NSTAGE]; // tile filled / tile consumed by all threads
    double *ring = reinterpret_cast<double *>(csm + (size_t)G * GS); // [NSTAGE][RT][N]
    // both twiddle tables (m complex values each) live in shared memory: with ~210 KB of it in use the L1 is too small to keep them
    // M entries per direction as before, split into the first T block twiddles and the [7][T] last-pass table
    double2 *twf = reinterpret_cast<double2 *>(ring + (size_t)NSTAGE * RT * N), *twlf = twf + T, *twi = twf + M, *twli = twi + T;
    for (int i = threadIdx.x; i < T; i += 512) {
        twf[i] = twf_g[i];
        twi[i] = twi_g[i];
    }
    for (int i = threadIdx.x; i < 7 * T; i += 512) {
        twlf[i] = twlf_g[i];
        twli[i] = twli_g[i];
    }
    const int cols = p.cols, C = cols * p.brk_size, K = p.base2k, bs = p.block_size;
    const int tid = threadIdx.x, slot = tid / T, t = tid % T;
    const int ct0 = blockIdx.x * G;
    const int mn_small = min(p.brk_size, p.out_size);
    const int a_start = min(p.out_size, p.brk_size);
    const int nblk = p.n_lwe / bs, tiles_per_blk = bs * C, total_tiles = nblk * tiles_per_blk;
    const uint32_t ring_s = smem_u32(ring), bar_s = smem_u32(s_bar), empty_s = smem_u32(s_empty);

    // tile gk -> key (gk / C), output poly gk % C: rows r at brk + key * brk_doubles + (r * C + q) * N
    auto issue = [&](int gk) {
        const int st = gk % NSTAGE, key = gk / C, q = gk % C;
        const uint32_t bar = bar_s + st * 8;
        mbar_expect_tx(bar, TILE);
        const double *src = p.brk + (size_t)key * p.brk_doubles + (size_t)q * N;
#pragma unroll
        for (int r = 0; r < RT; r++) bulk_g2s(ring_s + (uint32_t)(st * RT + r) * CHUNK, src + (size_t)r * C * N, CHUNK, bar);
    };
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; s++) {
            mbar_init(bar_s + s * 8, 1);
            mbar_init(empty_s + s * 8, NT);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0)
        for (int gk = 0; gk < NSTAGE && gk < total_tiles; gk++) issue(gk);
    const int gp = tid % GH, f = tid / GH; // this thread's frequency and its two ciphertexts gp, gp + GH
    double2 *mine0 = csm + (size_t)gp * GS + FPAD(f), *mine1 = csm + (size_t)(gp + GH) * GS + FPAD(f);
    int gk = 0;

    for (int blk = 0; blk + bs <= p.n_lwe; blk += bs) {
        if (tid < G * bs) {
            const int g = tid / bs, tt = tid % bs, ct = ct0 + g;
            const long long ai = ct < p.batch ? p.lwe[(size_t)ct * p.lwe_stride + 1 + blk + tt] : 0;
            s_pos[g * 8 + tt] = (int)((ai + (long long)(2 * N)) & (long long)(2 * N - 1));
        }
        // ---- acc_dft = FFT(acc) ------------------------------------------------------------------------------------------
        for (int base = 0; base < G * RT; base += NSLOT) {
            const int job = base + slot;
            const bool valid = job < G * RT;
            const int g = valid ? job / RT : 0, r = valid ? job % RT : 0, limb = r / cols, col = r % cols;
            double2 *buf = csm + g * GS + r * PL;
            if (valid) {
                const int ct = ct0 + g;
                const bool live = ct < p.batch && limb < p.out_size;
                const long long *src = p.res + (size_t)(live ? ct : 0) * p.res_stride + (size_t)(limb * cols + col) * N;
                double2 x[8];
#pragma unroll
                for (int jj = 0; jj < 8; jj++) {
                    const int idx = t + jj * T;
                    x[jj] = live ? make_double2((double)src[idx], (double)src[idx + M]) : make_double2(0.0, 0.0);
                }
                fct_radix8<FG::R0, true>(x, twf, 1u);
#pragma unroll
                for (int jj = 0; jj < 8; jj++) buf[FPAD(t + jj * T)] = x[jj];
            }
            poly_sync<T>(slot);
            SmFwdP<LM, FG::R0>::run(buf, twf, twlf, t, slot, valid);
        }
        __syncthreads(); // every transform of the block is complete before the key products read across them
        // ---- key products: tiles in (key, poly) order -----------------------------------------------------------------------
        {
            double a0r[RT], a0i[RT], a1r[RT], a1i[RT];
#pragma unroll
            for (int r = 0; r < RT; r++) {
                const double2 u = mine0[r * PL], v = mine1[r * PL];
                a0r[r] = u.x; a0i[r] = u.y; a1r[r] = v.x; a1i[r] = v.y;
            }
            double s0r[CT], s0i[CT], s1r[CT], s1i[CT];
#pragma unroll
            for (int q = 0; q < CT; q++) s0r[q] = s0i[q] = s1r[q] = s1i[q] = 0.0;
#pragma unroll 1
            for (int tt = 0; tt < bs; tt++) {
                const double *w0 = p.xpa + (size_t)s_pos[gp * 8 + tt] * N + f, *w1 = p.xpa + (size_t)s_pos[(gp + GH) * 8 + tt] * N + f;
                const double w0r = __ldg(w0), w0i = __ldg(w0 + M), w1r = __ldg(w1), w1i = __ldg(w1 + M);
#pragma unroll
                for (int q = 0; q < CT; q++) {
                    if (q < C) { // uniform
                        const int st = gk % NSTAGE;
                        mbar_wait(bar_s + st * 8, (uint32_t)((gk / NSTAGE) & 1));
                        const double *tile = ring + (size_t)st * RT * N + f;
                        double v0r = 0.0, v0i = 0.0, v1r = 0.0, v1i = 0.0;
#pragma unroll
                        for (int r = 0; r < RT; r++) { // row order of reim4_add_mul (reim4/arithmetic_ref.rs:223-232), FMA-contracted
                            const double br = tile[r * N], bi = tile[r * N + M];
                            v0r = fma(a0r[r], br, v0r); v0r = fma(-a0i[r], bi, v0r);
                            v0i = fma(a0r[r], bi, v0i); v0i = fma(a0i[r], br, v0i);
                            v1r = fma(a1r[r], br, v1r); v1r = fma(-a1i[r], bi, v1r);
                            v1i = fma(a1r[r], bi, v1i); v1i = fma(a1i[r], br, v1i);
                        }
                        const double p0r = fma(w0r, v0r, -(w0i * v0i)), p0i = fma(w0r, v0i, w0i * v0r); // svp: reim_mul(ppol, v)
                        const double p1r = fma(w1r, v1r, -(w1i * v1i)), p1i = fma(w1r, v1i, w1i * v1r);
                        s0r[q] = (s0r[q] + p0r) - v0r; s0i[q] = (s0i[q] + p0i) - v0i; // dft_add_assign then dft_sub_assign
                        s1r[q] = (s1r[q] + p1r) - v1r; s1i[q] = (s1i[q] + p1i) - v1i;
                        // consumer release; thread 0 refills the stage once all 512 threads have released it (only its warp waits).  Measured
                        // alternatives: one arrival per warp after __syncwarp 94.9 k bootstraps/s (against 99.7 k), a non-blocking thread 0
                        // that polls at tile boundaries 75.6 k -- the refill must be requested the moment the stage is free; version 4 below
                        // gives the ring its own warp
                        mbar_arrive(empty_s + st * 8);
                        if (tid == 0 && gk + NSTAGE < total_tiles) {
                            mbar_wait(empty_s + st * 8, (uint32_t)((gk / NSTAGE) & 1));
                            issue(gk + NSTAGE);
                        }
                        gk++;
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < CT; q++)
                if (q < C) {
                    mine0[q * PL] = make_double2(s0r[q], s0i[q]);
                    mine1[q * PL] = make_double2(s1r[q], s1i[q]);
                }
        }
        __syncthreads();
        // ---- acc = normalize(round(IFFT(out) / m) + acc) -----------------------------------------------------------------------
        for (int base = 0; base < G * C; base += NSLOT) {
            const int job = base + slot;
            const bool valid = job < G * C;
            const int g = valid ? job / C : 0, q = valid ? job % C : 0;
            double2 *buf = csm + g * GS + q * PL;
            if (valid) {
                double2 x[8];
#pragma unroll
                for (int jj = 0; jj < 8; jj++) x[jj] = buf[FPAD(8 * t + jj)];
                double2 w[7];
                load_tw7<true>(w, twli, T, t);
                fgs_radix8_w(x, w);
#pragma unroll
                for (int jj = 0; jj < 8; jj++) buf[FPAD(8 * t + jj)] = x[jj];
            }
            poly_sync<T>(slot);
            SmInvP<LM, (LM - 6 >= FG::R0) ? LM - 6 : -1>::run(buf, twi, t, slot, valid);
            double2 x[8];
            if (valid) {
#pragma unroll
                for (int jj = 0; jj < 8; jj++) x[jj] = buf[FPAD(t + jj * T)];
                fgs_radix8<FG::R0, true>(x, twi, 1u);
            }
            poly_sync<T>(slot); // the transform has read its inputs: the buffer is reused for the rounded i64 coefficients
            if (valid) {
                long long *big = reinterpret_cast<long long *>(buf);
#pragma unroll
                for (int jj = 0; jj < 8; jj++) {
                    const int idx = t + jj * T;
                    big[idx] = (long long)round(x[jj].x * inv_m); // reim_to_znx_i64 (conversion.rs:43-52)
                    big[idx + M] = (long long)round(x[jj].y * inv_m);
                }
            }
        }
        __syncthreads();
        // brk_size <= 4 here (checked on the host): all loads of a coefficient are issued before its carry chain; 32-bit offsets
        {
            const int limb_w = cols * N, bsz = p.brk_size;
            const bool two_to_one = bsz == 2 && p.out_size == 1; // the bench shape (k_brk = 2 limbs, k_glwe = 1 limb): no predicated loads
            for (int g = 0; g < G; g++) {
                if (ct0 + g >= p.batch) break;
                long long *acc_g = p.res + (size_t)(ct0 + g) * p.res_stride;
                const long long *big_g = reinterpret_cast<const long long *>(csm + (size_t)g * GS);
                if (two_to_one) {
                    for (int col = 0; col < cols; col++) {
#pragma unroll
                        for (int i = tid; i < N; i += NT) {
                            const long long v0 = big_g[col * (2 * PL) + i], v1 = big_g[(cols + col) * (2 * PL) + i];
                            long long *ap = acc_g + col * N + i;
                            const long long a0 = *ap;
                            const long long o1 = (long long)((unsigned long long)v1 << (64 - K)) >> (64 - K);
                            const long long c = (long long)((unsigned long long)v1 - (unsigned long long)o1) >> K;
                            const long long tsum = (long long)((unsigned long long)v0 + (unsigned long long)a0 + (unsigned long long)c);
                            *ap = (long long)((unsigned long long)tsum << (64 - K)) >> (64 - K);
                        }
                    }
                    continue;
                }
                for (int col = 0; col < cols; col++) {
#pragma unroll
                    for (int i = tid; i < N; i += NT) {
                        const int o = col * N + i;
                        long long vv[4], aa[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            vv[j] = j < bsz ? big_g[(j * cols + col) * (2 * PL) + i] : 0;
                            aa[j] = j < mn_small ? acc_g[o + j * limb_w] : 0;
                        }
                        long long c = 0;
#pragma unroll
                        for (int j = 3; j >= 0; j--) {
                            if (j < bsz) {
                                const long long tsum = (long long)((unsigned long long)vv[j] + (unsigned long long)aa[j] + (unsigned long long)c);
                                const long long out = (long long)((unsigned long long)tsum << (64 - K)) >> (64 - K);
                                c = (long long)((unsigned long long)tsum - (unsigned long long)out) >> K;
                                if (j < a_start) acc_g[o + j * limb_w] = out;
                            }
                        }
                        for (int j = a_start; j < p.out_size; j++) acc_g[o + j * limb_w] = 0;
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ---- version 4: the key ring gets its own warp ------------------------------------------------------------------------------
// ncu on version 3 at the BASELINE shape (profiles/r2_ncu_cggi.md): 18.7 % of the stall samples and 24 % of the executed instructions are
// the try-wait spin on the tile-full barrier -- the compute warps outrun a ring whose refills are requested by a thread that is itself
// one of the consumers.  Here warp 16 does nothing but the ring (as in cggi_ntt_fused.cu, where the same wait is 3 % of the instructions).
// A 17th warp caps the kernel at 96 registers per thread, so the key products are reorganised around ONE output poly at a time (tiles in
// (poly, key) order): 4 partial-sum doubles per thread instead of 64, the X^{a_t} factors of all keys of the block loaded once, no spills.
// Per (ciphertext, frequency, output poly) the floating-point operations and their order are exactly those of version 3 (rows in row
// order inside a key, keys in block order), so the results are bit-identical to it.
// CG4_PRODUCER_WARP = 0 builds the measured alternative without a producer: the warp whose lane 0 is the LAST of the 16 to finish a tile (a
// shared-memory counter per stage, ordering through the "empty" barrier) issues the bulk copies of the tile NSTAGE ahead; the 16-warp CTA
// gets 126 registers per thread instead of 96 and no spill.  Bit-identical, and 23 % SLOWER (98.4 k vs 127.4 k bootstraps/s): the refill is
// ~40 instructions executed by one lane of a consumer warp, which stalls that warp exactly when the ring is shortest -- the same flaw as
// version 3's thread-0 producer.  The dedicated warp is what makes version 4 fast, not its register budget.
#ifndef CG4_PRODUCER_WARP
#define CG4_PRODUCER_WARP 1
#endif
constexpr int CG4_THREADS = 512 + (CG4_PRODUCER_WARP ? 32 : 0);
template <int LM, int G, int RT, int CT, int NSTAGE, int BS> __global__ void __launch_bounds__(CG4_THREADS, 1)
cggi_fused4_fft64_kernel(CggiFusedArgs p, const double2 *__restrict__ twf_g, const double2 *__restrict__ twi_g, const double2 *__restrict__ twlf_g,
                         const double2 *__restrict__ twli_g, double inv_m) {
    typedef FGeo<LM> FG;
    constexpr int M = 1 << LM, N = 2 * M, T = FG::T, NT = 512, PL = FG::PLANE, NSLOT = NT / T, GH = G / 2;
    constexpr int PMAX = RT > CT ? RT : CT;
    // planes of consecutive ciphertexts start 64 bytes (mod 128) apart: in the key products neighbouring threads hold the same frequency of
    // two different ciphertexts, and a plane stride that is a multiple of 128 bytes put both on the same banks (2-way conflicts)
    constexpr int GS = PMAX * PL + 4;
    constexpr uint32_t CHUNK = N * 8, TILE = RT * CHUNK; // bytes
    static_assert(GH * M == NT && GH >= 1 && LM > FG::R0, "geometry");
    extern __shared__ __align__(128) double2 csm[];
    __shared__ int s_pos[G * 8];
    __shared__ __align__(8) unsigned long long s_bar[NSTAGE], s_empty[NSTAGE]; // tile filled / tile consumed by all threads
    __shared__ unsigned int s_done[NSTAGE]; // warps that have finished the tile of this stage (producer-less refill)
    __shared__ __align__(8) unsigned long long s_peer[NSTAGE]; // rank 0 of a cluster: the other CTAs have drained this stage and armed their barrier
    double *ring = reinterpret_cast<double *>(csm + (size_t)G * GS); // [NSTAGE][RT][N]
    // both twiddle tables (m complex values each) live in shared memory: with ~210 KB of it in use the L1 is too small to keep them
    // M entries per direction as before, split into the first T block twiddles and the [7][T] last-pass table
    double2 *twf = reinterpret_cast<double2 *>(ring + (size_t)NSTAGE * RT * N), *twlf = twf + T, *twi = twf + M, *twli = twi + T;
    for (int i = threadIdx.x; i < T; i += CG4_THREADS) {
        twf[i] = twf_g[i];
        twi[i] = twi_g[i];
    }
    for (int i = threadIdx.x; i < 7 * T; i += CG4_THREADS) {
        twlf[i] = twlf_g[i];
        twli[i] = twli_g[i];
    }
    const int cols = p.cols, C = cols * p.brk_size, K = p.base2k, bs = p.block_size;
    const int tid = threadIdx.x, slot = tid / T, t = tid % T;
    const int ct0 = blockIdx.x * G;
    const int mn_small = min(p.brk_size, p.out_size);
    const int a_start = min(p.out_size, p.brk_size);
    const int nblk = p.n_lwe / bs, tiles_per_blk = bs * C, total_tiles = nblk * tiles_per_blk;
    const uint32_t ring_s = smem_u32(ring), bar_s = smem_u32(s_bar), empty_s = smem_u32(s_empty), peer_s = smem_u32(s_peer);
    // Cluster launch (PGB_OPT_CGGI_CLUSTER = 2): the CTAs of a cluster walk the same key tiles, so rank 0 fetches each tile ONCE and the copy is
    // multicast into the ring of every CTA (same stage, same offset) -- half the L2 -> SM key traffic.  The other ranks' producers only arm
    // their own barrier for the stage once their consumers have drained it, and tell rank 0.
    const uint32_t csize = cluster_nctarank(), crank = csize > 1 ? cluster_ctarank() : 0;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; s++) {
            mbar_init(bar_s + s * 8, 1);
            mbar_init(empty_s + s * 8, CG4_PRODUCER_WARP ? NT : NT / 32);
            mbar_init(peer_s + s * 8, csize > 1 ? csize - 1 : 1);
            s_done[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (csize > 1) cluster_sync_all(); // every CTA's barriers exist before a peer arrives on them or a multicast copy signals them
    // ---- key stream: tiles in (block, output poly c, key t) order ------------------------------------------------------------------------
    auto issue_src = [&](const double *src, const uint32_t st_) { // one thread: bulk copies of one tile (RT key rows of one output poly)
        const uint32_t bar = bar_s + st_ * 8;
        mbar_expect_tx(bar, TILE);
#pragma unroll
        for (int r = 0; r < RT; r++) bulk_g2s(ring_s + (uint32_t)(st_ * RT + r) * CHUNK, src + (size_t)r * C * N, CHUNK, bar);
    };
    auto issue_tile = [&](const uint32_t gk_, const uint32_t st_) { // the same from the tile index (producer-less mode)
        const uint32_t b_ = gk_ / (uint32_t)tiles_per_blk, rem_ = gk_ - b_ * (uint32_t)tiles_per_blk, c_ = rem_ / (uint32_t)bs, t_ = rem_ - c_ * (uint32_t)bs;
        issue_src(p.brk + ((size_t)b_ * bs + t_) * p.brk_doubles + (size_t)c_ * N, st_);
    };
    if (CG4_PRODUCER_WARP) {
        if (tid >= NT) { // warp 16: refills a stage the moment all consumers have released it
            if (tid == NT) {
                int tt = 0, c = 0;
                const double *key0 = p.brk; // first key of the current block
                const uint16_t mask = (uint16_t)((1u << csize) - 1u);
                for (int gk = 0; gk < total_tiles; gk++) {
                    const int st = gk % NSTAGE;
                    if (gk >= NSTAGE) mbar_wait(empty_s + st * 8, (uint32_t)((gk / NSTAGE - 1) & 1));
                    if (csize == 1) {
                        issue_src(key0 + (size_t)tt * p.brk_doubles + (size_t)c * N, (uint32_t)st);
                    } else if (crank != 0) {
                        mbar_expect_tx(bar_s + st * 8, TILE);                  // armed before rank 0 can send
                        mbar_arrive_remote(mapa_u32(peer_s + st * 8, 0));
                    } else {
                        mbar_wait_cluster(peer_s + st * 8, (uint32_t)((gk / NSTAGE) & 1)); // every other rank has drained and armed the stage
                        const double *src = key0 + (size_t)tt * p.brk_doubles + (size_t)c * N;
                        const uint32_t bar = bar_s + st * 8;
                        mbar_expect_tx(bar, TILE);
#pragma unroll
                        for (int r = 0; r < RT; r++) bulk_g2s_multicast(ring_s + (uint32_t)(st * RT + r) * CHUNK, src + (size_t)r * C * N, CHUNK, bar, mask);
                    }
                    if (++tt == bs) {
                        tt = 0;
                        if (++c == C) {
                            c = 0;
                            key0 += (size_t)bs * p.brk_doubles;
                        }
                    }
                }
            }
            return;
        }
    } else if (tid == 0) {
        for (int gk = 0; gk < NSTAGE && gk < total_tiles; gk++) issue_tile((uint32_t)gk, (uint32_t)gk);
    }

    const int gp = tid % GH, f = tid / GH; // this thread's frequency and its two ciphertexts gp, gp + GH
    double2 *mine0 = csm + (size_t)gp * GS + FPAD(f), *mine1 = csm + (size_t)(gp + GH) * GS + FPAD(f);
    uint32_t gk = 0;

    for (int blk = 0; blk + bs <= p.n_lwe; blk += bs) {
        if (tid < G * bs) {
            const int g = tid / bs, tt = tid % bs, ct = ct0 + g;
            const long long ai = ct < p.batch ? p.lwe[(size_t)ct * p.lwe_stride + 1 + blk + tt] : 0;
            s_pos[g * 8 + tt] = (int)((ai + (long long)(2 * N)) & (long long)(2 * N - 1));
        }
        // ---- acc_dft = FFT(acc) ------------------------------------------------------------------------------------------
        for (int base = 0; base < G * RT; base += NSLOT) {
            const int job = base + slot;
            const bool valid = job < G * RT;
            const int g = valid ? job / RT : 0, r = valid ? job % RT : 0, limb = r / cols, col = r % cols;
            double2 *buf = csm + g * GS + r * PL;
            if (valid) {
                const int ct = ct0 + g;
                const bool live = ct < p.batch && limb < p.out_size;
                const long long *src = p.res + (size_t)(live ? ct : 0) * p.res_stride + (size_t)(limb * cols + col) * N;
                double2 x[8];
#pragma unroll
                for (int jj = 0; jj < 8; jj++) {
                    const int idx = t + jj * T;
                    x[jj] = live ? make_double2((double)src[idx], (double)src[idx + M]) : make_double2(0.0, 0.0);
                }
                fct_radix8<FG::R0, true>(x, twf, 1u);
#pragma unroll
                for (int jj = 0; jj < 8; jj++) buf[FPAD(t + jj * T)] = x[jj];
            }
            poly_sync<T>(slot);
            SmFwdP<LM, FG::R0>::run(buf, twf, twlf, t, slot, valid);
        }
        named_sync(15, NT); // every transform of the block is complete before the key products read across them
        // ---- key products: tiles in (output poly, key) order; the partial sums of ONE output poly at a time live in registers -------------
        {
            double a0r[RT], a0i[RT], a1r[RT], a1i[RT];
#pragma unroll
            for (int r = 0; r < RT; r++) {
                const double2 u = mine0[r * PL], v = mine1[r * PL];
                a0r[r] = u.x; a0i[r] = u.y; a1r[r] = v.x; a1i[r] = v.y;
            }
            double w0r[BS], w0i[BS], w1r[BS], w1i[BS]; // X^{a_t} at this frequency for both ciphertexts and every key of the block
#pragma unroll
            for (int tt = 0; tt < BS; tt++) {
                if (tt < bs) {
                    const double *w0 = p.xpa + (size_t)s_pos[gp * 8 + tt] * N + f, *w1 = p.xpa + (size_t)s_pos[(gp + GH) * 8 + tt] * N + f;
                    w0r[tt] = __ldg(w0); w0i[tt] = __ldg(w0 + M); w1r[tt] = __ldg(w1); w1i[tt] = __ldg(w1 + M);
                }
            }
#pragma unroll 1
            for (int q = 0; q < C; q++) {
                double s0r = 0.0, s0i = 0.0, s1r = 0.0, s1i = 0.0;
#pragma unroll
                for (int tt = 0; tt < BS; tt++) {
                    if (tt < bs) { // uniform
                        const uint32_t st = gk % NSTAGE;
                        mbar_wait(bar_s + st * 8, (gk / NSTAGE) & 1u);
                        const double *tile = ring + (size_t)st * RT * N + f;
                        double v0r = 0.0, v0i = 0.0, v1r = 0.0, v1i = 0.0;
#pragma unroll
                        for (int r = 0; r < RT; r++) { // row order of reim4_add_mul (reim4/arithmetic_ref.rs:223-232), FMA-contracted
                            const double br = tile[r * N], bi = tile[r * N + M];
                            v0r = fma(a0r[r], br, v0r); v0r = fma(-a0i[r], bi, v0r);
                            v0i = fma(a0r[r], bi, v0i); v0i = fma(a0i[r], br, v0i);
                            v1r = fma(a1r[r], br, v1r); v1r = fma(-a1i[r], bi, v1r);
                            v1i = fma(a1r[r], bi, v1i); v1i = fma(a1i[r], br, v1i);
                        }
                        if (CG4_PRODUCER_WARP) {
                            mbar_arrive(empty_s + st * 8); // the key values are in registers: release the stage to the producer warp
                        } else {
                            __syncwarp(); // every lane's key values are in registers
                            if ((tid & 31) == 0) {
                                mbar_arrive(empty_s + st * 8);
                                // the last warp out refills the stage.  The counter only elects it; the ordering comes from the barrier:
                                // every warp arrived (release) before it counted itself, so the wait (acquire) completes at once
                                if (atomicAdd(&s_done[st], 1u) == NT / 32 - 1) {
                                    s_done[st] = 0;
                                    mbar_wait(empty_s + st * 8, (gk / NSTAGE) & 1u);
                                    if (gk + NSTAGE < (uint32_t)total_tiles) {
                                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                                        issue_tile(gk + NSTAGE, st);
                                    }
                                }
                            }
                        }
                        const double p0r = fma(w0r[tt], v0r, -(w0i[tt] * v0i)), p0i = fma(w0r[tt], v0i, w0i[tt] * v0r); // svp: reim_mul(ppol, v)
                        const double p1r = fma(w1r[tt], v1r, -(w1i[tt] * v1i)), p1i = fma(w1r[tt], v1i, w1i[tt] * v1r);
                        s0r = (s0r + p0r) - v0r; s0i = (s0i + p0i) - v0i; // dft_add_assign then dft_sub_assign, keys in block order
                        s1r = (s1r + p1r) - v1r; s1i = (s1i + p1i) - v1i;
                        gk++;
                    }
                }
                mine0[q * PL] = make_double2(s0r, s0i); // the planes of polys < RT were read into registers above: in place is safe
                mine1[q * PL] = make_double2(s1r, s1i);
            }
        }
        named_sync(15, NT);
        // ---- acc = normalize(round(IFFT(out) / m) + acc) -----------------------------------------------------------------------
        for (int base = 0; base < G * C; base += NSLOT) {
            const int job = base + slot;
            const bool valid = job < G * C;
            const int g = valid ? job / C : 0, q = valid ? job % C : 0;
            double2 *buf = csm + g * GS + q * PL;
            if (valid) {
                double2 x[8];
#pragma unroll
                for (int jj = 0; jj < 8; jj++) x[jj] = buf[FPAD(8 * t + jj)];
                double2 w[7];
                load_tw7<true>(w, twli, T, t);
                fgs_radix8_w(x, w);
#pragma unroll
                for (int jj = 0; jj < 8; jj++) buf[FPAD(8 * t + jj)] = x[jj];
            }
            poly_sync<T>(slot);
            SmInvP<LM, (LM - 6 >= FG::R0) ? LM - 6 : -1>::run(buf, twi, t, slot, valid);
            double2 x[8];
            if (valid) {
#pragma unroll
                for (int jj = 0; jj < 8; jj++) x[jj] = buf[FPAD(t + jj * T)];
                fgs_radix8<FG::R0, true>(x, twi, 1u);
            }
            poly_sync<T>(slot); // the transform has read its inputs: the buffer is reused for the rounded i64 coefficients
            if (valid) {
                long long *big = reinterpret_cast<long long *>(buf);
#pragma unroll
                for (int jj = 0; jj < 8; jj++) {
                    const int idx = t + jj * T;
                    big[idx] = (long long)round(x[jj].x * inv_m); // reim_to_znx_i64 (conversion.rs:43-52)
                    big[idx + M] = (long long)round(x[jj].y * inv_m);
                }
            }
        }
        named_sync(15, NT);
        // brk_size <= 4 here (checked on the host): all loads of a coefficient are issued before its carry chain; 32-bit offsets
        {
            const int limb_w = cols * N, bsz = p.brk_size;
            const bool two_to_one = bsz == 2 && p.out_size == 1; // the bench shape (k_brk = 2 limbs, k_glwe = 1 limb): no predicated loads
            for (int g = 0; g < G; g++) {
                if (ct0 + g >= p.batch) break;
                long long *acc_g = p.res + (size_t)(ct0 + g) * p.res_stride;
                const long long *big_g = reinterpret_cast<const long long *>(csm + (size_t)g * GS);
                if (two_to_one) {
                    for (int col = 0; col < cols; col++) {
#pragma unroll
                        for (int i = tid; i < N; i += NT) {
                            const long long v0 = big_g[col * (2 * PL) + i], v1 = big_g[(cols + col) * (2 * PL) + i];
                            long long *ap = acc_g + col * N + i;
                            const long long a0 = *ap;
                            const long long o1 = (long long)((unsigned long long)v1 << (64 - K)) >> (64 - K);
                            const long long c = (long long)((unsigned long long)v1 - (unsigned long long)o1) >> K;
                            const long long tsum = (long long)((unsigned long long)v0 + (unsigned long long)a0 + (unsigned long long)c);
                            *ap = (long long)((unsigned long long)tsum << (64 - K)) >> (64 - K);
                        }
                    }
                    continue;
                }
                for (int col = 0; col < cols; col++) {
#pragma unroll
                    for (int i = tid; i < N; i += NT) {
                        const int o = col * N + i;
                        long long vv[4], aa[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            vv[j] = j < bsz ? big_g[(j * cols + col) * (2 * PL) + i] : 0;
                            aa[j] = j < mn_small ? acc_g[o + j * limb_w] : 0;
                        }
                        long long c = 0;
#pragma unroll
                        for (int j = 3; j >= 0; j--) {
                            if (j < bsz) {
                                const long long tsum = (long long)((unsigned long long)vv[j] + (unsigned long long)aa[j] + (unsigned long long)c);
                                const long long out = (long long)((unsigned long long)tsum << (64 - K)) >> (64 - K);
                                c = (long long)((unsigned long long)tsum - (unsigned long long)out) >> K;
                                if (j < a_start) acc_g[o + j * limb_w] = out;
                            }
                        }
                        for (int j = a_start; j < p.out_size; j++) acc_g[o + j * limb_w] = 0;
                    }
                }
            }
        }
        named_sync(15, NT);
    }
}

template <int LM, int G, int RT, int CT, int NSTAGE> static int launch_cggi3(pgb_module *m, const CggiFusedArgs &p) {
    typedef FGeo<LM> FG;
    constexpr int PMAX = RT > CT ? RT : CT;
    const size_t smem = (size_t)G * (PMAX * FG::PLANE + 4) * sizeof(double2) + (size_t)NSTAGE * RT * (2 << LM) * 8 + (size_t)2 * (1 << LM) * sizeof(double2);
    PGB_CHECK_CUDA(cudaFuncSetAttribute(cggi_fused3_fft64_kernel<LM, G, RT, CT, NSTAGE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (p.batch + G - 1) / G;
    { ProfScope _ps(m, PROF_OTHER);
    cggi_fused3_fft64_kernel<LM, G, RT, CT, NSTAGE><<<grid, 512, smem, m->stream>>>(p, m->fft_fwd, m->fft_inv, m->fft_last_f, m->fft_last_i,
                                                                                   1.0 / (double)(1 << LM));
    }
    PGB_CHECK_CUDA(cudaGetLastError());
    return PGB_OK;
}
template <int LM, int G, int NSTAGE> static int launch_cggi3_shape(pgb_module *m, const CggiFusedArgs &p, int R, int C, bool *handled) {
    *handled = true;
    if (R == 4 && C > 4 && C <= 8) return launch_cggi3<LM, G, 4, 8, NSTAGE>(m, p);
    if (R == 4 && C <= 4) return launch_cggi3<LM, G, 4, 4, NSTAGE>(m, p);
    if (R == 2 && C > 4 && C <= 8) return launch_cggi3<LM, G, 2, 8, NSTAGE>(m, p);
    if (R == 2 && C <= 4) return launch_cggi3<LM, G, 2, 4, NSTAGE>(m, p);
    *handled = false;
    return PGB_OK;
}

template <int LM, int G, int RT, int CT, int NSTAGE, int BS> static int launch_cggi4(pgb_module *m, const CggiFusedArgs &p) {
    typedef FGeo<LM> FG;
    constexpr int PMAX = RT > CT ? RT : CT;
    const size_t smem = (size_t)G * (PMAX * FG::PLANE + 4) * sizeof(double2) + (size_t)NSTAGE * RT * (2 << LM) * 8 + (size_t)2 * (1 << LM) * sizeof(double2);
    static bool attr_dev[32] = {};
    if (!attr_dev[m->device & 31]) {
        PGB_CHECK_CUDA(cudaFuncSetAttribute(cggi_fused4_fft64_kernel<LM, G, RT, CT, NSTAGE, BS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_dev[m->device & 31] = true;
    }
    int grid = (p.batch + G - 1) / G;
    const int cl = (CG4_PRODUCER_WARP && m->opt[PGB_OPT_CGGI_CLUSTER] == 2 && grid >= 2) ? 2 : 1;
    grid = (grid + cl - 1) / cl * cl; // a CTA past the batch still walks the ring with its peers (its ciphertexts are not live)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(CG4_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = m->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    { ProfScope _ps(m, PROF_OTHER);
    PGB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, cggi_fused4_fft64_kernel<LM, G, RT, CT, NSTAGE, BS>, p, (const double2 *)m->fft_fwd, (const double2 *)m->fft_inv,
                                      (const double2 *)m->fft_last_f, (const double2 *)m->fft_last_i, 1.0 / (double)(1 << LM)));
    }
    PGB_CHECK_CUDA(cudaGetLastError());
    return PGB_OK;
}

template <int LM, int G, int RT, int CT> static int launch_cggi2(pgb_module *m, const CggiFusedArgs &p) {
    typedef FGeo<LM> FG;
    constexpr int PMAX = RT > CT ? RT : CT;
    const size_t smem = (size_t)G * PMAX * FG::PLANE * sizeof(double2);
    PGB_CHECK_CUDA(cudaFuncSetAttribute(cggi_fused2_fft64_kernel<LM, G, RT, CT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = (p.batch + G - 1) / G;
    { ProfScope _ps(m, PROF_OTHER);
    cggi_fused2_fft64_kernel<LM, G, RT, CT><<<grid, 512, smem, m->stream>>>(p, m->fft_fwd, m->fft_inv, 1.0 / (double)(1 << LM));
    }
    PGB_CHECK_CUDA(cudaGetLastError());
    return PGB_OK;
}
// shapes the register-blocked kernel is instantiated for: rows in {2, 4}, at most 8 output polys, at most 8 keys per block
template <int LM, int G> static int launch_cggi2_shape(pgb_module *m, const CggiFusedArgs &p, int R, int C, bool *handled) {
    *handled = true;
    if (R == 4 && C > 4 && C <= 8) return launch_cggi2<LM, G, 4, 8>(m, p);
    if (R == 4 && C <= 4) return launch_cggi2<LM, G, 4, 4>(m, p);
    if (R == 2 && C > 4 && C <= 8) return launch_cggi2<LM, G, 2, 8>(m, p);
    if (R == 2 && C <= 4) return launch_cggi2<LM, G, 2, 4>(m, p);
    *handled = false;
    return PGB_OK;
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
    if (block_size <= 3 && m->opt[PGB_OPT_CGGI_VARIANT] == 0 && (brk_doubles % 2) == 0 && brk_size <= 4 && m->log_n >= 8 && m->log_n <= 10 &&
        (R == 4 || R == 2) && C <= 8) {
        // version 4 (dedicated ring warp): n = 256 / 512 / 1024 (8 / 4 / 2 ciphertexts per CTA), four (rank 3) or two (rank 1) input polys,
        // blocks of at most three keys -- the BASELINE family
#define CGGI4_SHAPES(LM, G, NS)                                                                                          \
        if (R == 4) return C > 4 ? launch_cggi4<LM, G, 4, 8, NS, 3>(m, p) : launch_cggi4<LM, G, 4, 4, NS, 3>(m, p);       \
        return C > 4 ? launch_cggi4<LM, G, 2, 8, NS, 3>(m, p) : launch_cggi4<LM, G, 2, 4, NS, 3>(m, p);
        if (m->log_n == 8) { CGGI4_SHAPES(7, 8, 4) }
        if (m->log_n == 9) { CGGI4_SHAPES(8, 4, 4) }
        { CGGI4_SHAPES(9, 2, 2) }
#undef CGGI4_SHAPES
    }
    if (block_size <= 8 && (m->opt[PGB_OPT_CGGI_VARIANT] == 0 || m->opt[PGB_OPT_CGGI_VARIANT] >= 3) && (brk_doubles % 2) == 0 && brk_size <= 4) {
        // TMA key stream (tiles must be 16-byte aligned: brk is a cudaMalloc'd / 64-byte aligned buffer of whole polys)
        bool handled = false;
        int s = PGB_OK;
        switch (m->log_n) {
        case 8: s = launch_cggi3_shape<7, 8, 4>(m, p, R, C, &handled); break;
        case 9: s = launch_cggi3_shape<8, 4, 4>(m, p, R, C, &handled); break;
        case 10: s = launch_cggi3_shape<9, 2, 2>(m, p, R, C, &handled); break;
        default: break;
        }
        if (handled) return s;
    }
    if (block_size <= 8 && m->opt[PGB_OPT_CGGI_VARIANT] != 1) {
        bool handled = false;
        int s = PGB_OK;
        switch (m->log_n) {
        case 8: s = launch_cggi2_shape<7, 8>(m, p, R, C, &handled); break;
        case 9: s = launch_cggi2_shape<8, 4>(m, p, R, C, &handled); break;
        case 10: s = launch_cggi2_shape<9, 2>(m, p, R, C, &handled); break;
        case 11: s = launch_cggi2_shape<10, 1>(m, p, R, C, &handled); break;
        default: break;
        }
        if (handled) return s;
    }
    switch (m->log_n) {
    case 8: return launch_cggi_rt<7, 8>(m, p, R, C);
    case 9: return launch_cggi_rt<8, 4>(m, p, R, C);
    case 10: return launch_cggi_rt<9, 2>(m, p, R, C);
    case 11: return launch_cggi_rt<10, 1>(m, p, R, C);
    default: pgb_set_error("cggi fused: unsupported n"); return PGB_ERR_UNSUPPORTED;
    }
}
