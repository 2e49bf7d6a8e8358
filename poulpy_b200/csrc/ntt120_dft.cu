// ntt120_dft.cu -- batched forward / inverse negacyclic NTT of the NTT120 flavour (K1 / K2 of SURVEY.md 2.8).
//
// What is computed (reference: poulpy-cpu-ref/src/reference/ntt120/ntt.rs:558-684, vec_znx_dft.rs:177-409):
//   forward : i64 limb -> four planes of canonical residues, out_k[p] = sum_i a_i * psi_k^(i * (2*bitrev(p) + 1)),
//             i.e. the reference's bit-reversed frequency order (its DIF network after the psi^i twist);
//   inverse : planes -> centred i128 via CRT (arithmetic.rs:119-140), 1/n folded into the CRT constant.
// How: one CTA per limb (all four primes), the polynomial lives in padded shared memory, radix-8 register passes
// (three butterfly levels per shared-memory round trip) in the twist-free "block twiddle" Cooley-Tukey form
// (forward) / Gentleman-Sande form (inverse), Shoup multiplication with Harvey lazy reduction.  Global loads and
// stores are fused into the first / last pass and are fully coalesced (i64 in: 8 B/thread contiguous per warp;
// planes out: 32 B/thread; i128 out: 16 B/thread contiguous per warp).
#include "internal.h"
#include "ntt120.cuh"

using namespace n120;

__constant__ CrtConsts c_crt;

// conflict-free padding for strides 1 (128-bit), 8, 64, 512: 4 extra words per 32
__device__ __forceinline__ int PAD(int idx) { return idx + ((idx >> 5) << 2); }
template <int L> struct Geo {
    static constexpr int NB = 1 << L;
    static constexpr int T = NB >= 8 ? NB / 8 : 1;         // threads per limb, 8 coefficients each
    static constexpr int R0 = (L % 3 == 0) ? 3 : (L % 3);  // levels of the (possibly short) top pass
    static constexpr int PLANE = NB + (NB >> 5) * 4 + 4;   // padded words per prime plane
};

template <int K, int NLEV> __device__ __forceinline__ void ct_radix8(uint32_t (&x)[8], const uint2 *__restrict__ tw, uint32_t hi) {
    constexpr uint32_t q = Prime<K>::q;
    {
        uint2 w = __ldg(tw + hi);
#pragma unroll
        for (int j = 0; j < 4; j++) ct_bf(x[j], x[j + 4], w, q);
    }
    if (NLEV >= 2) {
        uint2 w0 = __ldg(tw + 2 * hi), w1 = __ldg(tw + 2 * hi + 1);
        ct_bf(x[0], x[2], w0, q);
        ct_bf(x[1], x[3], w0, q);
        ct_bf(x[4], x[6], w1, q);
        ct_bf(x[5], x[7], w1, q);
    }
    if (NLEV >= 3) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint2 w = __ldg(tw + 4 * hi + j);
            ct_bf(x[2 * j], x[2 * j + 1], w, q);
        }
    }
}
template <int K, int NLEV> __device__ __forceinline__ void gs_radix8(uint32_t (&x)[8], const uint2 *__restrict__ tw, uint32_t hi) {
    constexpr uint32_t q = Prime<K>::q;
    if (NLEV >= 3) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint2 w = __ldg(tw + 4 * hi + j);
            gs_bf(x[2 * j], x[2 * j + 1], w, q);
        }
    }
    if (NLEV >= 2) {
        uint2 w0 = __ldg(tw + 2 * hi), w1 = __ldg(tw + 2 * hi + 1);
        gs_bf(x[0], x[2], w0, q);
        gs_bf(x[1], x[3], w0, q);
        gs_bf(x[4], x[6], w1, q);
        gs_bf(x[5], x[7], w1, q);
    }
    {
        uint2 w = __ldg(tw + hi);
#pragma unroll
        for (int j = 0; j < 4; j++) gs_bf(x[j], x[j + 4], w, q);
    }
}

struct NttJobs {
    LimbSet in, out;
    int jobs_per_batch;
    int total_jobs;
};

// ---------------------------------------------------------------------------------------------- forward
template <int K, int L> __device__ __forceinline__ void fwd_prime(const long long (&v)[8], uint32_t *__restrict__ plane,
                                                                 uint32_t *__restrict__ gout, const uint2 *__restrict__ tw,
                                                                 int t, bool active) {
    typedef Geo<L> G;
    constexpr uint32_t q = Prime<K>::q;
    uint32_t x[8];
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = from_i64<K>(v[j]);
    ct_radix8<K, G::R0>(x, tw, 1u);
    if (L == G::R0) { // n == 8: the top pass is the whole transform
        if (active) {
#pragma unroll
            for (int j = 0; j < 8; j++) gout[j] = csub(csub(x[j], 2 * q), q);
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < 8; j++) plane[PAD(t + j * G::T)] = x[j];
}

template <int K, int L, int L0> __device__ __forceinline__ void fwd_mid(uint32_t *__restrict__ plane, uint32_t *__restrict__ gout,
                                                                       const uint2 *__restrict__ tw, int t, bool active) {
    constexpr uint32_t q = Prime<K>::q;
    constexpr int SL = L - L0 - 3; // log2 of the in-group stride
    const int a = t >> SL, b = t & ((1 << SL) - 1);
    const int base = (a << (SL + 3)) | b;
    const uint32_t hi = (1u << L0) | (uint32_t)a;
    uint32_t x[8];
    if (SL == 0) {
        const uint4 *p = reinterpret_cast<const uint4 *>(plane + PAD(base));
        uint4 u0 = p[0], u1 = p[1];
        x[0] = u0.x; x[1] = u0.y; x[2] = u0.z; x[3] = u0.w;
        x[4] = u1.x; x[5] = u1.y; x[6] = u1.z; x[7] = u1.w;
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) x[j] = plane[PAD(base + (j << SL))];
    }
    ct_radix8<K, 3>(x, tw, hi);
    if (SL == 0) {
        if (active) {
#pragma unroll
            for (int j = 0; j < 8; j++) x[j] = csub(csub(x[j], 2 * q), q);
            uint4 *o = reinterpret_cast<uint4 *>(gout + base);
            o[0] = make_uint4(x[0], x[1], x[2], x[3]);
            o[1] = make_uint4(x[4], x[5], x[6], x[7]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) plane[PAD(base + (j << SL))] = x[j];
    }
}

template <int L, int L0> struct FwdMid {
    static __device__ __forceinline__ void run(uint32_t *sm, uint32_t *gout, const uint2 *tw, int n, int t, bool active) {
        typedef Geo<L> G;
        fwd_mid<0, L, L0>(sm + 0 * G::PLANE, gout + 0 * n, tw + 0 * n, t, active);
        fwd_mid<1, L, L0>(sm + 1 * G::PLANE, gout + 1 * n, tw + 1 * n, t, active);
        fwd_mid<2, L, L0>(sm + 2 * G::PLANE, gout + 2 * n, tw + 2 * n, t, active);
        fwd_mid<3, L, L0>(sm + 3 * G::PLANE, gout + 3 * n, tw + 3 * n, t, active);
        if (L0 + 3 < L) __syncthreads();
        FwdMid<L, (L0 + 3 < L) ? L0 + 3 : L>::run(sm, gout, tw, n, t, active);
    }
};
template <int L> struct FwdMid<L, L> {
    static __device__ __forceinline__ void run(uint32_t *, uint32_t *, const uint2 *, int, int, bool) {}
};

template <int L, int LPC> __global__ void __launch_bounds__(Geo<L>::T *LPC) ntt120_fwd_kernel(NttJobs jb, const uint2 *__restrict__ tw) {
    typedef Geo<L> G;
    extern __shared__ __align__(16) uint32_t smem[];
    const int slot = threadIdx.x / G::T, t = threadIdx.x % G::T;
    const int job = blockIdx.x * LPC + slot;
    const bool active = job < jb.total_jobs;
    const int b = active ? job / jb.jobs_per_batch : 0, j = active ? job % jb.jobs_per_batch : 0;
    const long long *gin = reinterpret_cast<const long long *>(jb.in.base + (size_t)b * jb.in.batch_stride + (size_t)j * jb.in.limb_stride);
    uint32_t *gout = reinterpret_cast<uint32_t *>(jb.out.base + (size_t)b * jb.out.batch_stride + (size_t)j * jb.out.limb_stride);
    uint32_t *sm = smem + slot * 4 * G::PLANE;
    constexpr int n = G::NB;

    long long v[8];
#pragma unroll
    for (int jj = 0; jj < 8; jj++) v[jj] = active ? __ldg(gin + t + jj * G::T) : 0;
    fwd_prime<0, L>(v, sm + 0 * G::PLANE, gout + 0 * n, tw + 0 * n, t, active);
    fwd_prime<1, L>(v, sm + 1 * G::PLANE, gout + 1 * n, tw + 1 * n, t, active);
    fwd_prime<2, L>(v, sm + 2 * G::PLANE, gout + 2 * n, tw + 2 * n, t, active);
    fwd_prime<3, L>(v, sm + 3 * G::PLANE, gout + 3 * n, tw + 3 * n, t, active);
    if (L > G::R0) {
        __syncthreads();
        FwdMid<L, (L > G::R0) ? G::R0 : L>::run(sm, gout, tw, n, t, active);
    }
}

// ---------------------------------------------------------------------------------------------- inverse
// bottom pass: levels L-3..L-1, inputs read from global planes (8 consecutive residues per thread)
template <int K, int L> __device__ __forceinline__ void inv_bottom(uint32_t *__restrict__ plane, const uint32_t *__restrict__ gin,
                                                                  const uint2 *__restrict__ tw, int t, bool active) {
    constexpr int L0 = L - 3;
    const uint32_t hi = (1u << L0) | (uint32_t)t;
    uint32_t x[8];
    uint4 u0 = make_uint4(0, 0, 0, 0), u1 = u0;
    if (active) {
        const uint4 *p = reinterpret_cast<const uint4 *>(gin + 8 * t);
        u0 = __ldg(p);
        u1 = __ldg(p + 1);
    }
    x[0] = u0.x; x[1] = u0.y; x[2] = u0.z; x[3] = u0.w;
    x[4] = u1.x; x[5] = u1.y; x[6] = u1.z; x[7] = u1.w;
    gs_radix8<K, 3>(x, tw, hi);
    uint4 *o = reinterpret_cast<uint4 *>(plane + PAD(8 * t));
    o[0] = make_uint4(x[0], x[1], x[2], x[3]);
    o[1] = make_uint4(x[4], x[5], x[6], x[7]);
}
template <int K, int L, int L0> __device__ __forceinline__ void inv_mid(uint32_t *__restrict__ plane, const uint2 *__restrict__ tw, int t) {
    constexpr int SL = L - L0 - 3;
    const int a = t >> SL, b = t & ((1 << SL) - 1);
    const int base = (a << (SL + 3)) | b;
    const uint32_t hi = (1u << L0) | (uint32_t)a;
    uint32_t x[8];
#pragma unroll
    for (int j = 0; j < 8; j++) x[j] = plane[PAD(base + (j << SL))];
    gs_radix8<K, 3>(x, tw, hi);
#pragma unroll
    for (int j = 0; j < 8; j++) plane[PAD(base + (j << SL))] = x[j];
}
// passes from l0 = L-6 down to R0 (exclusive of the top pass)
template <int L, int L0> struct InvMid {
    static __device__ __forceinline__ void run(uint32_t *sm, const uint2 *tw, int n, int t) {
        typedef Geo<L> G;
        inv_mid<0, L, L0>(sm + 0 * G::PLANE, tw + 0 * n, t);
        inv_mid<1, L, L0>(sm + 1 * G::PLANE, tw + 1 * n, t);
        inv_mid<2, L, L0>(sm + 2 * G::PLANE, tw + 2 * n, t);
        inv_mid<3, L, L0>(sm + 3 * G::PLANE, tw + 3 * n, t);
        __syncthreads();
        InvMid<L, (L0 - 3 >= G::R0) ? L0 - 3 : -1>::run(sm, tw, n, t);
    }
};
template <int L> struct InvMid<L, -1> {
    static __device__ __forceinline__ void run(uint32_t *, const uint2 *, int, int) {}
};

// top pass for one prime: R0 levels, then t_k = x * CRT_k / n and accumulation of t_k * (Q / Q_k)
template <int K, int L> __device__ __forceinline__ void inv_top(const uint32_t *__restrict__ plane, const uint32_t *__restrict__ gin,
                                                               const uint2 *__restrict__ tw, int t, bool active,
                                                               const Ntt120Consts &nc, u128 (&acc)[8]) {
    typedef Geo<L> G;
    constexpr uint32_t q = Prime<K>::q;
    uint32_t x[8];
    if (L == G::R0) {
#pragma unroll
        for (int j = 0; j < 8; j++) x[j] = active ? __ldg(gin + j) : 0;
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) x[j] = plane[PAD(t + j * G::T)];
    }
    gs_radix8<K, G::R0>(x, tw, 1u);
    const unsigned long long mlo = c_crt.m_lo[K], mhi = c_crt.m_hi[K];
#pragma unroll
    for (int j = 0; j < 8; j++) {
        uint32_t tk = csub(mul_shoup(x[j], nc.crt_ninv[K], nc.crt_ninv_sh[K], q), q);
        u128 p = (u128)tk * mlo + ((u128)((unsigned long long)tk * mhi) << 64);
        acc[j] += p;
    }
}

__device__ __forceinline__ i128 crt_finish(u128 v) {
    const u128 Q = ((u128)c_crt.q_hi << 64) | c_crt.q_lo;
    const u128 H = ((u128)c_crt.half_hi << 64) | c_crt.half_lo;
    unsigned qa = (unsigned)(v >> 120);
    v -= (u128)qa * Q;
    if (v >= Q) v -= Q;
    return v >= H ? (i128)v - (i128)Q : (i128)v;
}

// OUT_I128 = true : write centred i128 coefficients (VecZnxBig of the NTT120 flavour)
template <int L, int LPC> __global__ void __launch_bounds__(Geo<L>::T *LPC) ntt120_inv_kernel(NttJobs jb, const uint2 *__restrict__ tw,
                                                                                             Ntt120Consts nc) {
    typedef Geo<L> G;
    extern __shared__ __align__(16) uint32_t smem[];
    const int slot = threadIdx.x / G::T, t = threadIdx.x % G::T;
    const int job = blockIdx.x * LPC + slot;
    const bool active = job < jb.total_jobs;
    const int b = active ? job / jb.jobs_per_batch : 0, j = active ? job % jb.jobs_per_batch : 0;
    const uint32_t *gin = reinterpret_cast<const uint32_t *>(jb.in.base + (size_t)b * jb.in.batch_stride + (size_t)j * jb.in.limb_stride);
    i128 *gout = reinterpret_cast<i128 *>(jb.out.base + (size_t)b * jb.out.batch_stride + (size_t)j * jb.out.limb_stride);
    uint32_t *sm = smem + slot * 4 * G::PLANE;
    constexpr int n = G::NB;

    if (L > G::R0) {
        inv_bottom<0, L>(sm + 0 * G::PLANE, gin + 0 * n, tw + 0 * n, t, active);
        inv_bottom<1, L>(sm + 1 * G::PLANE, gin + 1 * n, tw + 1 * n, t, active);
        inv_bottom<2, L>(sm + 2 * G::PLANE, gin + 2 * n, tw + 2 * n, t, active);
        inv_bottom<3, L>(sm + 3 * G::PLANE, gin + 3 * n, tw + 3 * n, t, active);
        __syncthreads();
        InvMid<L, (L - 6 >= G::R0) ? L - 6 : -1>::run(sm, tw, n, t);
    }
    u128 acc[8];
#pragma unroll
    for (int jj = 0; jj < 8; jj++) acc[jj] = 0;
    inv_top<0, L>(sm + 0 * G::PLANE, gin + 0 * n, tw + 0 * n, t, active, nc, acc);
    inv_top<1, L>(sm + 1 * G::PLANE, gin + 1 * n, tw + 1 * n, t, active, nc, acc);
    inv_top<2, L>(sm + 2 * G::PLANE, gin + 2 * n, tw + 2 * n, t, active, nc, acc);
    inv_top<3, L>(sm + 3 * G::PLANE, gin + 3 * n, tw + 3 * n, t, active, nc, acc);
    if (L == G::R0) __syncthreads(); // n == 8, in-place safety: every thread has read before anyone writes
    if (active) {
#pragma unroll
        for (int jj = 0; jj < 8; jj++) gout[t + jj * G::T] = crt_finish(acc[jj]);
    }
}

// ---------------------------------------------------------------------------------------------- host side
static uint32_t modpow(uint32_t x, uint64_t e, uint32_t q) {
    uint64_t r = 1, b = x;
    while (e) {
        if (e & 1) r = r * b % q;
        b = b * b % q;
        e >>= 1;
    }
    return (uint32_t)r;
}

int ntt120_module_init(pgb_module *m) {
    const uint64_t n = m->n;
    // block-twiddle exponents: E[1] = n/2, E[2i] = E[i]/2, E[2i+1] = E[i]/2 + n/2  (W[i] = psi^E[i], psi = 2n-th root)
    uint64_t *E = (uint64_t *)malloc(sizeof(uint64_t) * n);
    if (n >= 2) E[1] = n / 2;
    for (uint64_t i = 1; 2 * i + 1 < n; i++) {
        E[2 * i] = E[i] / 2;
        E[2 * i + 1] = E[i] / 2 + n / 2;
    }
    uint2 *hf = (uint2 *)calloc(4 * n, sizeof(uint2)), *hi = (uint2 *)calloc(4 * n, sizeof(uint2));
    for (int k = 0; k < 4; k++) {
        uint32_t q = qk(k);
        uint32_t psi = modpow(OMEGA[k], (1u << 16) / n, q); // ntt.rs:164-167
        for (uint64_t i = 1; i < n; i++) {
            uint32_t w = modpow(psi, E[i], q);
            uint32_t wi = modpow(psi, 2 * n - E[i], q);
            hf[k * n + i] = make_uint2(w, (uint32_t)(((uint64_t)w << 32) / q));
            hi[k * n + i] = make_uint2(wi, (uint32_t)(((uint64_t)wi << 32) / q));
        }
        uint32_t ninv = modpow((uint32_t)(n % q), q - 2, q);
        uint32_t c = (uint32_t)((uint64_t)CRT_CST[k] * ninv % q);
        m->nc.crt_ninv[k] = c;
        m->nc.crt_ninv_sh[k] = (uint32_t)(((uint64_t)c << 32) / q);
    }
    free(E);
    PGB_CHECK_CUDA(cudaMalloc(&m->ntt_fwd, 4 * n * sizeof(uint2)));
    PGB_CHECK_CUDA(cudaMalloc(&m->ntt_inv, 4 * n * sizeof(uint2)));
    PGB_CHECK_CUDA(cudaMemcpy(m->ntt_fwd, hf, 4 * n * sizeof(uint2), cudaMemcpyHostToDevice));
    PGB_CHECK_CUDA(cudaMemcpy(m->ntt_inv, hi, 4 * n * sizeof(uint2), cudaMemcpyHostToDevice));
    free(hf);
    free(hi);
    CrtConsts cc;
    u128 Q = 1;
    for (int k = 0; k < 4; k++) Q *= qk(k);
    for (int k = 0; k < 4; k++) {
        u128 mk = Q / qk(k);
        cc.m_lo[k] = (unsigned long long)mk;
        cc.m_hi[k] = (unsigned long long)(mk >> 64);
    }
    cc.q_lo = (unsigned long long)Q;
    cc.q_hi = (unsigned long long)(Q >> 64);
    u128 H = (Q + 1) / 2;
    cc.half_lo = (unsigned long long)H;
    cc.half_hi = (unsigned long long)(H >> 64);
    PGB_CHECK_CUDA(cudaMemcpyToSymbol(c_crt, &cc, sizeof cc));
    return PGB_OK;
}

template <int L> static constexpr int lpc_for() { return Geo<L>::T >= 128 ? 1 : (128 / Geo<L>::T > 16 ? 16 : 128 / Geo<L>::T); }

template <int L> static int launch_fwd(pgb_module *m, const NttJobs &jb) {
    constexpr int LPC = lpc_for<L>();
    typedef Geo<L> G;
    size_t smem = (size_t)LPC * 4 * G::PLANE * sizeof(uint32_t);
    static bool attr_set = false;
    if (!attr_set) {
        PGB_CHECK_CUDA(cudaFuncSetAttribute(ntt120_fwd_kernel<L, LPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    int grid = (jb.total_jobs + LPC - 1) / LPC;
    { ProfScope _ps(m, PROF_DFT_FWD);
    ntt120_fwd_kernel<L, LPC><<<grid, G::T * LPC, smem, m->stream>>>(jb, m->ntt_fwd);
    }
    PGB_CHECK_CUDA(cudaGetLastError());
    return PGB_OK;
}
template <int L> static int launch_inv(pgb_module *m, const NttJobs &jb) {
    constexpr int LPC = lpc_for<L>();
    typedef Geo<L> G;
    size_t smem = (size_t)LPC * 4 * G::PLANE * sizeof(uint32_t);
    static bool attr_set = false;
    if (!attr_set) {
        PGB_CHECK_CUDA(cudaFuncSetAttribute(ntt120_inv_kernel<L, LPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    int grid = (jb.total_jobs + LPC - 1) / LPC;
    { ProfScope _ps(m, PROF_DFT_INV);
    ntt120_inv_kernel<L, LPC><<<grid, G::T * LPC, smem, m->stream>>>(jb, m->ntt_inv, m->nc);
    }
    PGB_CHECK_CUDA(cudaGetLastError());
    return PGB_OK;
}

#define NTT_DISPATCH(fn)                         \
    switch (m->log_n) {                          \
    case 3: return fn<3>(m, jb);                 \
    case 4: return fn<4>(m, jb);                 \
    case 5: return fn<5>(m, jb);                 \
    case 6: return fn<6>(m, jb);                 \
    case 7: return fn<7>(m, jb);                 \
    case 8: return fn<8>(m, jb);                 \
    case 9: return fn<9>(m, jb);                 \
    case 10: return fn<10>(m, jb);               \
    case 11: return fn<11>(m, jb);               \
    case 12: return fn<12>(m, jb);               \
    case 13: return fn<13>(m, jb);               \
    default:                                     \
        pgb_set_error("NTT120: n = 2^%d not supported by the single-CTA path (8 <= n <= 8192)", m->log_n); \
        return PGB_ERR_UNSUPPORTED;              \
    }

// in: i64 limbs, out: 16 B/coef DFT limbs
int ntt120_forward(pgb_module *m, LimbSet in, LimbSet out, int jobs_per_batch, int batch) {
    NttJobs jb = {in, out, jobs_per_batch, jobs_per_batch * batch};
    if (jb.total_jobs == 0) return PGB_OK;
    NTT_DISPATCH(launch_fwd)
}
// in: DFT limbs, out: i128 limbs (may alias `in` limb for limb: in-place consume)
int ntt120_inverse_big(pgb_module *m, LimbSet in, LimbSet out, int jobs_per_batch, int batch) {
    NttJobs jb = {in, out, jobs_per_batch, jobs_per_batch * batch};
    if (jb.total_jobs == 0) return PGB_OK;
    NTT_DISPATCH(launch_inv)
}
