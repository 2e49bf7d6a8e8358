// ntt120_gadget.cu -- the whole gadget product (GLWE key-switch / GGSW x GLWE external product, dsize = 1, same base2k) of the
// NTT120 flavour as ONE persistent kernel: i64 GLWE in -> i64 GLWE out, nothing else touches HBM.
//
// Restates, per ciphertext, the HAL sequence of poulpy-core/src/keyswitching/glwe.rs:207-239,106-108 and
// external_product/glwe.rs:197-271,138-140:
//     vec_znx_dft_apply (R limbs) -> vmp_apply_dft_to_dft -> vec_znx_idft_apply_consume -> vec_znx_big_add_small_assign
//     -> vec_znx_big_normalize (same base2k, offset 0)
// using the collapsed key of ntt120_dft.cu (one inverse transform per output column; bit-identical to the per-limb route whenever
// the integers stay below 2^118, which is decided per ciphertext on the device -- the others are flagged for the per-limb kernels).
//
// B200 mapping: a thread-block CLUSTER of four CTAs owns one ciphertext, CTA k works modulo Q[k] (primes.rs:80-90):
//   * max(R, cols_out) polynomials live in swizzled (unpadded) shared memory, 16 coefficients per thread;
//   * radix-16 register passes (four butterfly levels per shared-memory round trip), Shoup multiplication, Harvey lazy ranges;
//   * the R forward transforms run in lock step so that the per-thread twiddles are fetched once per pass;
//   * products with the collapsed key accumulate in u64 and are reduced once;
//   * CRT needs all four residues of a coefficient: CTA k reconstructs the k-th quarter of the coefficients and reads the other
//     three residues from its peers through distributed shared memory (ld.shared::cluster), then emits the base-2^K digits.
// Four independent CTAs (of different clusters) share an SM, so global loads, barriers and the CRT tail of one ciphertext overlap
// with the butterflies of another.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "internal.h"
#include "ntt120.cuh"

namespace cg = cooperative_groups;
using namespace n120;

namespace {

struct GadgetArgs {
    const char *in;  unsigned long long in_bs;      // GLWE inputs (i64), limb (j, col) at ((j * in_cols + col) * n) words
    char *res;       unsigned long long res_bs;     // GLWE outputs (i64), limb (j, col) at ((j * cols_out + col) * n) words
    const uint32_t *ckey;                           // collapsed key [r][col][k][chunk 0..3][t][4 words]
    const int *key_bits;                            // device: max bit length of the key's integer coefficients
    int *ok;                                        // out, per ciphertext: 1 = done here, 0 = needs the per-limb route
    int in_cols, row_cols, row_col0, R, cols_out;
    int small_size;                                 // limbs of input column 0 added to output column 0 (key-switch), 0 = none
    int K, S, res_size;                             // base2k, key size (digits of the collapsed integer), output limbs
    int bound_bits;                                 // an input passes when its bit length <= bound_bits - key_bits (see ntt120_gadget_fused)
    int batch;
    // automorphism epilogue (AUT instances only): output coefficient j' takes source coefficient j = j' * aut_pinv mod 2n (sign flipped
    // when that product lands in [n, 2n)).  x = CRT value + body.  mode 1: res = normalize(aut(x) + a), 2: normalize(aut(x) - a),
    // 3: normalize(a - aut(x)) with a = limbs of the INPUT ciphertext (all columns, the first post_size limbs) -- the limb-wise
    // big_automorphism / big_(add|sub)_small / big_normalize sequence of automorphism/glwe_ct.rs:95-275; mode 4: res = aut(normalize(x))
    // (glwe_automorphism, glwe_ct.rs:51-72: the digits are permuted and negated after the normalisation)
    int aut_mode, post_size;
    uint32_t aut_pinv;
    uint32_t zero;                                  // always 0 (see ct_bfz)
    // CRT over the NP primes the launch works with (Q = their product): NP = 4 is the reference's Q120 (arithmetic.rs:119-140), NP = 3
    // the same integers reconstructed from three residues when they are known to stay below Q[0] Q[1] Q[2] / 2
    uint32_t m_w[4][4];                             // M_k = Q / Q[k], four 32-bit words each
    uint32_t nq_w[4];                               // 2^128 - Q
    unsigned long long half_lo, half_hi;            // sum_j 2^(K-1) 2^(jK), j < S (mod 2^128)
    uint32_t inv60[4];                              // floor(2^60 / Q[k]) (31 bits)
    uint32_t crt_ninv[4], crt_ninv_sh[4];           // (Q / Q[k])^-1 / n mod Q[k] (CRT_CST[k] / n for NP = 4) and its Shoup companion
    struct { uint32_t q, c32, c32s, neg64, qni; } prime[4]; // qni = -q^-1 mod 2^32 (Montgomery reduction of the key products)
};

template <int L> struct GGeo {
    static constexpr int N = 1 << L;
    static constexpr int T = N / 16;
    // pass p covers butterfly levels [LV + 4 - NLEV, LV + 4); in-group stride 2^SG
    static constexpr int LV2 = (L >= 12) ? 4 : (L == 11 ? 4 : 3);
    static constexpr int NL2 = (L >= 11) ? 4 : 3;
    static constexpr int NL3 = L - (LV2 + 4) ;  // levels left for the last pass
    static constexpr int SG2 = L - LV2 - 4;     // log2 stride of pass 2
    static constexpr bool SIG3 = SG2 == 3;
};

// conflict-free XOR swizzle of a word index for the three access patterns (stride T scalar, stride 2^SG2 scalar, 16 consecutive
// words as four 128-bit accesses); only bits >= 2 change, so aligned quads stay contiguous
template <int L> __device__ __forceinline__ int swz(int i) {
    int r = i ^ (((i >> 5) & 3) << 2);
    if (GGeo<L>::SIG3) r ^= ((i >> 7) & 3) << 3;
    else r ^= ((i >> 8) & 1) << 4;
    return r;
}

// Butterflies as in ntt120.cuh, with one twist: `z` is a kernel parameter that is always 0 but unknown to the compiler.  ptxas turns
// two-input integer adds into IMAD.IADD on the FMA-heavy pipe, which the three IMADs of the Shoup product already saturate inside the
// butterfly sections; a three-input add can only be an IADD3 on the ALU pipe.
__device__ __forceinline__ void ct_bfz(uint32_t &x, uint32_t &y, uint2 w, uint32_t q, uint32_t z) {
    const uint32_t xr = csub(x, 2 * q);
    const uint32_t t = mul_shoup(y, w.x, w.y, q);
    x = xr + t + z;
    y = xr - t + 2 * q;
}
__device__ __forceinline__ void gs_bfz(uint32_t &x, uint32_t &y, uint2 w, uint32_t q, uint32_t z) {
    const uint32_t s = csub(x + y + z, 2 * q);
    const uint32_t d = x - y + 2 * q;
    x = s;
    y = mul_shoup(d, w.x, w.y, q);
}

// ---- 16-point butterfly networks; level j of the network has distance 8 >> j; NLEV < 4 skips the first 4 - NLEV levels ------------
template <int NLEV, class TW> __device__ __forceinline__ void ct16(uint32_t (&x)[16], const TW &tw, const uint32_t q, const uint32_t z) {
    if (NLEV >= 4) {
        const uint2 w = tw.get1();
#pragma unroll
        for (int j = 0; j < 8; j++) ct_bfz(x[j], x[j + 8], w, q, z);
    }
    if (NLEV >= 3) {
        uint2 w[2];
        tw.get2(w);
#pragma unroll
        for (int b = 0; b < 2; b++)
#pragma unroll
            for (int j = 0; j < 4; j++) ct_bfz(x[8 * b + j], x[8 * b + j + 4], w[b], q, z);
    }
    if (NLEV >= 2) {
        uint2 w[4];
        tw.get4(w);
#pragma unroll
        for (int b = 0; b < 4; b++)
#pragma unroll
            for (int j = 0; j < 2; j++) ct_bfz(x[4 * b + j], x[4 * b + j + 2], w[b], q, z);
    }
    {
        uint2 w[8];
        tw.get8(w);
#pragma unroll
        for (int b = 0; b < 8; b++) ct_bfz(x[2 * b], x[2 * b + 1], w[b], q, z);
    }
}
template <int NLEV, class TW> __device__ __forceinline__ void gs16(uint32_t (&x)[16], const TW &tw, const uint32_t q, const uint32_t z) {
    {
        uint2 w[8];
        tw.get8(w);
#pragma unroll
        for (int b = 0; b < 8; b++) gs_bfz(x[2 * b], x[2 * b + 1], w[b], q, z);
    }
    if (NLEV >= 2) {
        uint2 w[4];
        tw.get4(w);
#pragma unroll
        for (int b = 0; b < 4; b++)
#pragma unroll
            for (int j = 0; j < 2; j++) gs_bfz(x[4 * b + j], x[4 * b + j + 2], w[b], q, z);
    }
    if (NLEV >= 3) {
        uint2 w[2];
        tw.get2(w);
#pragma unroll
        for (int b = 0; b < 2; b++)
#pragma unroll
            for (int j = 0; j < 4; j++) gs_bfz(x[8 * b + j], x[8 * b + j + 4], w[b], q, z);
    }
    if (NLEV >= 4) {
        const uint2 w = tw.get1();
#pragma unroll
        for (int j = 0; j < 8; j++) gs_bfz(x[j], x[j + 8], w, q, z);
    }
}

// twiddles of tree node `hi` and its descendants from the [n] table of one prime (block-twiddle order, ntt120_dft.cu)
struct TwGlobal {
    const uint2 *tw;
    uint32_t hi;
    __device__ __forceinline__ uint2 get1() const { return __ldg(tw + hi); }
    __device__ __forceinline__ void get2(uint2 (&w)[2]) const {
        const uint4 a = __ldg(reinterpret_cast<const uint4 *>(tw + 2 * hi));
        w[0] = make_uint2(a.x, a.y); w[1] = make_uint2(a.z, a.w);
    }
    __device__ __forceinline__ void get4(uint2 (&w)[4]) const {
        const uint4 *p = reinterpret_cast<const uint4 *>(tw + 4 * hi);
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const uint4 a = __ldg(p + i);
            w[2 * i] = make_uint2(a.x, a.y); w[2 * i + 1] = make_uint2(a.z, a.w);
        }
    }
    __device__ __forceinline__ void get8(uint2 (&w)[8]) const {
        const uint4 *p = reinterpret_cast<const uint4 *>(tw + 8 * hi);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint4 a = __ldg(p + i);
            w[2 * i] = make_uint2(a.x, a.y); w[2 * i + 1] = make_uint2(a.z, a.w);
        }
    }
};
// the same 15 twiddles from the per-thread table [8][T] uint4 (common.cuh: ntt_last16_*): every load of a warp is one contiguous 512 bytes,
// where the 64-byte-strided get8 / get4 of TwGlobal touch four / two times the lines they need
struct TwLast {
    const uint4 *p; // table of this prime + t
    int T;
    __device__ __forceinline__ uint2 get1() const {
        const uint4 a = __ldg(p);
        return make_uint2(a.x, a.y);
    }
    __device__ __forceinline__ void get2(uint2 (&w)[2]) const {
        const uint4 a = __ldg(p + T);
        w[0] = make_uint2(a.x, a.y); w[1] = make_uint2(a.z, a.w);
    }
    __device__ __forceinline__ void get4(uint2 (&w)[4]) const {
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const uint4 a = __ldg(p + (2 + i) * T);
            w[2 * i] = make_uint2(a.x, a.y); w[2 * i + 1] = make_uint2(a.z, a.w);
        }
    }
    __device__ __forceinline__ void get8(uint2 (&w)[8]) const {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint4 a = __ldg(p + (4 + i) * T);
            w[2 * i] = make_uint2(a.x, a.y); w[2 * i + 1] = make_uint2(a.z, a.w);
        }
    }
};
// the 15 twiddles of a pass held in registers: loaded once per pass and reused by every polynomial of the pass
struct TwRegs {
    uint2 w[16]; // node order: w[1], w[2..3], w[4..7], w[8..15]
    template <class TW> __device__ __forceinline__ void load(const TW &src, const int nlev) {
        if (nlev >= 4) w[1] = src.get1();
        if (nlev >= 3) { uint2 a[2]; src.get2(a); w[2] = a[0]; w[3] = a[1]; }
        if (nlev >= 2) { uint2 a[4]; src.get4(a);
#pragma unroll
            for (int i = 0; i < 4; i++) w[4 + i] = a[i]; }
        { uint2 a[8]; src.get8(a);
#pragma unroll
            for (int i = 0; i < 8; i++) w[8 + i] = a[i]; }
    }
    __device__ __forceinline__ uint2 get1() const { return w[1]; }
    __device__ __forceinline__ void get2(uint2 (&o)[2]) const { o[0] = w[2]; o[1] = w[3]; }
    __device__ __forceinline__ void get4(uint2 (&o)[4]) const {
#pragma unroll
        for (int i = 0; i < 4; i++) o[i] = w[4 + i];
    }
    __device__ __forceinline__ void get8(uint2 (&o)[8]) const {
#pragma unroll
        for (int i = 0; i < 8; i++) o[i] = w[8 + i];
    }
};

// L2 prefetch of a contiguous global range by one thread (TMA bulk prefetch; bytes must be a multiple of 16)
__device__ __forceinline__ void prefetch_l2_bulk(const void *gptr, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}
// Cluster barrier halves.  Only shared memory is exchanged between the CTAs of a cluster (ld.shared::cluster), never global data, so
// the default release/acquire forms -- which ptxas implements as MEMBAR.ALL.GPU + an L1 invalidation (CCTL.IVALL) that also throws
// away the cached twiddles -- are replaced by a CTA-scope fence (orders this thread's st.shared) followed by the relaxed arrive.
#ifdef PGB_CLUSTER_BARRIER_DEFAULT
__device__ __forceinline__ void cl_arrive() { asm volatile("barrier.cluster.arrive.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.aligned;" ::: "memory"); }
#else
__device__ __forceinline__ void cl_arrive() {
    asm volatile("fence.acq_rel.cta;" ::: "memory");
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
}
__device__ __forceinline__ void cl_wait() {
    asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}
#endif

__device__ __forceinline__ uint32_t ld_cluster(uint32_t cluster_smem_addr) {
    uint32_t v;
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(cluster_smem_addr));
    return v;
}

// one pass over a plane: element j of thread t sits at word (a << (SG + 4)) | b + (j << SG), a = t >> SG, b = t & (2^SG - 1)
template <int L, int SG> __device__ __forceinline__ int pass_base(int t) {
    const int a = t >> SG, b = t & ((1 << SG) - 1);
    return (a << (SG + 4)) | b;
}

// i64 -> residue in [0, 3q) for a run-time prime (same steps as n120::from_i64<K>)
struct PrimeCtx {
    uint32_t q, c32, c32s, neg64; // q, 2^32 mod q, its Shoup companion, q - (2^64 mod q)
};
__device__ __forceinline__ uint32_t from_i64_rt(long long v, const PrimeCtx &pc) {
    const uint32_t uh = (uint32_t)((unsigned long long)v >> 32), ul = (uint32_t)v;
    uint32_t r = csub(mul_shoup(uh, pc.c32, pc.c32s, pc.q) + (ul - (ul >> 30) * pc.q), 2 * pc.q);
    if (v < 0) r += pc.neg64;
    return r;
}

// compile-time part of the swizzle of a word offset whose bits do not overlap the thread's own index bits
template <int L> __host__ __device__ constexpr int fmask(int i) {
    return (((i >> 5) & 3) << 2) ^ (GGeo<L>::SIG3 ? (((i >> 7) & 3) << 3) : (((i >> 8) & 1) << 4));
}

// Low 32-bit words of v = sum_k t_k M_k - e Q (mod 2^128) for the coefficient at byte offset `off` of every CTA's planes, with
// e = round(sum t_k / Q[k]): |v| < Q (1/2 - 2^-27) here, so a 2^-28-accurate estimate of the fraction (60 fractional bits, truncated
// constants) always rounds to the right integer.  Column sums of 32-bit words, -e Q folded in as + e (2^128 - Q); `nwords` = how many
// words the digits need (ceil(S K / 32)), the others are returned as 0.
template <int NP> __device__ __forceinline__ void crt_low_words(const GadgetArgs &p, const uint32_t (&rb)[NP], const uint32_t off, const int nwords,
                                                               uint32_t &w0, uint32_t &w1, uint32_t &w2, uint32_t &w3) {
    uint32_t tk[NP];
#pragma unroll
    for (int k = 0; k < NP; k++) tk[k] = ld_cluster(rb[k] + off);
    unsigned long long fr = 1ull << 59;
#pragma unroll
    for (int k = 0; k < NP; k++) fr += (unsigned long long)tk[k] * p.inv60[k];
    const uint32_t e = (uint32_t)(fr >> 60);
    unsigned long long a0 = (unsigned long long)e * p.nq_w[0];
#pragma unroll
    for (int k = 0; k < NP; k++) a0 += (unsigned long long)tk[k] * p.m_w[k][0];
    unsigned long long a1 = (a0 >> 32) + (unsigned long long)e * p.nq_w[1];
#pragma unroll
    for (int k = 0; k < NP; k++) a1 += (unsigned long long)tk[k] * p.m_w[k][1];
    w0 = (uint32_t)a0; w1 = (uint32_t)a1; w2 = 0; w3 = 0;
    if (nwords > 2) {
        unsigned long long a2 = (a1 >> 32) + (unsigned long long)e * p.nq_w[2];
        if (NP == 4) { // three-prime M_k are below 2^60
#pragma unroll
            for (int k = 0; k < NP; k++) a2 += (unsigned long long)tk[k] * p.m_w[k][2];
        }
        w2 = (uint32_t)a2;
        if (nwords > 3) {
            unsigned long long a3 = (a2 >> 32) + (unsigned long long)e * p.nq_w[3];
            if (NP == 4) {
#pragma unroll
                for (int k = 0; k < NP; k++) a3 += (unsigned long long)tk[k] * p.m_w[k][3];
            }
            w3 = (uint32_t)a3;
        }
    }
}

// ---- CRT + balanced digits of one output column for base2k < 32, G coefficients of a thread at a time ------------------------------------
// CTA K owns the coefficient chunks K, K + NP, ... (16 chunks of T coefficients per polynomial); a thread processes G of its coefficients
// together so that the CTA-uniform work -- the CRT constants, the per-step decisions (does limb j take a body limb, is digit j stored),
// pointer updates -- is paid once per group and step instead of once per coefficient and step (round 1: ~250 thread instructions per
// coefficient, a quarter of them predicates, branches and constant loads).
//
// Per coefficient: v = sum_k t_k M_k - e Q (mod 2^(32 NW)) with e = round(sum t_k / Q[k]) as in crt_low_words, half = sum_j 2^(K-1) 2^(jK)
// folded into the column sums.  Balanced digits of W = v + sum_j body_j 2^((S-1-j)K): the unsigned K-bit fields of u = W + half, each minus
// 2^(K-1).  u is kept modulo 2^(32 NW) >= 2^(S K) and shifted right by K per digit; body limb j joins at bit 0 just before its own digit is
// read (the carry out of the top digit is dropped as in vec_znx_big_normalize).  Nothing above bit S K ever moves down into a digit that is
// still to be read, so NW = ceil(S K / 32) words are the whole state.
template <int L, int NP, int NW, bool SMALL>
__device__ __forceinline__ void gadget_tail_narrow(const GadgetArgs &p, const uint32_t (&rb)[NP], const int K, const int t, const int st,
                                                   const int o, const long long *__restrict__ in, long long *__restrict__ res) {
    typedef GGeo<L> Geo;
    constexpr int n = Geo::N, T = Geo::T;
    constexpr int G = NP == 4 ? 4 : 3;       // NP = 4: chunks K, K+4, K+8, K+12; NP = 3: two groups (K, K+3, K+6), (K+9, K+12, K+15 if < 16)
    const int Kb = p.K, S = p.S, cols_out = p.cols_out;
    const int a_start = p.res_size < S ? p.res_size : S; // digits j >= a_start are discarded (carry only)
    const size_t res_ls = (size_t)cols_out * n, in_ls = (size_t)p.in_cols * n;
    const uint32_t kmask32 = (1u << Kb) - 1u, khalf32 = 1u << (Kb - 1);
    const uint32_t hw[4] = {(uint32_t)p.half_lo, (uint32_t)(p.half_lo >> 32), (uint32_t)p.half_hi, (uint32_t)(p.half_hi >> 32)};
#pragma unroll 1
    for (int c0 = K; c0 < 16; c0 += NP * G) {
        bool valid[G];
        uint32_t tk[G][NP];
#pragma unroll
        for (int g = 0; g < G; g++) {
            const int chunk = c0 + g * NP;
            valid[g] = NP == 4 || chunk < 16;
            const uint32_t off = (uint32_t)(o * n + (swz<L>((valid[g] ? chunk : c0) * T) ^ st)) * 4u;
#pragma unroll
            for (int k = 0; k < NP; k++) tk[g][k] = ld_cluster(rb[k] + off);
        }
        // first body limbs in flight before the CRT arithmetic
        const long long *bp = in + (size_t)(S - 1) * in_ls + c0 * T + t;       // body limb S - 1 (column 0), coefficient of g = 0
        long long body[G];
        if (SMALL) {
#pragma unroll
            for (int g = 0; g < G; g++) body[g] = (S - 1 < p.small_size && valid[g]) ? __ldg(bp + g * NP * T) : 0;
        }
        uint32_t w[G][NW];
        {
            unsigned long long a[G];
            uint32_t e[G];
#pragma unroll
            for (int g = 0; g < G; g++) a[g] = 1ull << 59;
#pragma unroll
            for (int k = 0; k < NP; k++) {
                const uint32_t c = p.inv60[k];
#pragma unroll
                for (int g = 0; g < G; g++) a[g] += (unsigned long long)tk[g][k] * c;
            }
#pragma unroll
            for (int g = 0; g < G; g++) e[g] = (uint32_t)(a[g] >> 60);
#pragma unroll
            for (int wi = 0; wi < NW; wi++) {
                const uint32_t nq = p.nq_w[wi], h = hw[wi];
#pragma unroll
                for (int g = 0; g < G; g++) a[g] = (wi ? (a[g] >> 32) : 0ull) + (unsigned long long)e[g] * nq + h;
                if (NP == 4 || wi < 2) { // three-prime M_k are below 2^60
#pragma unroll
                    for (int k = 0; k < NP; k++) {
                        const uint32_t mw = p.m_w[k][wi];
#pragma unroll
                        for (int g = 0; g < G; g++) a[g] += (unsigned long long)tk[g][k] * mw;
                    }
                }
#pragma unroll
                for (int g = 0; g < G; g++) w[g][wi] = (uint32_t)a[g];
            }
        }
        long long *out_p = res + (size_t)(S - 1) * res_ls + (size_t)o * n + c0 * T + t;
        // one digit step; the body limbs of the NEXT step are requested before this step's arithmetic, the two buffers alternate
#define TAIL_STEP(J, CUR, NXT)                                                                                                     \
    {                                                                                                                              \
        const int j_ = (J);                                                                                                        \
        if (SMALL) {                                                                                                               \
            bp -= in_ls;                                                                                                           \
            _Pragma("unroll") for (int g = 0; g < G; g++)                                                                          \
                NXT[g] = (j_ >= 1 && j_ - 1 < p.small_size && valid[g]) ? __ldg(bp + g * NP * T) : 0;                              \
            if (j_ < p.small_size) {                                                                                               \
                _Pragma("unroll") for (int g = 0; g < G; g++) {                                                                    \
                    const uint32_t lo = (uint32_t)CUR[g], hi = (uint32_t)((unsigned long long)CUR[g] >> 32), sx = (uint32_t)(CUR[g] >> 63); \
                    if (NW == 1) w[g][0] += lo;                                                                                    \
                    else if (NW == 2) asm("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(w[g][0]), "+r"(w[g][1]) : "r"(lo), "r"(hi)); \
                    else if (NW == 3)                                                                                              \
                        asm("add.cc.u32 %0, %0, %3; addc.cc.u32 %1, %1, %4; addc.u32 %2, %2, %5;"                                  \
                            : "+r"(w[g][0]), "+r"(w[g][1]), "+r"(w[g][NW > 2 ? 2 : 0]) : "r"(lo), "r"(hi), "r"(sx));               \
                    else                                                                                                           \
                        asm("add.cc.u32 %0, %0, %4; addc.cc.u32 %1, %1, %5; addc.cc.u32 %2, %2, %6; addc.u32 %3, %3, %6;"          \
                            : "+r"(w[g][0]), "+r"(w[g][1]), "+r"(w[g][NW > 2 ? 2 : 0]), "+r"(w[g][NW > 3 ? 3 : 0])                 \
                            : "r"(lo), "r"(hi), "r"(sx));                                                                          \
                }                                                                                                                  \
            }                                                                                                                      \
        }                                                                                                                          \
        if (j_ < a_start) {                                                                                                        \
            _Pragma("unroll") for (int g = 0; g < G; g++)                                                                          \
                if (valid[g]) out_p[g * NP * T] = (long long)((int)(w[g][0] & kmask32) - (int)khalf32);                            \
        }                                                                                                                          \
        out_p -= res_ls;                                                                                                           \
        _Pragma("unroll") for (int g = 0; g < G; g++) {                                                                            \
            _Pragma("unroll") for (int wi = 0; wi + 1 < NW; wi++) w[g][wi] = __funnelshift_r(w[g][wi], w[g][wi + 1], Kb);          \
            w[g][NW - 1] >>= Kb;                                                                                                   \
        }                                                                                                                          \
    }
        long long body2[G];
        int j = S - 1;
#pragma unroll 1
        for (; j >= 1; j -= 2) {
            TAIL_STEP(j, body, body2)
            TAIL_STEP(j - 1, body2, body)
        }
        if (j == 0) TAIL_STEP(0, body, body2)
#undef TAIL_STEP
        long long *zp = res + (size_t)o * n + c0 * T + t;
        for (int j = a_start; j < p.res_size; j++) {
#pragma unroll
            for (int g = 0; g < G; g++)
                if (valid[g]) zp[(size_t)j * res_ls + g * NP * T] = 0;
        }
    }
}

// Products with the collapsed key for the 16 coefficients of a thread (four chunks of four): out[o] = sum_r a[r] * key[r][o] mod q for CO
// output columns.  The key is stored times 2^32 (gadget_collapse_key_kernel), the u64 row sums of up to four rows (< 8 q^2 < 2^63) take
// one Montgomery reduction (two instructions) and one conditional subtraction, results in [0, 2q).  Every key address is the thread's
// base plus a compile-time offset.
template <int L, int CO>
__device__ __forceinline__ void gadget_mac(uint32_t *__restrict__ sm, const uint4 *__restrict__ ck, const int R, const int (&qa)[4], const uint32_t q,
                                           const uint32_t qni) {
    constexpr int n = GGeo<L>::N, T = GGeo<L>::T;
    constexpr int col_stride = n, row_stride = n * CO; // in uint4: [r][col][k 0..3][chunk][t]
#pragma unroll 1
    for (int c = 0; c < 4; c++) {
        // rows in groups of four; the partial result of a group goes straight to the output planes: rows 0..3 have been consumed by then and
        // CO <= 4, and the words are this thread's own
#pragma unroll 1
        for (int r0 = 0; r0 < R; r0 += 4) {
            unsigned long long acc[CO][4];
            const uint4 *ckr = ck + (size_t)r0 * row_stride + c * T;
            const uint32_t *ar = sm + r0 * n + qa[c];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                if (u == 0 || r0 + u < R) {
                    const uint4 a = *reinterpret_cast<const uint4 *>(ar + u * n);
#pragma unroll
                    for (int o = 0; o < CO; o++) {
                        const uint4 m = __ldg(ckr + u * row_stride + o * col_stride);
                        if (u == 0) {
                            acc[o][0] = (unsigned long long)a.x * m.x; acc[o][1] = (unsigned long long)a.y * m.y;
                            acc[o][2] = (unsigned long long)a.z * m.z; acc[o][3] = (unsigned long long)a.w * m.w;
                        } else {
                            acc[o][0] += (unsigned long long)a.x * m.x; acc[o][1] += (unsigned long long)a.y * m.y;
                            acc[o][2] += (unsigned long long)a.z * m.z; acc[o][3] += (unsigned long long)a.w * m.w;
                        }
                    }
                }
            }
#pragma unroll
            for (int o = 0; o < CO; o++) {
                uint4 *dst = reinterpret_cast<uint4 *>(sm + o * n + qa[c]);
                uint32_t y[4];
#pragma unroll
                for (int i = 0; i < 4; i++) y[i] = csub(redc64(acc[o][i], q, qni), 2 * q); // redc < 2^31 + q < 4q
                if (r0 != 0) {
                    const uint4 prev = *dst;
                    y[0] = csub(y[0] + prev.x, 2 * q); y[1] = csub(y[1] + prev.y, 2 * q);
                    y[2] = csub(y[2] + prev.z, 2 * q); y[3] = csub(y[3] + prev.w, 2 * q);
                }
                *dst = make_uint4(y[0], y[1], y[2], y[3]);
            }
        }
    }
}

// The same for the automorphism epilogue (see GadgetArgs): a thread still owns OUTPUT coefficients (coalesced stores and reads of `a`); the
// residues and the body limbs are gathered at the source coefficient js = j' * p^-1 mod 2n, whose sign (flip) is applied to x before the
// digits (modes 1-3) or to the digits (mode 4).  Per group: the CRT constants, the step decisions and the pointers are shared as in
// gadget_tail_narrow; per coefficient: source index, sign, gathered loads.
template <int L, int NP, int NW>
__device__ __forceinline__ void gadget_tail_narrow_aut(const GadgetArgs &p, const uint32_t (&rb)[NP], const int K, const int t, const int o,
                                                       const long long *__restrict__ in, long long *__restrict__ res) {
    typedef GGeo<L> Geo;
    constexpr int n = Geo::N, T = Geo::T;
    constexpr int G = NP == 4 ? 4 : 3;
    const int Kb = p.K, S = p.S, cols_out = p.cols_out, mode = p.aut_mode;
    const int a_start = p.res_size < S ? p.res_size : S;
    const size_t res_ls = (size_t)cols_out * n, in_ls = (size_t)p.in_cols * n;
    const uint32_t kmask32 = (1u << Kb) - 1u, khalf32 = 1u << (Kb - 1);
    const uint32_t hw[4] = {(uint32_t)p.half_lo, (uint32_t)(p.half_lo >> 32), (uint32_t)p.half_hi, (uint32_t)(p.half_hi >> 32)};
    const bool with_small = o == 0 && p.small_size > 0;
#pragma unroll 1
    for (int c0 = K; c0 < 16; c0 += NP * G) {
        bool valid[G], flip[G], negx[G];
        int js[G];
        uint32_t tk[G][NP];
#pragma unroll
        for (int g = 0; g < G; g++) {
            const int chunk = c0 + g * NP;
            valid[g] = NP == 4 || chunk < 16;
            const int jo = (valid[g] ? chunk : c0) * T + t;
            const uint32_t j0 = ((uint32_t)jo * p.aut_pinv) & (uint32_t)(2 * n - 1);
            js[g] = (int)(j0 & (uint32_t)(n - 1));
            flip[g] = j0 >= (uint32_t)n;                                          // sign of the permuted coefficient
            negx[g] = mode == 4 ? false : (mode == 3 ? !flip[g] : flip[g]);       // sign applied to x before the digits
            const uint32_t off = (uint32_t)(o * n + swz<L>(js[g])) * 4u;
#pragma unroll
            for (int k = 0; k < NP; k++) tk[g][k] = ld_cluster(rb[k] + off);
        }
        uint32_t w[G][NW];
        {
            unsigned long long a[G];
            uint32_t e[G];
#pragma unroll
            for (int g = 0; g < G; g++) a[g] = 1ull << 59;
#pragma unroll
            for (int k = 0; k < NP; k++) {
                const uint32_t c = p.inv60[k];
#pragma unroll
                for (int g = 0; g < G; g++) a[g] += (unsigned long long)tk[g][k] * c;
            }
#pragma unroll
            for (int g = 0; g < G; g++) e[g] = (uint32_t)(a[g] >> 60);
#pragma unroll
            for (int wi = 0; wi < NW; wi++) {
                const uint32_t nq = p.nq_w[wi];
#pragma unroll
                for (int g = 0; g < G; g++) a[g] = (wi ? (a[g] >> 32) : 0ull) + (unsigned long long)e[g] * nq;
                if (NP == 4 || wi < 2) {
#pragma unroll
                    for (int k = 0; k < NP; k++) {
                        const uint32_t mw = p.m_w[k][wi];
#pragma unroll
                        for (int g = 0; g < G; g++) a[g] += (unsigned long long)tk[g][k] * mw;
                    }
                }
#pragma unroll
                for (int g = 0; g < G; g++) w[g][wi] = (uint32_t)a[g];
            }
        }
        // u = +-v + half (mod 2^(32 NW))
#pragma unroll
        for (int g = 0; g < G; g++) {
            uint32_t cy = negx[g] ? 1u : 0u;
            const uint32_t xm = negx[g] ? 0xffffffffu : 0u;
#pragma unroll
            for (int wi = 0; wi < NW; wi++) { // (w ^ xm) + cy + half, word by word with explicit carries
                const unsigned long long s2 = (unsigned long long)(w[g][wi] ^ xm) + cy + hw[wi];
                w[g][wi] = (uint32_t)s2;
                cy = (uint32_t)(s2 >> 32);
            }
        }
        const long long *bp = in + (size_t)(S - 1) * in_ls;                                   // body (column 0), gathered at js
        const long long *pp = in + (size_t)(S - 1) * in_ls + (size_t)o * n + c0 * T + t;      // a, column o, at the output coefficient
        long long *out_p = res + (size_t)(S - 1) * res_ls + (size_t)o * n + c0 * T + t;
        // signed 64-bit value joined at bit 0 of the NW-word state, added (sub = false) or subtracted
        auto join = [&](uint32_t (&ww)[NW], const long long sv, const bool sub) {
            const uint32_t xm = sub ? 0xffffffffu : 0u;
            const uint32_t lo = (uint32_t)sv ^ xm, hi = (uint32_t)((unsigned long long)sv >> 32) ^ xm, sx = (uint32_t)(sv >> 63) ^ xm;
            unsigned long long s2 = (unsigned long long)ww[0] + lo + (sub ? 1u : 0u);
            ww[0] = (uint32_t)s2;
            if (NW > 1) { s2 = (s2 >> 32) + ww[NW > 1 ? 1 : 0] + hi; ww[NW > 1 ? 1 : 0] = (uint32_t)s2; }
            if (NW > 2) { s2 = (s2 >> 32) + ww[NW > 2 ? 2 : 0] + sx; ww[NW > 2 ? 2 : 0] = (uint32_t)s2; }
            if (NW > 3) { s2 = (s2 >> 32) + ww[NW > 3 ? 3 : 0] + sx; ww[NW > 3 ? 3 : 0] = (uint32_t)s2; }
        };
#pragma unroll 1
        for (int j = S - 1; j >= 0; j--) {
            if (with_small && j < p.small_size) {
                long long b[G];
#pragma unroll
                for (int g = 0; g < G; g++) b[g] = valid[g] ? __ldg(bp + js[g]) : 0;
#pragma unroll
                for (int g = 0; g < G; g++) join(w[g], b[g], negx[g]);
            }
            if (mode != 4 && j < p.post_size) {
                long long b[G];
#pragma unroll
                for (int g = 0; g < G; g++) b[g] = valid[g] ? __ldg(pp + g * NP * T) : 0;
#pragma unroll
                for (int g = 0; g < G; g++) join(w[g], b[g], mode == 2);
            }
            bp -= in_ls;
            pp -= in_ls;
            if (j < a_start) {
#pragma unroll
                for (int g = 0; g < G; g++) {
                    int d = (int)(w[g][0] & kmask32) - (int)khalf32;
                    if (mode == 4 && flip[g]) d = -d;
                    if (valid[g]) out_p[g * NP * T] = (long long)d;
                }
            }
            out_p -= res_ls;
#pragma unroll
            for (int g = 0; g < G; g++) {
#pragma unroll
                for (int wi = 0; wi + 1 < NW; wi++) w[g][wi] = __funnelshift_r(w[g][wi], w[g][wi + 1], Kb);
                w[g][NW - 1] >>= Kb;
            }
        }
        long long *zp = res + (size_t)o * n + c0 * T + t;
        for (int j = a_start; j < p.res_size; j++) {
#pragma unroll
            for (int g = 0; g < G; g++)
                if (valid[g]) zp[(size_t)j * res_ls + g * NP * T] = 0;
        }
    }
}

template <int L, bool AUT, int NP> __device__ __forceinline__ void gadget_body(const GadgetArgs &p, uint32_t *__restrict__ sm, const uint2 *__restrict__ twf,
                                                             const uint2 *__restrict__ twi, const uint4 *__restrict__ lastf,
                                                             const uint4 *__restrict__ lasti, const int K) {
    typedef GGeo<L> G;
    constexpr int n = G::N, T = G::T;
    const PrimeCtx pc = {p.prime[K].q, p.prime[K].c32, p.prime[K].c32s, p.prime[K].neg64};
    const uint32_t q = pc.q, z = p.zero;
    const int t = threadIdx.x;
    cg::cluster_group cluster = cg::this_cluster();
    const int nclusters = gridDim.x / NP, cid = blockIdx.x / NP;
    const TwGlobal top_f = {twf, 1u}, top_i = {twi, 1u}; // CTA-uniform addresses: one L1 broadcast per load
    const int R = p.R, cols_out = p.cols_out;
    const int amax_allowed = p.bound_bits - __ldg(p.key_bits); // see ntt120_gadget_fused (collapsed-key bound)
    const uint32_t sm_base = (uint32_t)__cvta_generic_to_shared(sm);
    uint32_t rb[NP]; // shared::cluster addresses of the NP CTAs' plane 0
#pragma unroll
    for (int k = 0; k < NP; k++) asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb[k]) : "r"(sm_base), "r"(k));
    // swizzled word addresses: stride-T pattern = (st ^ const) + j*T, 16-consecutive pattern = qa[c], stride-2^SG2 pattern = sb ^ const
    const int st = swz<L>(t);
    const int sb = swz<L>(pass_base<L, G::SG2>(t));
    int qa[4];
#pragma unroll
    for (int c = 0; c < 4; c++) qa[c] = swz<L>(16 * t + 4 * c);
#define P1_ADDR(j) ((st ^ fmask<L>((j) * T)) + (j) * T)
#define P2_ADDR(j) (sb ^ swz<L>((j) << G::SG2))

    for (int ct = cid; ct < p.batch; ct += nclusters) {
        const long long *in = reinterpret_cast<const long long *>(p.in + (size_t)ct * p.in_bs);
        // L2 prefetch (one thread per 8n-byte limb, spread over the four CTAs): the body limbs this ciphertext needs in its CRT phase and
        // the mask limbs of the cluster's next ciphertext
        if ((t % NP) == K && (t / NP) < R + p.small_size) {
            const int u = t / NP;
            if (u < p.small_size) {
                prefetch_l2_bulk(in + (size_t)u * p.in_cols * n, n * 8);
            } else if (ct + nclusters < p.batch) {
                const int r = u - p.small_size, limb = r / p.row_cols, col = r % p.row_cols + p.row_col0;
                prefetch_l2_bulk(in + (size_t)nclusters * (p.in_bs / 8) + ((size_t)limb * p.in_cols + col) * n, n * 8);
            }
        }
        // ---- forward pass 1 (levels 0..3, stride T) fused with the i64 load and the magnitude scan -------------------------------
        uint32_t mag = 0;       // OR of |v|-like patterns of the values that fit 32 bits
        bool too_big = false;
        long long v[16]; // the loads of polynomial r + 1 are issued before the butterflies of polynomial r
        {
            const long long *src = in + (size_t)p.row_col0 * n + t;
#pragma unroll
            for (int j = 0; j < 16; j++) v[j] = __ldg(src + j * T);
        }
        for (int r = 0; r < R; r++) {
            uint32_t x[16];
            uint32_t wide = 0;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                const int lo = (int)v[j], hi = (int)(v[j] >> 32), sg = lo >> 31;
                wide |= (uint32_t)(hi ^ sg);
                mag |= (uint32_t)(lo ^ sg);
                x[j] = (uint32_t)lo + ((uint32_t)sg & (3u * q)); // [0, 3q) when v fits 32 bits (3q > 2^31)
            }
            if (__any_sync(0xffffffffu, wide != 0)) { // rare: digits beyond 32 bits, full-range conversion and magnitude
                unsigned long long m64 = 0;
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    m64 |= (unsigned long long)(v[j] ^ (v[j] >> 63));
                    x[j] = from_i64_rt(v[j], pc);
                }
                too_big |= amax_allowed < 0 || (m64 >> (amax_allowed > 63 ? 63 : amax_allowed)) != 0;
            }
            if (r + 1 < R) {
                const int limb = (r + 1) / p.row_cols, col = (r + 1) % p.row_cols + p.row_col0;
                const long long *src = in + ((size_t)limb * p.in_cols + col) * n + t;
#pragma unroll
                for (int j = 0; j < 16; j++) v[j] = __ldg(src + j * T);
            }
            ct16<4>(x, top_f, q, z);
            if (r == 0) cl_wait(); // peers have finished reading my planes (CRT of the previous ciphertext)
            uint32_t *pl = sm + r * n;
#pragma unroll
            for (int j = 0; j < 16; j++) pl[P1_ADDR(j)] = x[j];
        }
        too_big |= amax_allowed < 0 || (amax_allowed < 32 && (mag >> amax_allowed) != 0);
        const int bad = __syncthreads_or(too_big);
        if (bad) { // identical decision in all four CTAs (same inputs): this ciphertext takes the per-limb kernels
            if (t == 0 && K == 0) { // ok[batch] counts the flagged ciphertexts, ok[batch + 1 ..] lists them for the per-limb kernels
                p.ok[ct] = 0;
                p.ok[p.batch + 1 + atomicAdd(p.ok + p.batch, 1)] = ct;
            }
            cl_arrive();
            continue;
        }
        // ---- forward pass 2 ---------------------------------------------------------------------------------------------------
        {
            const TwGlobal tw = {twf, (1u << G::LV2) | (uint32_t)(t >> G::SG2)};
            for (int r = 0; r < R; r++) {
                uint32_t *pl = sm + r * n;
                uint32_t x[16];
#pragma unroll
                for (int j = 0; j < 16; j++) x[j] = pl[P2_ADDR(j)];
                ct16<G::NL2>(x, tw, q, z);
#pragma unroll
                for (int j = 0; j < 16; j++) pl[P2_ADDR(j)] = x[j];
            }
        }
        __syncthreads();
        // ---- forward pass 3 (16 consecutive words per thread), results reduced to [0, 2q) ------------------------------------------
        {
            const TwLast tws = {lastf + t, T};
            TwRegs tw;
            tw.load(tws, G::NL3);
            for (int r = 0; r < R; r++) {
                uint32_t *pl = sm + r * n;
                uint32_t x[16];
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(pl + qa[c]);
                    x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
                }
                ct16<G::NL3>(x, tw, q, z);
#pragma unroll
                for (int c = 0; c < 4; c++)
                    *reinterpret_cast<uint4 *>(pl + qa[c]) = make_uint4(csub(x[4 * c], 2 * q), csub(x[4 * c + 1], 2 * q),
                                                                        csub(x[4 * c + 2], 2 * q), csub(x[4 * c + 3], 2 * q));
            }
        }
        // ---- products with the collapsed key (thread-private words: no barrier) ---------------------------------------------------
        {
            const uint4 *ck = reinterpret_cast<const uint4 *>(p.ckey) + (size_t)K * (n / 4) + t;
            const uint32_t qni = p.prime[K].qni;
            if (cols_out == 2) gadget_mac<L, 2>(sm, ck, R, qa, q, qni);
            else if (cols_out == 3) gadget_mac<L, 3>(sm, ck, R, qa, q, qni);
            else if (cols_out == 4) gadget_mac<L, 4>(sm, ck, R, qa, q, qni);
            else gadget_mac<L, 1>(sm, ck, R, qa, q, qni);
        }
        // ---- inverse pass 1 (levels L-1 .. L-NL3, 16 consecutive words) --------------------------------------------------------------
        {
            const TwLast tws = {lasti + t, T};
            TwRegs tw;
            tw.load(tws, G::NL3);
            for (int o = 0; o < cols_out; o++) {
                uint32_t *pl = sm + o * n;
                uint32_t x[16];
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(pl + qa[c]);
                    x[4 * c] = v.x; x[4 * c + 1] = v.y; x[4 * c + 2] = v.z; x[4 * c + 3] = v.w;
                }
                gs16<G::NL3>(x, tw, q, z);
#pragma unroll
                for (int c = 0; c < 4; c++) *reinterpret_cast<uint4 *>(pl + qa[c]) = make_uint4(x[4 * c], x[4 * c + 1], x[4 * c + 2], x[4 * c + 3]);
            }
        }
        __syncthreads();
        // ---- inverse pass 2 ---------------------------------------------------------------------------------------------------
        {
            const TwGlobal tw = {twi, (1u << G::LV2) | (uint32_t)(t >> G::SG2)};
            for (int o = 0; o < cols_out; o++) {
                uint32_t *pl = sm + o * n;
                uint32_t x[16];
#pragma unroll
                for (int j = 0; j < 16; j++) x[j] = pl[P2_ADDR(j)];
                gs16<G::NL2>(x, tw, q, z);
#pragma unroll
                for (int j = 0; j < 16; j++) pl[P2_ADDR(j)] = x[j];
            }
        }
        __syncthreads();
        // ---- inverse pass 3 (levels 3..0, stride T), scale by CRT_k / n, canonical residues back to the plane ----------------------
        {
            const uint32_t cn = p.crt_ninv[K], cns = p.crt_ninv_sh[K];
            TwRegs top_r;
            top_r.load(top_i, 4);
            for (int o = 0; o < cols_out; o++) {
                uint32_t *pl = sm + o * n;
                uint32_t x[16];
#pragma unroll
                for (int j = 0; j < 16; j++) x[j] = pl[P1_ADDR(j)];
                gs16<4>(x, top_r, q, z);
#pragma unroll
                for (int j = 0; j < 16; j++) pl[P1_ADDR(j)] = csub(mul_shoup(x[j], cn, cns, q), q);
            }
        }
        cl_arrive();
        cl_wait(); // all four residues of every coefficient are in place
        // ---- CRT + digits: CTA K takes the coefficient chunks K, K + NP, ... of T consecutive coefficients (16 chunks per polynomial) ------
        {
            const int Kb = p.K, S = p.S;
            const int a_start = p.res_size < S ? p.res_size : S; // digits j >= a_start are discarded (carry only)
            const size_t res_ls = (size_t)cols_out * n, in_ls = (size_t)p.in_cols * n;
            const int nwords = (S * Kb + 31) >> 5;
            const bool narrow = Kb < 32; // digit fields inside one 32-bit word: funnel shifts over four words
            const unsigned long long kmask = (1ull << Kb) - 1, khalf = 1ull << (Kb - 1);
            long long *res_ct = reinterpret_cast<long long *>(p.res + (size_t)ct * p.res_bs) + t;
            const long long *in_kt = in + t + (size_t)(S - 1) * in_ls; // body limb S - 1 (column 0)
            if constexpr (AUT) {
                // Automorphism epilogue (see GadgetArgs): this thread still owns OUTPUT coefficients chunk T + t (coalesced stores and
                // reads of `a`); residues and body limbs are gathered at the source coefficient.
                const int mode = p.aut_mode;
                if (narrow) { // base2k < 32: groups of coefficients, as many 32-bit words as the digits need
                    long long *res_base = reinterpret_cast<long long *>(p.res + (size_t)ct * p.res_bs);
                    for (int o = 0; o < cols_out; o++) {
                        if (nwords <= 2) gadget_tail_narrow_aut<L, NP, 2>(p, rb, K, t, o, in, res_base);
                        else if (nwords == 3) gadget_tail_narrow_aut<L, NP, 3>(p, rb, K, t, o, in, res_base);
                        else gadget_tail_narrow_aut<L, NP, 4>(p, rb, K, t, o, in, res_base);
                    }
                } else
                for (int o = 0; o < cols_out; o++) {
                    const bool with_small = o == 0 && p.small_size > 0;
#pragma unroll 1
                    for (int chunk = K; chunk < 16; chunk += NP) {
                        const int jo = chunk * T + t;
                        const uint32_t j0 = ((uint32_t)jo * p.aut_pinv) & (uint32_t)(2 * n - 1);
                        const int js = (int)(j0 & (uint32_t)(n - 1));
                        const bool flip = j0 >= (uint32_t)n;            // sign of the permuted coefficient
                        const bool negx = mode == 4 ? false : (mode == 3 ? !flip : flip); // sign applied to x before the digits
                        const uint32_t off = (uint32_t)(o * n + swz<L>(js)) * 4u;
                        uint32_t w0, w1, w2, w3;
                        crt_low_words<NP>(p, rb, off, 4, w0, w1, w2, w3);
                        // u = +-v + half as a 128-bit integer in two 64-bit words (mod 2^128; only the low S K bits are read)
                        unsigned long long lo = (unsigned long long)w0 | ((unsigned long long)w1 << 32), hi = (unsigned long long)w2 | ((unsigned long long)w3 << 32);
                        if (negx) {
                            lo = ~lo + 1;
                            hi = ~hi + (lo == 0);
                        }
                        lo += p.half_lo;
                        hi += p.half_hi + (lo < p.half_lo);
                        const long long *bp = in + (size_t)(S - 1) * in_ls + js;                     // body (column 0) at the source coefficient
                        const long long *pp = in + (size_t)(S - 1) * in_ls + (size_t)o * n + jo;     // a, column o, at the output coefficient
                        long long *out_p = reinterpret_cast<long long *>(p.res + (size_t)ct * p.res_bs) + (size_t)(S - 1) * res_ls + (size_t)o * n + jo;
                        for (int j = S - 1; j >= 0; j--) {
                            if (with_small && j < p.small_size) {
                                const long long sv = __ldg(bp);
                                const unsigned long long sl = (unsigned long long)sv, sh = (unsigned long long)(sv >> 63);
                                if (negx) {
                                    const unsigned long long nl = lo - sl;
                                    hi = hi - sh - (lo < sl);
                                    lo = nl;
                                } else {
                                    lo += sl;
                                    hi += sh + (lo < sl);
                                }
                            }
                            if (mode != 4 && j < p.post_size) {
                                const long long sv = __ldg(pp);
                                const unsigned long long sl = (unsigned long long)sv, sh = (unsigned long long)(sv >> 63);
                                if (mode == 2) {
                                    const unsigned long long nl = lo - sl;
                                    hi = hi - sh - (lo < sl);
                                    lo = nl;
                                } else {
                                    lo += sl;
                                    hi += sh + (lo < sl);
                                }
                            }
                            bp -= in_ls;
                            pp -= in_ls;
                            long long d = (long long)(lo & kmask) - (long long)khalf;
                            if (mode == 4 && flip) d = -d;
                            if (j < a_start) *out_p = d;
                            out_p -= res_ls;
                            lo = (lo >> Kb) | (hi << (64 - Kb));
                            hi >>= Kb;
                        }
                        long long *zp = reinterpret_cast<long long *>(p.res + (size_t)ct * p.res_bs) + (size_t)o * n + jo;
                        for (int j = a_start; j < p.res_size; j++) zp[(size_t)j * res_ls] = 0;
                    }
                }
            } else if (narrow) {
                long long *res_base = reinterpret_cast<long long *>(p.res + (size_t)ct * p.res_bs);
                for (int o = 0; o < cols_out; o++) {
                    if (o == 0 && p.small_size > 0) {
                        if (nwords <= 2) gadget_tail_narrow<L, NP, 2, true>(p, rb, K, t, st, o, in, res_base);
                        else if (nwords == 3) gadget_tail_narrow<L, NP, 3, true>(p, rb, K, t, st, o, in, res_base);
                        else gadget_tail_narrow<L, NP, 4, true>(p, rb, K, t, st, o, in, res_base);
                    } else {
                        if (nwords <= 2) gadget_tail_narrow<L, NP, 2, false>(p, rb, K, t, st, o, in, res_base);
                        else if (nwords == 3) gadget_tail_narrow<L, NP, 3, false>(p, rb, K, t, st, o, in, res_base);
                        else gadget_tail_narrow<L, NP, 4, false>(p, rb, K, t, st, o, in, res_base);
                    }
                }
            } else
            for (int o = 0; o < cols_out; o++) { // base2k >= 32: 64-bit digit fields, one coefficient at a time
                const bool with_small = o == 0 && p.small_size > 0;
#pragma unroll 1
                for (int chunk = K; chunk < 16; chunk += NP) {
                    const long long *sp = in_kt + chunk * T;
                    const uint32_t off = (uint32_t)(o * n + (swz<L>(chunk * T) ^ st)) * 4u;
                    uint32_t w0, w1, w2, w3;
                    crt_low_words<NP>(p, rb, off, nwords, w0, w1, w2, w3);
                    // digits as in gadget_tail_narrow, on two 64-bit halves (words the digits do not need are 0 and stay unread)
                    unsigned long long lo = ((unsigned long long)w1 << 32) | w0, hi = ((unsigned long long)w3 << 32) | w2;
                    lo += p.half_lo;
                    hi += p.half_hi + (lo < p.half_lo);
                    long long *out_p = res_ct + (size_t)(S - 1) * res_ls + (size_t)o * n + chunk * T;
                    for (int j = S - 1; j >= 0; j--) {
                        if (with_small && j < p.small_size) {
                            const long long sv = __ldg(sp);
                            const unsigned long long sl = (unsigned long long)sv;
                            lo += sl;
                            hi += (unsigned long long)(sv >> 63) + (lo < sl);
                        }
                        sp -= in_ls;
                        if (j < a_start) *out_p = (long long)(lo & kmask) - (long long)khalf;
                        out_p -= res_ls;
                        lo = (lo >> Kb) | (hi << (64 - Kb));
                        hi >>= Kb;
                    }
                    long long *zp = res_ct + (size_t)o * n + chunk * T;
                    for (int j = a_start; j < p.res_size; j++) zp[(size_t)j * res_ls] = 0;
                }
            }
        }
        if (t == 0 && K == 0) p.ok[ct] = 1;
        cl_arrive();
    }
    cl_wait();
#undef P1_ADDR
#undef P2_ADDR
}

template <int L, int MB, bool AUT, int NP>
__global__ void __cluster_dims__(NP, 1, 1) __launch_bounds__(GGeo<L>::T, MB * 256 / GGeo<L>::T)
    ntt120_gadget_kernel(const __grid_constant__ GadgetArgs p, const uint2 *__restrict__ twf, const uint2 *__restrict__ twi,
                         const uint4 *__restrict__ lastf, const uint4 *__restrict__ lasti) {
    extern __shared__ __align__(16) uint32_t smem[];
    constexpr int n = GGeo<L>::N;
    cl_arrive(); // primes the arrive/wait pairing used by the per-ciphertext loop
    const int k = blockIdx.x % NP; // = rank in the cluster; one code body for all primes (instruction-cache footprint)
    gadget_body<L, AUT, NP>(p, smem, twf + (size_t)k * n, twi + (size_t)k * n, lastf + (size_t)k * (n / 2), lasti + (size_t)k * (n / 2), k);
}

// collapsed key in the gadget kernel's layout: out[r][col][k][chunk][t][4] = sum_j 2^((S-1-j)K) * pmat[r][j * cols_out + col][k][16t + 4 chunk + w]
struct CollapseArgs2 {
    const char *pmat;
    uint32_t *out;
    int n, R, C, cols_out, S;
    // output row r (the r-th input poly of the gadget kernel) <- key row src_row[r], key limbs shifted by di[r], limbs j < jmax[r] only
    // (dsize == 1: src_row = r, di = 0, jmax = S; dsize > 1: the digit groups of keyswitching/glwe.rs:332-379, see ntt120_gadget_fused)
    signed char src_row[8], di[8], jmax[8];
    uint32_t c[32][4];
};
__global__ void __launch_bounds__(256) gadget_collapse_key_kernel(CollapseArgs2 p) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x; // quad index inside one prime plane
    const int n4 = p.n / 4;
    if (u >= n4) return;
    const int k = blockIdx.z;
    const PrimeRt pr(k);
    const int r = blockIdx.y / p.cols_out, col = blockIdx.y % p.cols_out;
    const size_t poly4 = (size_t)4 * n4; // uint4 per poly
    const int jm = p.jmax[r], di = p.di[r];
    const uint4 *src = reinterpret_cast<const uint4 *>(p.pmat) + ((size_t)p.src_row[r] * p.C + (size_t)di * p.cols_out + col) * poly4 + (size_t)k * n4 + u;
    unsigned long long acc[4] = {0, 0, 0, 0};
    for (int j0 = 0; j0 < jm; j0 += 16) {
        const int j1 = min(j0 + 16, jm);
        for (int j = j0; j < j1; j++) {
            const uint4 v = __ldg(src + (size_t)j * p.cols_out * poly4);
            const unsigned long long cj = p.c[j][k];
            acc[0] += v.x * cj; acc[1] += v.y * cj; acc[2] += v.z * cj; acc[3] += v.w * cj;
        }
#pragma unroll
        for (int i = 0; i < 4; i++) acc[i] = pr.reduce(acc[i]);
    }
    const int T = p.n / 16, t = u >> 2, c = u & 3;
    uint4 *dst = reinterpret_cast<uint4 *>(p.out) + ((size_t)r * p.cols_out + col) * poly4 + (size_t)k * n4 + (size_t)c * T + t;
    *dst = make_uint4((uint32_t)acc[0], (uint32_t)acc[1], (uint32_t)acc[2], (uint32_t)acc[3]);
}

static uint32_t pow2_mod(uint64_t e, uint32_t q) {
    uint64_t r = 1, b = 2;
    while (e) {
        if (e & 1) r = r * b % q;
        b = b * b % q;
        e >>= 1;
    }
    return (uint32_t)r;
}

template <int L, int MB, bool AUT, int NP> int launch_gadget_mb(pgb_module *m, const GadgetArgs &p, size_t smem) {
    typedef GGeo<L> G;
    static bool configured[32] = {};
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3(G::T);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = m->stream;
    if (!configured[m->device & 31]) {
        PGB_CHECK_CUDA(cudaFuncSetAttribute(ntt120_gadget_kernel<L, MB, AUT, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(96 << 10)));
        PGB_CHECK_CUDA(cudaFuncSetAttribute(ntt120_gadget_kernel<L, MB, AUT, NP>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
        configured[m->device & 31] = true;
    }
    // resident clusters for this shared-memory footprint (depends on R through smem)
    cfg.gridDim = dim3(NP * 148);
    int nc = 0;
    PGB_CHECK_CUDA(cudaOccupancyMaxActiveClusters(&nc, ntt120_gadget_kernel<L, MB, AUT, NP>, &cfg));
    if (nc < 1) {
        pgb_set_error("gadget kernel: no resident cluster fits (smem %zu)", smem);
        return PGB_ERR_UNSUPPORTED;
    }
    const int clusters = p.batch < nc ? p.batch : nc;
    cfg.gridDim = dim3(NP * clusters);
    { ProfScope _ps(m, PROF_GADGET);
    PGB_CHECK_CUDA(cudaLaunchKernelEx(&cfg, ntt120_gadget_kernel<L, MB, AUT, NP>, p, (const uint2 *)m->ntt_fwd, (const uint2 *)m->ntt_inv,
                                      (const uint4 *)m->ntt_last16_f, (const uint4 *)m->ntt_last16_i));
    }
    return PGB_OK;
}
template <int L> int launch_gadget(pgb_module *m, const GadgetArgs &p, size_t smem, int np) {
    if (np == 3) {
        if (p.aut_mode) return launch_gadget_mb<L, 3, true, 3>(m, p, smem);
        return launch_gadget_mb<L, 3, false, 3>(m, p, smem);
    }
    if (p.aut_mode) return launch_gadget_mb<L, 3, true, 4>(m, p, smem);
    if (m->opt[PGB_OPT_GADGET_MB] == 4) return launch_gadget_mb<L, 4, false, 4>(m, p, smem);
    return launch_gadget_mb<L, 3, false, 4>(m, p, smem);
}

} // namespace

bool ntt120_gadget_supported(const pgb_module *m, int R, int cols_out, int S, int base2k, int batch) {
    if (m->flavour != PGB_NTT120 || m->log_n < 10 || m->log_n > 12) return false;
    if (opt_on(m, PGB_OPT_NO_GADGET)) return false;
    const int planes = R > cols_out ? R : cols_out;
    if ((size_t)planes * m->n * 4 > (size_t)(96 << 10)) return false;
    if (cols_out < 1 || cols_out > 4 || R < 1 || R > 8) return false;
    if (S < 1 || S > 8 || base2k < 2 || base2k > 62) return false;
    if ((int64_t)S * base2k > 128 || (S - 1) * base2k + 3 >= 118) return false; // digits must come from the low 128 bits
    return batch >= 1;
}

// ok_out (device, 2 * batch + 1 ints, caller scratch): ok[b] = 1 where the ciphertext was finished here, 0 where the per-limb route is
// needed; ok[batch] = number of the latter, ok[batch + 1 ..] = their indices
//
// dsize > 1 (keyswitching/glwe.rs:332-379, external_product/glwe.rs:225-270): input limb l belongs to digit group di = dsize-1 - l % dsize
// and is the (l / dsize)-th limb of that group; the group is multiplied by the key with its limbs shifted by di (vmp limb_offset) into
// the first size_di = S - max(dsize-di-2, 0) output limbs.  By linearity the collapsed key of input poly (l, ci) is therefore
//     sum_{j < min(S - di, size_di)} 2^((S-1-j)K) key[(l / dsize) * row_cols + ci][(j + di) * cols_out + col]
// and the kernel itself is unchanged: R = a_size * row_cols input polys in their natural order.  `key_rows` = rows * cols_in of the key,
// `group_limit` = the reference's bound on the limbs of a digit group (dnum for the key-switch, none = 0 for the external product).
int ntt120_gadget_fused(pgb_module *m, const char *in, uint64_t in_bs, int in_cols, int row_cols, int row_col0, int R, const char *pmat,
                        int C, int cols_out, int small_size, char *res, uint64_t res_bs, int res_size, int base2k, int batch, int *ok_out,
                        int dsize, int a_size, int key_rows, int group_limit, int aut_mode, int64_t aut_p, int post_size) {
    const uint64_t n = m->n, poly_bytes = 16 * n;
    const int S = C / cols_out;
    if (dsize < 1) dsize = 1;
    if (dsize == 1) key_rows = R;
    // workspace: [collapsed key | key coefficients (i128) | key_bits].  For a PINNED key (pgb_gadget_key_pin) the collapsed key and the bit
    // bound live in the module's key cache and the three pre-pass launches below run once, not once per call.
    const uint64_t ck_bytes = (uint64_t)R * cols_out * poly_bytes, key_bytes = (uint64_t)key_rows * C * poly_bytes;
    const uint64_t sig[KEY_SIG_WORDS] = {1, (uint64_t)R, (uint64_t)C, (uint64_t)cols_out, (uint64_t)base2k, (uint64_t)dsize, (uint64_t)a_size,
                                         (uint64_t)key_rows, (uint64_t)group_limit, (uint64_t)row_cols};
    const bool pinned = key_is_pinned(m, pmat);
    char *cached = pinned ? (char *)key_cache_find(m, pmat, sig) : nullptr;
    const bool have = cached != nullptr;
    const uint64_t need = have ? 0 : (pinned ? key_bytes : ck_bytes + key_bytes + 256);
    if (m->aux_len < need) {
        if (m->aux_ws) {
            PGB_CHECK_CUDA(cudaStreamSynchronize(m->stream));
            cudaFree(m->aux_ws);
        }
        m->aux_ws = nullptr;
        m->aux_len = 0;
        PGB_CHECK_CUDA(cudaMalloc(&m->aux_ws, need));
        m->aux_len = need;
    }
    if (pinned && !have) PGB_TRY(key_cache_insert(m, pmat, key_bytes, sig, ck_bytes + 256, (void **)&cached));
    char *ck = pinned ? cached : (char *)m->aux_ws;
    char *kcoef = pinned ? (char *)m->aux_ws : ck + ck_bytes;
    int *key_bits = pinned ? (int *)(cached + ck_bytes) : (int *)(kcoef + key_bytes);
    CollapseArgs2 ca;
    memset(&ca, 0, sizeof ca);
    ca.pmat = pmat; ca.out = (uint32_t *)ck; ca.n = (int)n; ca.R = R; ca.C = C; ca.cols_out = cols_out; ca.S = S;
    for (int j = 0; j < S; j++)
        for (int k = 0; k < 4; k++) ca.c[j][k] = pow2_mod((uint64_t)(S - 1 - j) * base2k + 32, qk(k)); // times 2^32: Montgomery form (gadget_mac)
    for (int r = 0; r < R; r++) {
        if (dsize == 1) {
            ca.src_row[r] = (signed char)r; ca.di[r] = 0; ca.jmax[r] = (signed char)S;
            continue;
        }
        const int l = r / row_cols, ci = r % row_cols, di = dsize - 1 - l % dsize, jl = l / dsize;
        int group = (a_size + di) / dsize; // limbs of digit group di
        if (group_limit > 0 && group > group_limit) group = group_limit;
        const int src = jl * row_cols + ci;
        const int cut = dsize - di - 2, size_di = S - (cut > 0 ? cut : 0);
        int jm = S - di < size_di ? S - di : size_di;
        if (jl >= group || src >= key_rows || jm < 0) jm = 0; // this limb takes no part: its collapsed key is zero
        ca.src_row[r] = (signed char)(jm ? src : 0); ca.di[r] = (signed char)(jm ? di : 0); ca.jmax[r] = (signed char)jm;
    }
    // a pinned key's bit bound is also kept on the host: it decides, before the launch, whether three primes carry the integers
    uint64_t sig_bits[KEY_SIG_WORDS];
    memcpy(sig_bits, sig, sizeof sig_bits);
    sig_bits[0] = 5;
    int64_t *bits_slot = pinned ? key_cache_host_slot(m, pmat, key_bytes, sig_bits) : nullptr;
    if (!have) {
        { ProfScope _ps(m, PROF_OTHER);
        gadget_collapse_key_kernel<<<dim3(((unsigned)(n / 4) + 255) / 256, R * cols_out, 4), 256, 0, m->stream>>>(ca);
        }
        PGB_CHECK_CUDA(cudaGetLastError());
        PGB_TRY(ntt120_key_max_bits(m, pmat, key_rows * C, kcoef, key_bits));
        if (bits_slot) {
            int kb = 0;
            PGB_CHECK_CUDA(cudaMemcpyAsync(&kb, key_bits, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
            PGB_CHECK_CUDA(cudaStreamSynchronize(m->stream));
            *bits_slot = (int64_t)kb + 1; // 0 = not set
        }
    }
    int rn_bits = 0;
    while (((uint64_t)1 << rn_bits) < (uint64_t)R * n) rn_bits++;
    // Bound.  v = sum over R n products of an input coefficient (|a| < 2^abits) with a collapsed-key coefficient
    // (|sum_j 2^((S-1-j)K) k_j| < 2^((S-1)K + 1 + key_bits)), so |v| < 2^(rn_bits + (S-1)K + 1 + key_bits + abits).  Four primes: the
    // reference's own range, kept two bits inside Q/4 ~ 2^118 as in round 1.  Three primes: Q3 = Q[0] Q[1] Q[2] > 2^89.99 and the CRT
    // rounding needs |v| < Q3 (1/2 - 2^-27), i.e. |v| < 2^88 with room to spare.  The three-prime launch is taken when the host knows the key
    // bound (pinned key) and normalised inputs (|a| <= 2^(K-1): K bits) are certain to pass; an input beyond the bound is flagged for
    // the per-limb route by the kernel either way, so the result never depends on which launch ran.
    const int bound4 = 118 - (rn_bits + (S - 1) * base2k + 3) - 1, bound3 = 88 - (rn_bits + (S - 1) * base2k + 1);
    int np = 4;
    if (bits_slot && *bits_slot > 0 && m->opt[PGB_OPT_GADGET_PRIMES] != 4 && bound3 - (int)(*bits_slot - 1) >= base2k) np = 3;

    m->opt[PGB_OPT_LAST_GADGET_PRIMES] = np;
    GadgetArgs p;
    memset(&p, 0, sizeof p);
    p.in = in; p.in_bs = in_bs; p.res = res; p.res_bs = res_bs; p.ckey = (const uint32_t *)ck; p.key_bits = key_bits; p.ok = ok_out;
    p.in_cols = in_cols; p.row_cols = row_cols; p.row_col0 = row_col0; p.R = R; p.cols_out = cols_out; p.small_size = small_size;
    p.K = base2k; p.S = S; p.res_size = res_size; p.batch = batch;
    p.aut_mode = aut_mode; p.post_size = post_size;
    if (aut_mode) { // p^-1 mod 2n (p odd): Newton iterations double the number of correct low bits
        const uint32_t pm = (uint32_t)(((aut_p % (int64_t)(2 * n)) + (int64_t)(2 * n)) % (int64_t)(2 * n));
        if (!(pm & 1)) {
            pgb_set_error("gadget kernel: automorphism index must be odd");
            return PGB_ERR_SHAPE;
        }
        uint32_t x = pm;
        for (int it = 0; it < 5; it++) x *= 2u - pm * x;
        p.aut_pinv = x & (uint32_t)(2 * n - 1);
    }
    u128 half = 0;
    for (int j = 0; j < S; j++) half += (u128)1 << (j * base2k + base2k - 1);
    p.half_lo = (unsigned long long)half;
    p.half_hi = (unsigned long long)(half >> 64);
    p.bound_bits = np == 3 ? bound3 : bound4;
    u128 Q = 1;
    for (int k = 0; k < np; k++) Q *= qk(k);
    uint64_t ninv_n[4]; // n^-1 mod Q[k]
    for (int k = 0; k < 4; k++) {
        const uint32_t q = qk(k);
        uint64_t r = 1, b = n % q, e = q - 2;
        while (e) {
            if (e & 1) r = r * b % q;
            b = b * b % q;
            e >>= 1;
        }
        ninv_n[k] = r;
    }
    for (int k = 0; k < 4; k++) {
        const uint32_t q = qk(k), c32 = (uint32_t)((1ull << 32) % q);
        p.prime[k].q = q;
        p.prime[k].c32 = c32;
        p.prime[k].c32s = (uint32_t)(((unsigned long long)c32 << 32) / q);
        p.prime[k].neg64 = q - (uint32_t)(((unsigned long long)c32 * c32) % q);
        uint32_t qinv = q; // Newton: q^-1 mod 2^32
        for (int it = 0; it < 5; it++) qinv *= 2u - q * qinv;
        p.prime[k].qni = 0u - qinv;
        p.nq_w[k] = (uint32_t)(((u128)0 - Q) >> (32 * k));
        if (k >= np) continue;
        const u128 mk = Q / q;
        for (int w = 0; w < 4; w++) p.m_w[k][w] = (uint32_t)(mk >> (32 * w));
        p.inv60[k] = (uint32_t)(((unsigned long long)1 << 60) / q);
        if (np == 4) {
            p.crt_ninv[k] = m->nc.crt_ninv[k];
            p.crt_ninv_sh[k] = m->nc.crt_ninv_sh[k];
        } else { // (Q3 / q)^-1 / n mod q
            uint64_t r = 1, b = (uint64_t)(mk % q), e = q - 2;
            while (e) {
                if (e & 1) r = r * b % q;
                b = b * b % q;
                e >>= 1;
            }
            const uint32_t c = (uint32_t)(r * ninv_n[k] % q);
            p.crt_ninv[k] = c;
            p.crt_ninv_sh[k] = (uint32_t)(((uint64_t)c << 32) / q);
        }
    }
    const int planes = R > cols_out ? R : cols_out;
    const size_t smem = (size_t)planes * n * 4;
    PGB_CHECK_CUDA(cudaMemsetAsync(ok_out + batch, 0, sizeof(int), m->stream));
    switch (m->log_n) {
    case 10: return launch_gadget<10>(m, p, smem, np);
    case 11: return launch_gadget<11>(m, p, smem, np);
    case 12: return launch_gadget<12>(m, p, smem, np);
    default: pgb_set_error("gadget kernel: unsupported n"); return PGB_ERR_UNSUPPORTED;
    }
}
