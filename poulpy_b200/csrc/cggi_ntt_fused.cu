// cggi_ntt_fused.cu -- CGGI blind rotation (block-binary), NTT120 flavour, as ONE persistent kernel per batch with an ADAPTIVE PRIME COUNT.
//
// Reference loop: poulpy-bin-fhe/src/blind_rotation/algorithms/cggi/algorithm.rs:338-367 (per block of `block_size` LWE coefficients:
// vec_znx_dft_apply of the accumulator, block_size x [vmp_apply_dft_to_dft, svp_apply_dft_to_dft, dft add / sub], then per column
// vec_znx_idft_apply + big_add_small_assign + big_normalize).  The NTT120 reference runs it modulo Q = q0 q1 q2 q3 (2^120) and recovers,
// at every vec_znx_idft_apply, the exact integer polynomial
//        acc_add[c] = sum_t (X^{a_t} - 1) * sum_r acc[r] (*) BRK_t[r][c]                     (negacyclic products over Z)
// as the centred representative modulo Q (reference/ntt120/arithmetic.rs:119-140).  Whenever that integer is smaller than q0 q1 / 2
// (2^59) in absolute value, the centred representative modulo Q2 = q0 q1 is THE SAME INTEGER, so two primes reconstruct the i128 of the
// reference bit for bit -- the argument that already legitimises the collapsed key of ntt120_gadget.cu.  The bound
//        block_size * 2 * R * n * max|acc| * max|key coefficient|  <  q0 q1 / 2
// is decided from DEVICE-MEASURED quantities: max|key coefficient| by an inverse transform of the whole key (cached for pinned keys),
// max|acc| checked by the kernel on every accumulator load (normalised digits are < 2^(K-1); the initial X^b * LUT may be anything) with
// a fail flag that sends the batch to the four-prime limb-wise path of cggi.cu.  At the BASELINE shape (n = 512, rank 3, block 3,
// base2k 18, 18-bit key digits) the integers are < 2^48.
//
// Organisation (the FFT64 twin is cggi_fused.cu): a CTA keeps the transform planes of G ciphertexts on chip for the whole bootstrap
// (8 polys x 2 primes x 2 KB each), 512 compute threads + ONE PRODUCER WARP that streams the bootstrapping key through a shared-memory
// ring with cp.async.bulk (a tile = the RT rows of one output poly of one key, both primes: RT contiguous 4 KB chunks of the
// [row][col][prime][n] VmpPMat); consumers release a ring stage with one mbarrier arrival per warp.  Transforms: radix-8 register passes
// through padded shared memory, 64 threads per transform synchronised by named barriers, Shoup / Harvey lazy butterflies, twiddles in
// shared memory (last pass from a conflict-free [7][T] table).  Key products: a thread owns four consecutive frequencies of one prime
// for two ciphertexts; rows accumulate as u64, the (X^{a_t} - 1) factors as u64 products of (w - 1) and a lazily reduced row sum, one
// canonical reduction per output poly.  Two-prime CRT (Garner) + add of the accumulator + base-2^K carry chain in the tail; nothing but
// the i64 accumulator (L2 resident), the LWE coefficients and the key stream touches global memory: 689 launches per batch -> 1.
#include <stdlib.h>

#include "internal.h"
#include "ntt120.cuh"
#include "tma.cuh"

using namespace n120;

namespace {

struct CggiNttArgs {
    long long *res;          uint64_t res_stride;   // GLWE VecZnx(cols, out_size), i64 words between ciphertexts
    const long long *lwe;    uint64_t lwe_stride;   // mod-switched LWE (b, a_0 .. a_{n_lwe-1}) per ciphertext
    const uint32_t *brk;     uint64_t brk_words;    // prepared GGSW i at brk + i * brk_words, layout [r][c][prime 0..3][n]
    const uint32_t *xpa;                             // x_pow_a table: 2n polys of [prime 0..3][n]
    int n_lwe, block_size, base2k, cols, dnum, brk_size, out_size, batch;
    long long acc_limit;                             // |accumulator coefficient| must stay below this (device-checked)
    int *fail;                                       // set to 1 when a loaded coefficient is not
    uint32_t q[2], qneg_inv[2];                      // primes 0 / 1, -q^-1 mod 2^32 (Montgomery)
    uint32_t cw[2], cw_sh[2];                        // 2^64 / n mod q (Shoup pair): pays the two Montgomery reductions and the 1 / n
    uint32_t q1inv, q1inv_sh;                        // q1^-1 mod q0 (Shoup pair)
    unsigned long long Q2, half2;                    // q0 q1, (q0 q1 + 1) / 2
};

__device__ __forceinline__ int NPAD(int idx) { return idx + ((idx >> 5) << 2); } // as ntt120_dft.cu: 4 extra words per 32
template <int L> struct NGeo {
    static constexpr int N = 1 << L;
    static constexpr int T = N / 8;
    static constexpr int R0 = (L % 3 == 0) ? 3 : (L % 3);
    static constexpr int PLANE = N + (N >> 5) * 4 + 4;
};

// radix-8 register passes, run-time prime, twiddles (w, floor(w 2^32 / q)) from shared memory in block-twiddle order
template <int NLEV> __device__ __forceinline__ void ct_r8(uint32_t (&x)[8], const uint2 *tw, uint32_t hi, uint32_t q) {
    {
        const uint2 w = tw[hi];
#pragma unroll
        for (int j = 0; j < 4; j++) ct_bf(x[j], x[j + 4], w, q);
    }
    if (NLEV >= 2) {
        const uint2 w0 = tw[2 * hi], w1 = tw[2 * hi + 1];
        ct_bf(x[0], x[2], w0, q);
        ct_bf(x[1], x[3], w0, q);
        ct_bf(x[4], x[6], w1, q);
        ct_bf(x[5], x[7], w1, q);
    }
    if (NLEV >= 3) {
#pragma unroll
        for (int j = 0; j < 4; j++) ct_bf(x[2 * j], x[2 * j + 1], tw[4 * hi + j], q);
    }
}
template <int NLEV> __device__ __forceinline__ void gs_r8(uint32_t (&x)[8], const uint2 *tw, uint32_t hi, uint32_t q) {
    if (NLEV >= 3) {
#pragma unroll
        for (int j = 0; j < 4; j++) gs_bf(x[2 * j], x[2 * j + 1], tw[4 * hi + j], q);
    }
    if (NLEV >= 2) {
        const uint2 w0 = tw[2 * hi], w1 = tw[2 * hi + 1];
        gs_bf(x[0], x[2], w0, q);
        gs_bf(x[1], x[3], w0, q);
        gs_bf(x[4], x[6], w1, q);
        gs_bf(x[5], x[7], w1, q);
    }
    {
        const uint2 w = tw[hi];
#pragma unroll
        for (int j = 0; j < 4; j++) gs_bf(x[j], x[j + 4], w, q);
    }
}
// last forward / first inverse pass: thread t owns node T | t; its seven twiddles come from the [7][T] table (one conflict-free load each)
__device__ __forceinline__ void ct_r8_w(uint32_t (&x)[8], const uint2 *twl, int T, int t, uint32_t q) {
    const uint2 w0 = twl[t];
#pragma unroll
    for (int j = 0; j < 4; j++) ct_bf(x[j], x[j + 4], w0, q);
    const uint2 w1 = twl[T + t], w2 = twl[2 * T + t];
    ct_bf(x[0], x[2], w1, q);
    ct_bf(x[1], x[3], w1, q);
    ct_bf(x[4], x[6], w2, q);
    ct_bf(x[5], x[7], w2, q);
#pragma unroll
    for (int j = 0; j < 4; j++) ct_bf(x[2 * j], x[2 * j + 1], twl[(3 + j) * T + t], q);
}
__device__ __forceinline__ void gs_r8_w(uint32_t (&x)[8], const uint2 *twl, int T, int t, uint32_t q) {
#pragma unroll
    for (int j = 0; j < 4; j++) gs_bf(x[2 * j], x[2 * j + 1], twl[(3 + j) * T + t], q);
    const uint2 w1 = twl[T + t], w2 = twl[2 * T + t];
    gs_bf(x[0], x[2], w1, q);
    gs_bf(x[1], x[3], w1, q);
    gs_bf(x[4], x[6], w2, q);
    gs_bf(x[5], x[7], w2, q);
    const uint2 w0 = twl[t];
#pragma unroll
    for (int j = 0; j < 4; j++) gs_bf(x[j], x[j + 4], w0, q);
}

// synchronisation among the T threads of one transform slot
template <int T> __device__ __forceinline__ void slot_sync(int slot) {
    if (T <= 32) __syncwarp();
    else named_sync(slot + 1, T);
}

// Offset of element base + j * 2^SL inside a padded plane relative to NPAD(base), for the bases the passes use (base = (a << (SL + 3)) | b
// with b < 2^SL): a compile-time constant per j, so a pass computes one padded address and the rest are immediates.
template <int SL> __device__ __forceinline__ constexpr int poff(int j) {
    return SL >= 5 ? j * ((1 << SL) + (1 << SL) / 8) : SL == 4 ? 16 * j + 4 * (j >> 1) : SL == 3 ? 8 * j + 4 * (j >> 2) : SL == 0 ? j : -1;
}
static_assert(poff<6>(3) == 216 && poff<3>(5) == 44 && poff<0>(7) == 7, "padded offsets");

// forward passes after the top one (L0 = levels done so far); the last pass leaves canonical residues
template <int L, int L0> struct NFwd {
    static __device__ __forceinline__ void run(uint32_t *buf, const uint2 *tw, const uint2 *twl, int t, int slot, bool valid, uint32_t q) {
        constexpr int SL = L - L0 - 3;
        static_assert(poff<SL>(1) > 0, "stride not covered by poff");
        if (valid) {
            const int a = t >> SL, b = t & ((1 << SL) - 1);
            uint32_t *pb = buf + NPAD((a << (SL + 3)) | b);
            uint32_t x[8];
            if (SL == 0) {
                const uint4 u0 = reinterpret_cast<const uint4 *>(pb)[0], u1 = reinterpret_cast<const uint4 *>(pb)[1];
                x[0] = u0.x; x[1] = u0.y; x[2] = u0.z; x[3] = u0.w; x[4] = u1.x; x[5] = u1.y; x[6] = u1.z; x[7] = u1.w;
                ct_r8_w(x, twl, NGeo<L>::T, t, q);
#pragma unroll
                for (int j = 0; j < 8; j++) x[j] = csub(csub(x[j], 2 * q), q);
                reinterpret_cast<uint4 *>(pb)[0] = make_uint4(x[0], x[1], x[2], x[3]);
                reinterpret_cast<uint4 *>(pb)[1] = make_uint4(x[4], x[5], x[6], x[7]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; j++) x[j] = pb[poff<SL>(j)];
                ct_r8<3>(x, tw, (1u << L0) | (uint32_t)a, q);
#pragma unroll
                for (int j = 0; j < 8; j++) pb[poff<SL>(j)] = x[j];
            }
        }
        slot_sync<NGeo<L>::T>(slot);
        NFwd<L, (L0 + 3 < L) ? L0 + 3 : L>::run(buf, tw, twl, t, slot, valid, q);
    }
};
template <int L> struct NFwd<L, L> {
    static __device__ __forceinline__ void run(uint32_t *, const uint2 *, const uint2 *, int, int, bool, uint32_t) {}
};
// inverse passes between the first (SL = 0) and the top one
template <int L, int L0> struct NInv {
    static __device__ __forceinline__ void run(uint32_t *buf, const uint2 *tw, int t, int slot, bool valid, uint32_t q) {
        constexpr int SL = L - L0 - 3;
        static_assert(poff<SL>(1) > 0, "stride not covered by poff");
        if (valid) {
            const int a = t >> SL, b = t & ((1 << SL) - 1);
            uint32_t *pb = buf + NPAD((a << (SL + 3)) | b);
            uint32_t x[8];
#pragma unroll
            for (int j = 0; j < 8; j++) x[j] = pb[poff<SL>(j)];
            gs_r8<3>(x, tw, (1u << L0) | (uint32_t)a, q);
#pragma unroll
            for (int j = 0; j < 8; j++) pb[poff<SL>(j)] = x[j];
        }
        slot_sync<NGeo<L>::T>(slot);
        NInv<L, (L0 - 3 >= NGeo<L>::R0) ? L0 - 3 : -1>::run(buf, tw, t, slot, valid, q);
    }
};
template <int L> struct NInv<L, -1> {
    static __device__ __forceinline__ void run(uint32_t *, const uint2 *, int, int, bool, uint32_t) {}
};

constexpr int CGN_COMPUTE = 512; // compute threads
// Key-stream refill.  1 (default): a 17th warp owns the ring (waits on the per-stage "empty" barriers, issues the bulk copies) -- the 17-warp
// CTA is capped at 96 registers per thread because one SM partition then holds five warps.  0: no producer at all -- the warp whose lane 0
// is the LAST of the 16 to finish a tile (shared-memory counter per stage, ordering through the "empty" barrier) issues the copy of the tile
// NSTAGE ahead into that stage; the 16-warp CTA gets 126 registers per thread and no spill.  Measured (profiles/r2_ncu_cggi.md): no faster
// here (68.1 k vs 68.6 k bootstraps/s, the kernel is bound by the FMA-heavy pipe, not by its 104 bytes of spill) and 23 % SLOWER in the
// FFT64 kernel, where the refill issued by a consumer lane stalls that consumer's warp -- so the producer warp stays.
#ifndef CGN_PRODUCER_WARP
#define CGN_PRODUCER_WARP 1
#endif
constexpr int CGN_THREADS = CGN_COMPUTE + (CGN_PRODUCER_WARP ? 32 : 0);
constexpr int CGN_BSMAX = 4;     // keys per block whose (X^a - 1) factors are held in registers

// EXACT > 0: RT and CT are the run-time row / output-poly counts and EXACT the block size (the BASELINE shape: 4, 8, 3), so every index
// computation is by constants and the (X^a - 1) factors take 8 EXACT registers instead of 8 CGN_BSMAX (the 17-warp CTA is capped at 96)
template <int L, int G, int RT, int CT, int NSTAGE, int EXACT> __global__ void __launch_bounds__(CGN_THREADS, 1)
cggi_fused_ntt120_p2_kernel(CggiNttArgs p, const uint2 *__restrict__ twf_g, const uint2 *__restrict__ twi_g) {
    typedef NGeo<L> NG;
    constexpr int N = NG::N, T = NG::T, NT = CGN_COMPUTE, PL = NG::PLANE, NSLOT = NT / T, GH = G / 2, P = 2;
    constexpr int PMAX = RT > CT ? RT : CT;
    constexpr int GS = PMAX * P * PL;                    // words per ciphertext
    constexpr int U4 = P * N / 4;                        // uint4 positions per poly (both primes)
    constexpr uint32_t CHUNK = P * N * 4, TILE = RT * CHUNK; // bytes: one key row of one output poly (primes 0 and 1 are adjacent planes)
    static_assert(GH * U4 == NT && L > NG::R0 && (GS % 4) == 0, "geometry");
    extern __shared__ __align__(128) uint32_t nsm[];
    __shared__ int s_pos[G * 8];
    __shared__ __align__(8) unsigned long long s_full[NSTAGE], s_empty[NSTAGE];
    __shared__ unsigned int s_done[NSTAGE]; // warps that have finished the tile in this stage (producer-less refill)
    uint32_t *planes = nsm;                                           // [G][PMAX][P][PL]
    uint32_t *ring = nsm + (size_t)G * GS;                            // [NSTAGE][RT][P][N]
    uint2 *tws = reinterpret_cast<uint2 *>(ring + (size_t)NSTAGE * RT * P * N); // [dir][prime][T + 7 T]: block twiddles < T, last-pass table
    const int tid = threadIdx.x;
    // twiddles: every thread (the producer warp included) helps, then one CTA-wide barrier
    for (int i = tid; i < 2 * P * 8 * T; i += CGN_THREADS) {
        const int e = i % (8 * T), k = (i / (8 * T)) % P, dir = i / (8 * T * P);
        const uint2 *src = (dir ? twi_g : twf_g) + (size_t)k * N;
        uint2 v;
        if (e < T) {
            v = src[e];
        } else {
            const int j = (e - T) / T, t = (e - T) % T, node = T + t; // j: 0 node, 1-2 children, 3-6 grandchildren
            v = j == 0 ? src[node] : (j < 3 ? src[2 * node + (j - 1)] : src[4 * node + (j - 3)]);
        }
        tws[i] = v;
    }
    const uint32_t ring_s = smem_u32(ring), full_s = smem_u32(s_full), empty_s = smem_u32(s_empty);
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; s++) {
            mbar_init(full_s + s * 8, 1);
            mbar_init(empty_s + s * 8, CGN_PRODUCER_WARP ? NT : NT / 32);
            s_done[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int cols = p.cols, C = EXACT ? CT : cols * p.brk_size, K = p.base2k, bs = EXACT ? EXACT : p.block_size;
    constexpr int BSA = EXACT ? EXACT : CGN_BSMAX;
    const int nblk = p.n_lwe / bs, total_tiles = nblk * C * bs;

    // ---- key stream: tile gk = (block, output poly c, key t) in that order ---------------------------------------------------------------
    auto issue_src = [&](const uint32_t *src, const uint32_t st_) { // one thread: bulk copies of one tile (RT key rows of one output poly)
        const uint32_t bar = full_s + st_ * 8;
        mbar_expect_tx(bar, TILE);
#pragma unroll
        for (int r = 0; r < RT; r++) bulk_g2s(ring_s + (uint32_t)(st_ * RT + r) * CHUNK, src + (size_t)r * C * 4 * N, CHUNK, bar);
    };
    auto issue_tile = [&](const uint32_t gk_, const uint32_t st_) { // the same from the tile index (producer-less mode)
        const uint32_t per_blk = (uint32_t)(C * bs), b_ = gk_ / per_blk, rem_ = gk_ - b_ * per_blk, c_ = rem_ / (uint32_t)bs, t_ = rem_ - c_ * (uint32_t)bs;
        issue_src(p.brk + ((size_t)b_ * bs + t_) * p.brk_words + (size_t)c_ * 4 * N, st_);
    };
    if (CGN_PRODUCER_WARP) {
        if (tid >= NT) {
            if (tid == NT) {
                int t = 0, c = 0;
                const uint32_t *key0 = p.brk; // first key of the current block
                for (int gk = 0; gk < total_tiles; gk++) {
                    const int st = gk % NSTAGE;
                    if (gk >= NSTAGE) {
                        const uint32_t bar = empty_s + st * 8, par = (uint32_t)((gk / NSTAGE - 1) & 1);
                        while (!mbar_try(bar, par)) __nanosleep(64); // polite: the spin would otherwise take issue slots of compute warps
                    }
                    issue_src(key0 + (size_t)t * p.brk_words + (size_t)c * 4 * N, (uint32_t)st);
                    if (++t == bs) {
                        t = 0;
                        if (++c == C) {
                            c = 0;
                            key0 += (size_t)bs * p.brk_words;
                        }
                    }
                }
            }
            return;
        }
    } else if (tid == 0) {
        for (int gk = 0; gk < NSTAGE && gk < total_tiles; gk++) issue_tile((uint32_t)gk, (uint32_t)gk);
    }

    const int slot = tid / T, t = tid % T, lane = tid & 31;
    const int ct0 = blockIdx.x * G;
    const int gp = tid / U4, u = tid % U4, kq = u / (N / 4), f4 = u % (N / 4); // products: ciphertexts gp, gp + GH; prime kq; freqs 4 f4 ..
    const uint32_t qk_ = p.q[kq], qni = p.qneg_inv[kq];
    const int mn_small = min(p.brk_size, p.out_size);
    const int a_start = min(p.out_size, p.brk_size); // same-base2k plan with offset 0: limbs >= a_start only feed the carry
    uint32_t gk = 0;
    bool bad = false;

    for (int blk = 0; blk + bs <= p.n_lwe; blk += bs) {
        if (tid < G * bs) {
            const int g = tid / bs, tt = tid % bs, ct = ct0 + g;
            const long long ai = ct < p.batch ? p.lwe[(size_t)ct * p.lwe_stride + 1 + blk + tt] : 0;
            s_pos[g * 8 + tt] = (int)((ai + (long long)(2 * N)) & (long long)(2 * N - 1));
        }
        // ---- acc_dft = NTT(acc) modulo q0 and q1 ---------------------------------------------------------------------------------
        for (int base = 0; base < G * RT * P; base += NSLOT) {
            const int job = base + slot;
            const bool valid = job < G * RT * P;
            const int g = valid ? job / (RT * P) : 0, r = valid ? (job / P) % RT : 0, k = job % P, limb = r / cols, col = r % cols;
            const uint32_t q = p.q[k];
            uint32_t *buf = planes + (size_t)g * GS + (size_t)(r * P + k) * PL;
            const uint2 *twf = tws + (size_t)k * 8 * T;
            if (valid) {
                const int ct = ct0 + g;
                const bool live = ct < p.batch && limb < p.out_size;
                const long long *src = p.res + (size_t)(live ? ct : 0) * p.res_stride + (size_t)(limb * cols + col) * N;
                uint32_t x[8];
                long long v[8];
#pragma unroll
                for (int jj = 0; jj < 8; jj++) v[jj] = live ? src[t + jj * T] : 0;
                if (blk == 0) { // only the initial accumulator (X^b * LUT) can leave the bound: later ones are digits this kernel wrote
#pragma unroll
                    for (int jj = 0; jj < 8; jj++) bad |= (v[jj] > p.acc_limit) | (v[jj] < -p.acc_limit);
                }
#pragma unroll
                for (int jj = 0; jj < 8; jj++) {
                    const int lo = (int)v[jj]; // |v| <= acc_limit <= 2^29 < q: the low word is the value
                    x[jj] = lo < 0 ? (uint32_t)lo + q : (uint32_t)lo;
                }
                ct_r8<NG::R0>(x, twf, 1u, q);
                uint32_t *pb = buf + NPAD(t);
#pragma unroll
                for (int jj = 0; jj < 8; jj++) pb[poff<L - 3>(jj)] = x[jj];
            }
            slot_sync<T>(slot);
            NFwd<L, NG::R0>::run(buf, twf, twf + T, t, slot, valid, q);
        }
        named_sync(15, NT); // every transform of the block is complete (and s_pos is visible) before the key products read across them
        // ---- key products, tiles in (output poly, key) order ---------------------------------------------------------------------
        {
            uint32_t *mine0 = planes + (size_t)gp * GS + (size_t)kq * PL + NPAD(4 * f4);
            uint32_t *mine1 = mine0 + (size_t)GH * GS;
            uint4 a0[RT], a1[RT];
#pragma unroll
            for (int r = 0; r < RT; r++) {
                a0[r] = *reinterpret_cast<const uint4 *>(mine0 + (size_t)r * P * PL);
                a1[r] = *reinterpret_cast<const uint4 *>(mine1 + (size_t)r * P * PL);
            }
            // (X^{a_t} - 1) * 2^64 / n at this thread's frequencies, per ciphertext and key: the two Montgomery reductions below each leave a
            // factor 2^-32 and the inverse transform a factor n; both are paid here, once per block, instead of per product
            uint4 w0[BSA], w1[BSA];
            const uint32_t cw = p.cw[kq], cws = p.cw_sh[kq];
            auto wfac = [&](uint32_t x) { return csub(mul_shoup(x ? x - 1 : qk_ - 1, cw, cws, qk_), qk_); };
#pragma unroll
            for (int tt = 0; tt < BSA; tt++) {
                if (tt < bs) {
                    const uint4 x0 = __ldg(reinterpret_cast<const uint4 *>(p.xpa + ((size_t)s_pos[gp * 8 + tt] * 4 + kq) * N) + f4);
                    const uint4 x1 = __ldg(reinterpret_cast<const uint4 *>(p.xpa + ((size_t)s_pos[(gp + GH) * 8 + tt] * 4 + kq) * N) + f4);
                    w0[tt] = make_uint4(wfac(x0.x), wfac(x0.y), wfac(x0.z), wfac(x0.w));
                    w1[tt] = make_uint4(wfac(x1.x), wfac(x1.y), wfac(x1.z), wfac(x1.w));
                }
            }
#pragma unroll 1
            for (int c = 0; c < C; c++) {
                unsigned long long s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
#pragma unroll
                for (int tt = 0; tt < BSA; tt++) {
                    if (tt < bs) { // uniform
                        const uint32_t st = gk % NSTAGE;
                        mbar_wait(full_s + st * 8, (gk / NSTAGE) & 1u);
                        const uint32_t *tile = ring + (size_t)st * RT * P * N + (size_t)kq * N + 4 * f4;
                        unsigned long long v0[4] = {0, 0, 0, 0}, v1[4] = {0, 0, 0, 0};
#pragma unroll
                        for (int r = 0; r < RT; r++) { // canonical a (< 2^30) x canonical key (< 2^30): RT <= 4 rows stay below 2^62
                            const uint4 kv = *reinterpret_cast<const uint4 *>(tile + (size_t)r * P * N);
                            v0[0] += (unsigned long long)a0[r].x * kv.x; v0[1] += (unsigned long long)a0[r].y * kv.y;
                            v0[2] += (unsigned long long)a0[r].z * kv.z; v0[3] += (unsigned long long)a0[r].w * kv.w;
                            v1[0] += (unsigned long long)a1[r].x * kv.x; v1[1] += (unsigned long long)a1[r].y * kv.y;
                            v1[2] += (unsigned long long)a1[r].z * kv.z; v1[3] += (unsigned long long)a1[r].w * kv.w;
                        }
                        // release the stage.  With the producer warp every thread arrives for itself once its key values are in registers (one
                        // arrival per warp after __syncwarp makes the release wait for the slowest lane; measured slower in the FFT64 kernel)
                        if (CGN_PRODUCER_WARP) {
                            mbar_arrive(empty_s + st * 8);
                        } else {
                            __syncwarp();
                            if (lane == 0) {
                                mbar_arrive(empty_s + st * 8);
                                // the last warp out refills the stage.  The counter only elects it; the ordering comes from the barrier: every
                                // warp arrived (release) before it counted itself, so the wait (acquire) below completes at once
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
                        // (s + w v) - v = s + (w - 1) v: u64 products of the canonical factor (< 2^30) and the Montgomery-reduced row sum
                        // (< 2^31): at most CGN_BSMAX = 4 terms of < 2^61 per sum
                        s0[0] += (unsigned long long)w0[tt].x * redc64(v0[0], qk_, qni); s0[1] += (unsigned long long)w0[tt].y * redc64(v0[1], qk_, qni);
                        s0[2] += (unsigned long long)w0[tt].z * redc64(v0[2], qk_, qni); s0[3] += (unsigned long long)w0[tt].w * redc64(v0[3], qk_, qni);
                        s1[0] += (unsigned long long)w1[tt].x * redc64(v1[0], qk_, qni); s1[1] += (unsigned long long)w1[tt].y * redc64(v1[1], qk_, qni);
                        s1[2] += (unsigned long long)w1[tt].z * redc64(v1[2], qk_, qni); s1[3] += (unsigned long long)w1[tt].w * redc64(v1[3], qk_, qni);
                        gk++;
                    }
                }
                // output poly c in [0, 2q) (s < 2^63: redc < 2^31 + q < 4q, one conditional subtraction), already scaled by 1 / n; the planes of
                // polys < RT were read into registers above, so writing in place is safe
                *reinterpret_cast<uint4 *>(mine0 + (size_t)c * P * PL) =
                    make_uint4(csub(redc64(s0[0], qk_, qni), 2 * qk_), csub(redc64(s0[1], qk_, qni), 2 * qk_),
                               csub(redc64(s0[2], qk_, qni), 2 * qk_), csub(redc64(s0[3], qk_, qni), 2 * qk_));
                *reinterpret_cast<uint4 *>(mine1 + (size_t)c * P * PL) =
                    make_uint4(csub(redc64(s1[0], qk_, qni), 2 * qk_), csub(redc64(s1[1], qk_, qni), 2 * qk_),
                               csub(redc64(s1[2], qk_, qni), 2 * qk_), csub(redc64(s1[3], qk_, qni), 2 * qk_));
            }
        }
        named_sync(15, NT);
        // ---- inverse transforms of the C output polys modulo both primes (the 1 / n is already in the products) ------------------------
        for (int base = 0; base < G * C * P; base += NSLOT) {
            const int job = base + slot;
            const bool valid = job < G * C * P;
            const int g = valid ? job / (C * P) : 0, qi = valid ? (job / P) % C : 0, k = job % P;
            const uint32_t q = p.q[k];
            uint32_t *buf = planes + (size_t)g * GS + (size_t)(qi * P + k) * PL;
            const uint2 *twi = tws + (size_t)(P + k) * 8 * T;
            if (valid) {
                uint32_t x[8];
                uint4 *pp = reinterpret_cast<uint4 *>(buf + NPAD(8 * t));
                const uint4 u0 = pp[0], u1 = pp[1];
                x[0] = u0.x; x[1] = u0.y; x[2] = u0.z; x[3] = u0.w; x[4] = u1.x; x[5] = u1.y; x[6] = u1.z; x[7] = u1.w;
                gs_r8_w(x, twi + T, T, t, q);
                pp[0] = make_uint4(x[0], x[1], x[2], x[3]);
                pp[1] = make_uint4(x[4], x[5], x[6], x[7]);
            }
            slot_sync<T>(slot);
            NInv<L, (L - 6 >= NG::R0) ? L - 6 : -1>::run(buf, twi, t, slot, valid, q);
            uint32_t x[8];
            uint32_t *pb = buf + NPAD(t);
            if (valid) {
#pragma unroll
                for (int jj = 0; jj < 8; jj++) x[jj] = pb[poff<L - 3>(jj)];
                gs_r8<NG::R0>(x, twi, 1u, q);
            }
            slot_sync<T>(slot); // every thread of the transform has read its inputs before the canonical values overwrite them
            if (valid) {
#pragma unroll
                for (int jj = 0; jj < 8; jj++) pb[poff<L - 3>(jj)] = csub(x[jj], q); // [0, 2q) -> canonical for the CRT
            }
        }
        named_sync(15, NT);
        // ---- two-prime CRT, + accumulator, base-2^K carry chain (brk_size <= 4 limbs), accumulator written back -----------------------
        {
            const int limb_w = cols * N, bsz = p.brk_size;
            const uint32_t q0 = p.q[0], q1 = p.q[1];
            auto crt = [&](const uint32_t *pl, int i) -> long long { // pl: plane of prime 0 of one poly; prime 1 follows at + PL
                const uint32_t v0 = pl[NPAD(i)], v1 = pl[PL + NPAD(i)];
                const uint32_t d = v0 >= v1 ? v0 - v1 : v0 + q0 - v1;                 // q1 < q0: v1 is already reduced modulo q0
                const uint32_t tq = csub(mul_shoup(d, p.q1inv, p.q1inv_sh, q0), q0);  // (v0 - v1) q1^-1 mod q0
                const unsigned long long V = (unsigned long long)v1 + (unsigned long long)q1 * tq; // in [0, q0 q1)
                return V >= p.half2 ? (long long)(V - p.Q2) : (long long)V;
            };
            const bool two_to_one = bsz == 2 && p.out_size == 1;
            for (int g = 0; g < G; g++) {
                if (ct0 + g >= p.batch) break;
                long long *acc_g = p.res + (size_t)(ct0 + g) * p.res_stride;
                const uint32_t *pg = planes + (size_t)g * GS;
                for (int col = 0; col < cols; col++) {
#pragma unroll
                    for (int i = tid; i < N; i += NT) {
                        const int o = col * N + i;
                        if (two_to_one) {
                            const long long a0v = acc_g[o];
                            const long long v0 = crt(pg + (size_t)col * P * PL, i), v1 = crt(pg + (size_t)(cols + col) * P * PL, i);
                            const long long o1 = (long long)((unsigned long long)v1 << (64 - K)) >> (64 - K);
                            const long long cy = (long long)((unsigned long long)v1 - (unsigned long long)o1) >> K;
                            const long long tsum = (long long)((unsigned long long)v0 + (unsigned long long)a0v + (unsigned long long)cy);
                            acc_g[o] = (long long)((unsigned long long)tsum << (64 - K)) >> (64 - K);
                            continue;
                        }
                        long long vv[4], aa[4];
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            vv[j] = j < bsz ? crt(pg + (size_t)(j * cols + col) * P * PL, i) : 0;
                            aa[j] = j < mn_small ? acc_g[o + j * limb_w] : 0;
                        }
                        long long cy = 0;
#pragma unroll
                        for (int j = 3; j >= 0; j--) {
                            if (j < bsz) {
                                const long long tsum = (long long)((unsigned long long)vv[j] + (unsigned long long)aa[j] + (unsigned long long)cy);
                                const long long out = (long long)((unsigned long long)tsum << (64 - K)) >> (64 - K);
                                cy = (long long)((unsigned long long)tsum - (unsigned long long)out) >> K;
                                if (j < a_start) acc_g[o + j * limb_w] = out;
                            }
                        }
                        for (int j = a_start; j < p.out_size; j++) acc_g[o + j * limb_w] = 0;
                    }
                }
            }
        }
        named_sync(15, NT); // the accumulator stores of this block are visible to the loads of the next one (same CTA, global memory)
    }
    if (bad) atomicOr(p.fail, 1);
}

template <int L, int G, int RT, int CT, int NSTAGE, int EXACT> int launch_p2(pgb_module *m, const CggiNttArgs &p) {
    typedef NGeo<L> NG;
    constexpr int PMAX = RT > CT ? RT : CT;
    const size_t smem = ((size_t)G * PMAX * 2 * NG::PLANE + (size_t)NSTAGE * RT * 2 * NG::N) * 4 + (size_t)2 * 2 * 8 * NG::T * sizeof(uint2);
    static bool attr_dev[32] = {};
    if (!attr_dev[m->device & 31]) {
        PGB_CHECK_CUDA(cudaFuncSetAttribute(cggi_fused_ntt120_p2_kernel<L, G, RT, CT, NSTAGE, EXACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_dev[m->device & 31] = true;
    }
    const int grid = (p.batch + G - 1) / G;
    { ProfScope _ps(m, PROF_OTHER);
    cggi_fused_ntt120_p2_kernel<L, G, RT, CT, NSTAGE, EXACT><<<grid, CGN_THREADS, smem, m->stream>>>(p, m->ntt_fwd, m->ntt_inv);
    }
    PGB_CHECK_CUDA(cudaGetLastError());
    return PGB_OK;
}

uint32_t modpow_u32(uint32_t b, uint32_t e, uint32_t q) {
    uint64_t r = 1, x = b % q;
    while (e) {
        if (e & 1) r = r * x % q;
        x = x * x % q;
        e >>= 1;
    }
    return (uint32_t)r;
}

__global__ void max_bits_i64_kernel(const long long *a, size_t count, int *bits) {
    int mb = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x) {
        const long long v = a[i];
        const unsigned long long mag = v < 0 ? ~(unsigned long long)v : (unsigned long long)v;
        mb = max(mb, 64 - __clzll(mag));
    }
    for (int o = 16; o; o >>= 1) mb = max(mb, __shfl_xor_sync(0xffffffffu, mb, o));
    if ((threadIdx.x & 31) == 0 && mb) atomicMax(bits, mb);
}

} // namespace

// Shapes the two-prime kernel is instantiated for: n = 512, 2 or 4 input polys, at most 8 output polys, at most CGN_BSMAX keys per block,
// key limbs <= 4 (the tail's carry chain), 16-byte aligned key polys.
bool cggi_ntt_fused_supported(const pgb_module *m, uint64_t cols, uint64_t dnum, uint64_t brk_size, uint64_t block_size) {
    if (m->flavour != PGB_NTT120 || m->log_n != 9) return false;
    const uint64_t R = cols * dnum, C = cols * brk_size;
    return (R == 2 || R == 4) && C >= 1 && C <= 8 && brk_size <= 4 && block_size >= 1 && block_size <= CGN_BSMAX;
}

// Largest bit length of the integer coefficients of `polys` prepared key polys (inverse transform + CRT in chunks through the module's
// aux workspace); *bits_host receives it (one 4-byte read-back).
static int brk_max_bits(pgb_module *m, const char *brk, uint64_t polys, int *bits_host) {
    const uint64_t poly_bytes = 16 * m->n;
    const uint64_t chunk = umin64(polys, ((uint64_t)32 << 20) / poly_bytes);
    const uint64_t need = chunk * poly_bytes + 256;
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
    char *coef = (char *)m->aux_ws;
    int *bits_dev = (int *)(coef + chunk * poly_bytes), *acc_dev = bits_dev + 1;
    PGB_CHECK_CUDA(cudaMemsetAsync(bits_dev, 0, 2 * sizeof(int), m->stream));
    int best = 0;
    for (uint64_t first = 0; first < polys; first += chunk) {
        const uint64_t cnt = umin64(chunk, polys - first);
        PGB_TRY(ntt120_key_max_bits(m, brk + first * poly_bytes, (int)cnt, coef, bits_dev)); // resets and fills bits_dev
        int b = 0;
        PGB_CHECK_CUDA(cudaMemcpyAsync(&b, bits_dev, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
        PGB_CHECK_CUDA(cudaStreamSynchronize(m->stream));
        if (b > best) best = b;
    }
    (void)acc_dev;
    *bits_host = best;
    return PGB_OK;
}

// res must already hold X^b * LUT in column 0 (and zeros elsewhere).  *handled = false: the bound or the shape sends the caller to the
// limb-wise path (res untouched); a fail flag raised by the kernel does the same after the fact (res must then be re-initialised).
int cggi_fused_ntt120(pgb_module *m, long long *res, uint64_t res_stride_words, const long long *lwe, uint64_t lwe_stride, const char *brk,
                      uint64_t brk_bytes, const char *xpa, int n_lwe, int block_size, int base2k, int cols, int dnum, int brk_size,
                      int out_size, int batch, const long long *lut, uint64_t lut_words, bool *handled) {
    *handled = false;
    const uint64_t n = m->n;
    const int R = cols * dnum, C = cols * brk_size;
    if (!cggi_ntt_fused_supported(m, cols, dnum, brk_size, block_size) || (brk_bytes % 16) != 0 || base2k > 30) return PGB_OK;
    if (m->opt[PGB_OPT_CGGI_NTT_PRIMES] == 4) return PGB_OK; // forced four-prime (limb-wise) path
    // ---- bound: key coefficients (cached on the host for a pinned key) and the LUT ---------------------------------------------------
    const uint64_t polys = (uint64_t)n_lwe * R * C;
    const uint64_t sig[KEY_SIG_WORDS] = {3, (uint64_t)n_lwe, (uint64_t)R, (uint64_t)C, 0, 0, 0, 0, 0, 0};
    int key_bits = 0;
    int64_t *slot = key_is_pinned(m, brk) ? key_cache_host_slot(m, brk, (uint64_t)n_lwe * brk_bytes, sig) : nullptr;
    if (slot && *slot > 0) {
        key_bits = (int)*slot;
    } else {
        PGB_TRY(brk_max_bits(m, brk, polys, &key_bits));
        if (slot) *slot = key_bits;
    }
    // the accumulator is the normalised LUT rotated, then normalised digits: |.| <= max(2^(K-1), max|LUT|); the kernel re-checks every load
    if (m->aux_len < 256) {
        if (m->aux_ws) {
            PGB_CHECK_CUDA(cudaStreamSynchronize(m->stream));
            cudaFree(m->aux_ws);
        }
        m->aux_ws = nullptr;
        m->aux_len = 0;
        PGB_CHECK_CUDA(cudaMalloc(&m->aux_ws, 4096));
        m->aux_len = 4096;
    }
    int *flags = (int *)((char *)m->aux_ws + m->aux_len - 64); // [lut bits | fail] at the end of the aux workspace
    PGB_CHECK_CUDA(cudaMemsetAsync(flags, 0, 2 * sizeof(int), m->stream));
    { ProfScope _ps(m, PROF_OTHER);
    max_bits_i64_kernel<<<8, 256, 0, m->stream>>>(lut, (size_t)lut_words, flags);
    }
    PGB_CHECK_CUDA(cudaGetLastError());
    int lut_bits = 0;
    PGB_CHECK_CUDA(cudaMemcpyAsync(&lut_bits, flags, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    PGB_CHECK_CUDA(cudaStreamSynchronize(m->stream));
    const int acc_bits = lut_bits > base2k - 1 ? lut_bits : base2k - 1; // |acc| <= 2^acc_bits (balanced digits reach -2^(K-1))
    int rn_bits = 0, bs_bits = 0;
    while (((uint64_t)1 << rn_bits) < (uint64_t)R * n) rn_bits++;
    while ((1 << bs_bits) < 2 * block_size) bs_bits++;
    // |acc_add| < 2^(acc_bits + key_bits + rn_bits + bs_bits) must stay below q0 q1 / 2 > 2^58.9
    if (acc_bits > 29 || acc_bits + key_bits + rn_bits + bs_bits > 58) return PGB_OK;

    CggiNttArgs p;
    memset(&p, 0, sizeof p);
    p.res = res; p.res_stride = res_stride_words; p.lwe = lwe; p.lwe_stride = lwe_stride;
    p.brk = (const uint32_t *)brk; p.brk_words = brk_bytes / 4; p.xpa = (const uint32_t *)xpa;
    p.n_lwe = n_lwe; p.block_size = block_size; p.base2k = base2k; p.cols = cols; p.dnum = dnum; p.brk_size = brk_size;
    p.out_size = out_size; p.batch = batch;
    p.acc_limit = 1ll << acc_bits;
    p.fail = flags + 1;
    for (int k = 0; k < 2; k++) {
        const uint32_t q = qk(k);
        p.q[k] = q;
        uint32_t inv = q; // Newton: q^-1 mod 2^32 (q odd), five doublings of the correct low bits
        for (int it = 0; it < 5; it++) inv *= 2u - q * inv;
        p.qneg_inv[k] = 0u - inv;
        const uint64_t r32 = ((uint64_t)1 << 32) % q, ninv = modpow_u32((uint32_t)(n % q), q - 2, q);
        p.cw[k] = (uint32_t)(r32 * r32 % q * ninv % q);
        p.cw_sh[k] = (uint32_t)(((uint64_t)p.cw[k] << 32) / q);
    }
    p.q1inv = modpow_u32(qk(1) % qk(0), qk(0) - 2, qk(0));
    p.q1inv_sh = (uint32_t)(((uint64_t)p.q1inv << 32) / qk(0));
    p.Q2 = (unsigned long long)qk(0) * qk(1);
    p.half2 = (p.Q2 + 1) / 2;
    int s;
    if (R == 4 && C == 8 && block_size == 3) s = launch_p2<9, 4, 4, 8, 4, 3>(m, p);
    else if (R == 4 && C > 4) s = launch_p2<9, 4, 4, 8, 4, 0>(m, p);
    else if (R == 4) s = launch_p2<9, 4, 4, 4, 4, 0>(m, p);
    else if (C > 4) s = launch_p2<9, 4, 2, 8, 4, 0>(m, p);
    else s = launch_p2<9, 4, 2, 4, 4, 0>(m, p);
    PGB_TRY(s);
    int fail = 0;
    PGB_CHECK_CUDA(cudaMemcpyAsync(&fail, flags + 1, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    PGB_CHECK_CUDA(cudaStreamSynchronize(m->stream));
    *handled = fail == 0;
    return PGB_OK;
}
