// ntt120.cuh -- device arithmetic of the NTT120 flavour (four ~30-bit primes, canonical u32 residues).
//
// Semantics follow poulpy-cpu-ref/src/reference/ntt120/{primes.rs:80-90, arithmetic.rs:39-60,119-140,
// ntt.rs:558-684}; the representation does not: the reference keeps lazy 64-bit residues (q120b) with a
// bit-growth schedule, this backend keeps canonical 32-bit residues with Shoup/Harvey lazy butterflies in
// [0, 4q).  The results agree because the reference canonicalises (x % Q[k]) before CRT
// (arithmetic.rs:132), so any residue representation yields the same i128.
#pragma once
#include <stdint.h>

namespace n120 {

// primes.rs:80-90 (Primes30)
template <int K> struct Prime;
template <> struct Prime<0> { static constexpr uint32_t q = (1u << 30) - 2u * (1u << 17) + 1u; };
template <> struct Prime<1> { static constexpr uint32_t q = (1u << 30) - 17u * (1u << 17) + 1u; };
template <> struct Prime<2> { static constexpr uint32_t q = (1u << 30) - 23u * (1u << 17) + 1u; };
template <> struct Prime<3> { static constexpr uint32_t q = (1u << 30) - 42u * (1u << 17) + 1u; };

__host__ __device__ __forceinline__ constexpr uint32_t qk(int k) {
    return k == 0 ? Prime<0>::q : k == 1 ? Prime<1>::q : k == 2 ? Prime<2>::q : Prime<3>::q;
}
static constexpr uint32_t OMEGA[4] = {1070907127u, 315046632u, 309185662u, 846468380u};
static constexpr uint32_t CRT_CST[4] = {43599465u, 292938863u, 594011630u, 140177212u};

// Montgomery reduction of a u64: x * 2^-32 mod q as a value in [0, x / 2^32 + q) -- two instructions (IMAD + IMAD.WIDE) where a Shoup-style
// reduction of a 64-bit value takes seven.  The factor 2^-32 is compensated in the constants the value is multiplied with (before or after).
// qneg_inv = -q^-1 mod 2^32.  x + m q is a multiple of 2^32 and < 2^64 whenever x < 2^63 (m q < 2^62).
__device__ __forceinline__ uint32_t redc64(unsigned long long x, uint32_t q, uint32_t qneg_inv) {
    const uint32_t m = (uint32_t)x * qneg_inv;
    return (uint32_t)((x + (unsigned long long)m * q) >> 32);
}

// x * w mod q in [0, 2q) for any x < 2^32, w < q, wp = floor(w * 2^32 / q)   (Shoup)
// -DPGB_SHOUP_WIDE takes the high word from a full 32x32->64 product (IMAD.WIDE, full IMAD rate in isolation) instead of __umulhi
// (IMAD.HI, half rate: profiles/r1_pipe_peaks.json); ptxas keeps the wide form when it is spelled as mul.wide + unpack.  Measured on the
// key-switch kernel: no gain (1.564 vs 1.541 ms per 4096), so the default stays __umulhi.
__device__ __forceinline__ uint32_t umulhi_wide(uint32_t a, uint32_t b) {
#ifndef PGB_SHOUP_WIDE
    return __umulhi(a, b);
#else
    uint32_t lo, hi;
    asm("{ .reg .u64 t; mul.wide.u32 t, %2, %3; mov.b64 {%0, %1}, t; }" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
    (void)lo;
    return hi;
#endif
}
__device__ __forceinline__ uint32_t mul_shoup(uint32_t x, uint32_t w, uint32_t wp, uint32_t q) {
    uint32_t h = umulhi_wide(x, wp);
    return x * w - h * q;
}
__device__ __forceinline__ uint32_t csub(uint32_t x, uint32_t m) { return min(x, x - m); } // x in [0, 2m) -> [0, m)

// Harvey lazy Cooley-Tukey butterfly: x, y in [0, 4q) -> (x + w*y, x - w*y) in [0, 4q)
__device__ __forceinline__ void ct_bf(uint32_t &x, uint32_t &y, uint2 w, uint32_t q) {
    uint32_t xr = csub(x, 2 * q);
    uint32_t t = mul_shoup(y, w.x, w.y, q);
    x = xr + t;
    y = xr - t + 2 * q;
}
// Lazy Gentleman-Sande butterfly: x, y in [0, 2q) -> (x + y, (x - y) * w) in [0, 2q)
__device__ __forceinline__ void gs_bf(uint32_t &x, uint32_t &y, uint2 w, uint32_t q) {
    uint32_t s = csub(x + y, 2 * q);
    uint32_t d = x - y + 2 * q;
    x = s;
    y = mul_shoup(d, w.x, w.y, q);
}

// i64 -> residue mod q in [0, 3q) (exact for the full i64 range, like arithmetic.rs:39-60 followed by % q; the lazy
// butterflies accept [0, 4q)).  x = uh * 2^32 + ul - [x < 0] * 2^64 with uh, ul the unsigned halves:
//   uh * (2^32 mod q) by Shoup -> [0, 2q);  ul - (ul >> 30) * q -> [0, 2q);  correction q - (2^64 mod q) for negatives.
template <int K> __device__ __forceinline__ uint32_t from_i64(long long v) {
    constexpr uint32_t q = Prime<K>::q;
    constexpr uint32_t c32 = (uint32_t)((1ull << 32) % q);
    constexpr uint32_t c32s = (uint32_t)(((unsigned long long)c32 << 32) / q);
    constexpr uint32_t c64 = (uint32_t)(((unsigned long long)c32 * c32) % q);
    const uint32_t uh = (uint32_t)((unsigned long long)v >> 32), ul = (uint32_t)v;
    uint32_t r = csub(mul_shoup(uh, c32, c32s, q) + (ul - (ul >> 30) * q), 2 * q); // [0, 2q)
    if (v < 0) r += q - c64;                                                        // [0, 3q)
    return r;
}

// u64 -> canonical residue, q compile-time so the compiler emits a multiply-high sequence
template <int K> __device__ __forceinline__ uint32_t red64(unsigned long long x) {
    return (uint32_t)(x % Prime<K>::q);
}
// runtime-prime variants: per-thread constants selected once from k
__host__ __device__ __forceinline__ constexpr uint32_t c32k(int k) { return (uint32_t)((1ull << 32) % qk(k)); }
__host__ __device__ __forceinline__ constexpr uint32_t c32sk(int k) { return (uint32_t)(((unsigned long long)c32k(k) << 32) / qk(k)); }
struct PrimeRt {
    uint32_t q, c32, c32s;
    __device__ __forceinline__ explicit PrimeRt(int k)
        : q(k == 0 ? qk(0) : k == 1 ? qk(1) : k == 2 ? qk(2) : qk(3)),
          c32(k == 0 ? c32k(0) : k == 1 ? c32k(1) : k == 2 ? c32k(2) : c32k(3)),
          c32s(k == 0 ? c32sk(0) : k == 1 ? c32sk(1) : k == 2 ? c32sk(2) : c32sk(3)) {}
    // any u64 -> canonical residue: hi * (2^32 mod q) by Shoup + folded low word, then two conditional subtractions
    __device__ __forceinline__ uint32_t reduce(unsigned long long x) const {
        const uint32_t hi = (uint32_t)(x >> 32), lo = (uint32_t)x;
        return csub(csub(mul_shoup(hi, c32, c32s, q) + (lo - (lo >> 30) * q), 2 * q), q);
    }
};
__device__ __forceinline__ uint32_t red64k(unsigned long long x, int k) {
    switch (k) {
    case 0: return red64<0>(x);
    case 1: return red64<1>(x);
    case 2: return red64<2>(x);
    default: return red64<3>(x);
    }
}

} // namespace n120

// CRT reconstruction constants (arithmetic.rs:119-140): M_k = Q / Q[k] as (lo, hi) 64-bit words, Q, (Q+1)/2
struct CrtConsts {
    unsigned long long m_lo[4], m_hi[4];
    unsigned long long q_lo, q_hi;
    unsigned long long half_lo, half_hi;
};
