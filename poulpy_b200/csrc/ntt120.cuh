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

// x * w mod q in [0, 2q) for any x < 2^32, w < q, wp = floor(w * 2^32 / q)   (Shoup)
__device__ __forceinline__ uint32_t mul_shoup(uint32_t x, uint32_t w, uint32_t wp, uint32_t q) {
    uint32_t h = __umulhi(x, wp);
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

// i64 -> canonical residue mod q (exact for the full i64 range, like arithmetic.rs:39-60 followed by % q)
template <int K> __device__ __forceinline__ uint32_t from_i64(long long v) {
    constexpr uint32_t q = Prime<K>::q;
    constexpr uint32_t two63 = (uint32_t)((1ull << 63) % q);
    unsigned long long u = (unsigned long long)v;
    uint32_t r = (uint32_t)((u & 0x7FFFFFFFFFFFFFFFull) % q);
    if (v < 0) r = csub(r + (q - two63), q);
    return r;
}

// u64 -> canonical residue, q compile-time so the compiler emits a multiply-high sequence
template <int K> __device__ __forceinline__ uint32_t red64(unsigned long long x) {
    return (uint32_t)(x % Prime<K>::q);
}
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
