// fft64.cuh -- device helpers shared by the FFT64 kernels (fft64.cu) and the fused CGGI kernel (cggi_fused.cu).
#pragma once
#include <stdint.h>

__device__ __forceinline__ int FPAD(int idx) { return idx + (idx >> 3); }
template <int L> struct FGeo {
    static constexpr int M = 1 << L;
    static constexpr int T = M >= 8 ? M / 8 : 1;
    static constexpr int R0 = (L % 3 == 0) ? 3 : (L % 3);
    static constexpr int PLANE = M + (M >> 3) + 2; // padded double2 elements
};

__device__ __forceinline__ void fct_bf(double2 &x, double2 &y, double2 w) { // (x + w*y, x - w*y)
    double dr = y.x * w.x - y.y * w.y;
    double di = y.x * w.y + y.y * w.x;
    y = make_double2(x.x - dr, x.y - di);
    x = make_double2(x.x + dr, x.y + di);
}
__device__ __forceinline__ void fgs_bf(double2 &x, double2 &y, double2 w) { // (x + y, (x - y)*w)
    double rd = x.x - y.x, id = x.y - y.y;
    x = make_double2(x.x + y.x, x.y + y.y);
    y = make_double2(rd * w.x - id * w.y, rd * w.y + id * w.x);
}
// twiddle load: read-only global tables go through the non-coherent path (TWS = false), shared-memory copies (fused CGGI
// kernel, TWS = true) through plain loads
template <bool TWS = false> __device__ __forceinline__ double2 ldw(const double2 *p) {
    if (TWS) return *p;
    return __ldg(p); // one 128-bit load
}

template <int NLEV, bool TWS = false> __device__ __forceinline__ void fct_radix8(double2 (&x)[8], const double2 *__restrict__ tw, uint32_t hi) {
    {
        double2 w = ldw<TWS>(tw + hi);
#pragma unroll
        for (int j = 0; j < 4; j++) fct_bf(x[j], x[j + 4], w);
    }
    if (NLEV >= 2) {
        double2 w0 = ldw<TWS>(tw + 2 * hi), w1 = ldw<TWS>(tw + 2 * hi + 1);
        fct_bf(x[0], x[2], w0);
        fct_bf(x[1], x[3], w0);
        fct_bf(x[4], x[6], w1);
        fct_bf(x[5], x[7], w1);
    }
    if (NLEV >= 3) {
#pragma unroll
        for (int j = 0; j < 4; j++) fct_bf(x[2 * j], x[2 * j + 1], ldw<TWS>(tw + 4 * hi + j));
    }
}
template <int NLEV, bool TWS = false> __device__ __forceinline__ void fgs_radix8(double2 (&x)[8], const double2 *__restrict__ tw, uint32_t hi) {
    if (NLEV >= 3) {
#pragma unroll
        for (int j = 0; j < 4; j++) fgs_bf(x[2 * j], x[2 * j + 1], ldw<TWS>(tw + 4 * hi + j));
    }
    if (NLEV >= 2) {
        double2 w0 = ldw<TWS>(tw + 2 * hi), w1 = ldw<TWS>(tw + 2 * hi + 1);
        fgs_bf(x[0], x[2], w0);
        fgs_bf(x[1], x[3], w0);
        fgs_bf(x[4], x[6], w1);
        fgs_bf(x[5], x[7], w1);
    }
    {
        double2 w = ldw<TWS>(tw + hi);
#pragma unroll
        for (int j = 0; j < 4; j++) fgs_bf(x[j], x[j + 4], w);
    }
}


// The last forward pass and the first inverse pass give thread t the node hi = T | t (T = m/8): its seven twiddles tw[hi], tw[2hi],
// tw[2hi+1], tw[4hi..4hi+3] sit 16 / 32 / 64 bytes apart between neighbouring threads, so a warp-wide 128-bit load of them touches up to
// 4x the lines (or shared-memory banks) it needs.  The module keeps the same values a second time as [7][T] ("last-pass tables",
// fft64_module_init): one coalesced / conflict-free load each.
template <bool TWS = false> __device__ __forceinline__ void load_tw7(double2 (&w)[7], const double2 *__restrict__ twl, int T, int t) {
#pragma unroll
    for (int j = 0; j < 7; j++) w[j] = ldw<TWS>(twl + j * T + t);
}
__device__ __forceinline__ void fct_radix8_w(double2 (&x)[8], const double2 (&w)[7]) {
#pragma unroll
    for (int j = 0; j < 4; j++) fct_bf(x[j], x[j + 4], w[0]);
    fct_bf(x[0], x[2], w[1]);
    fct_bf(x[1], x[3], w[1]);
    fct_bf(x[4], x[6], w[2]);
    fct_bf(x[5], x[7], w[2]);
#pragma unroll
    for (int j = 0; j < 4; j++) fct_bf(x[2 * j], x[2 * j + 1], w[3 + j]);
}
__device__ __forceinline__ void fgs_radix8_w(double2 (&x)[8], const double2 (&w)[7]) {
#pragma unroll
    for (int j = 0; j < 4; j++) fgs_bf(x[2 * j], x[2 * j + 1], w[3 + j]);
    fgs_bf(x[0], x[2], w[1]);
    fgs_bf(x[1], x[3], w[1]);
    fgs_bf(x[4], x[6], w[2]);
    fgs_bf(x[5], x[7], w[2]);
#pragma unroll
    for (int j = 0; j < 4; j++) fgs_bf(x[j], x[j + 4], w[0]);
}

// one step of the same-base2k carry chain on i64 (znx_normalize_{first,middle}_step with lsh = 0, normalization.rs:24-323): the digit
// of x and the digit of (digit + carry_in) are taken separately, exactly like the reference, so wrap-around cases agree too
__device__ __forceinline__ long long norm_step(long long x, long long &c, int K) {
    const long long d = (long long)((unsigned long long)x << (64 - K)) >> (64 - K);
    const long long co = (long long)((unsigned long long)x - (unsigned long long)d) >> K;
    const long long s = (long long)((unsigned long long)d + (unsigned long long)c);
    const long long out = (long long)((unsigned long long)s << (64 - K)) >> (64 - K);
    c = (long long)((unsigned long long)co + (unsigned long long)((long long)((unsigned long long)s - (unsigned long long)out) >> K));
    return out;
}
