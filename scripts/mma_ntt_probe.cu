// mma_ntt_probe.cu -- the bounded "tensor cores for the NTT?" experiment of VERDICT r1 (item 7), plus a Montgomery-vs-Shoup butterfly probe.
//
// An INT8 tensor-core NTT (tcgen05.mma.kind::i8) would run a 4096-point transform as two 64-point DFT steps: each step is a 64 x 64 matrix
// product per prime on BYTE-SPLIT residues (a 30-bit residue = 4 bytes, a twiddle = 4 bytes: 16 byte products per modular product,
// accumulated in int32 over the 64-long contraction).  The MMAs themselves are nearly free at 4.5 POP/s.  What is NOT free happens on the
// CUDA cores after every step: the 16 int32 partial sums S_ij of one output element must be recombined into sum_ij S_ij 2^(8(i+j)), reduced
// modulo q, (multiplied by the inter-step twiddle) and split into bytes again for the next step.  This probe measures exactly that
// per-element epilogue, in registers, on all SMs -- an UPPER bound on what an MMA-based NTT could reach (it assumes the MMAs, the
// TMEM -> register reads and the shared-memory staging cost nothing) -- next to the Shoup / Harvey butterfly the kernels use:
//     classic NTT, n = 4096: 12 levels x n/2 butterflies = 6 butterflies per element;
//     two-step MMA NTT:      2 epilogues per element (+ 1 twiddle product between the steps).
// Verdict rule: the MMA route is only worth building if 2 epilogues cost clearly less than 6 butterflies.
//
// Also measured: a Cooley-Tukey butterfly whose twiddle product is a Montgomery reduction (IMAD.WIDE + IMAD + IMAD.WIDE: three full-rate
// FMA-pipe slots) instead of Shoup's (IMAD.HI at half rate + 2 IMAD: four slots).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_ntt_probe mma_ntt_probe.cu && ./mma_ntt_probe > profiles/r2_mma_ntt_probe.json
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 2048
#define NCH 4

__device__ __forceinline__ uint32_t mul_shoup(uint32_t x, uint32_t w, uint32_t wp, uint32_t q) {
    uint32_t h = __umulhi(x, wp);
    return x * w - h * q;
}
__device__ __forceinline__ uint32_t csub(uint32_t x, uint32_t m) { return min(x, x - m); }
__device__ __forceinline__ uint32_t redc(unsigned long long x, uint32_t q, uint32_t qni) {
    const uint32_t m = (uint32_t)x * qni;
    return (uint32_t)((x + (unsigned long long)m * q) >> 32);
}

enum { T_EPILOGUE, T_CT_SHOUP, T_CT_MONT };

// one MMA-NTT epilogue: 16 partial sums (int32, |.| < 2^22) -> residue mod q -> x twiddle -> 4 bytes
__device__ __forceinline__ uint32_t epilogue(const uint32_t (&s)[16], const uint32_t (&c)[7], uint32_t q, uint32_t c32, uint32_t c32s, uint32_t tw, uint32_t tws) {
    // diagonals d = i + j
    const uint32_t d0 = s[0], d1 = s[1] + s[4], d2 = s[2] + s[5] + s[8], d3 = s[3] + s[6] + s[9] + s[12], d4 = s[7] + s[10] + s[13],
                   d5 = s[11] + s[14], d6 = s[15];
    // sum_d D_d * (2^(8d) mod q): seven 24-bit x 30-bit products, < 2^57 in total
    unsigned long long p = (unsigned long long)d0 * c[0] + (unsigned long long)d1 * c[1] + (unsigned long long)d2 * c[2] + (unsigned long long)d3 * c[3] +
                           (unsigned long long)d4 * c[4] + (unsigned long long)d5 * c[5] + (unsigned long long)d6 * c[6];
    const uint32_t hi = (uint32_t)(p >> 32), lo = (uint32_t)p;
    uint32_t r = csub(mul_shoup(hi, c32, c32s, q) + (lo - (lo >> 30) * q), 2 * q); // [0, 2q)
    r = csub(mul_shoup(r, tw, tws, q), q);                                            // inter-step twiddle, canonical
    return r;
}

template <int TEST> __global__ void __launch_bounds__(512, 2) k(uint32_t *out, unsigned long long *cycles, uint32_t s0, uint32_t s1, uint32_t q, uint32_t qni) {
    uint32_t x[2 * NCH];
#pragma unroll
    for (int i = 0; i < 2 * NCH; i++) x[i] = (threadIdx.x * 2654435761u + i * s0) % q;
    uint32_t c[7];
#pragma unroll
    for (int d = 0; d < 7; d++) c[d] = (uint32_t)(((1ull << (8 * d)) % q) ^ (s0 & 1)); // not compile-time constants
    const uint32_t c32 = (uint32_t)((1ull << 32) % q), c32s = (uint32_t)(((unsigned long long)c32 << 32) / q);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 2
    for (int it = 0; it < ITERS; it++) {
        if (TEST == T_EPILOGUE) {
#pragma unroll
            for (int ch = 0; ch < NCH; ch++) {
                // the 16 partial sums of one element: derived from the running value so that nothing folds (4 bytes x 4 "twiddle bytes")
                uint32_t s[16];
                const uint32_t v = x[ch];
                const uint32_t b0 = v & 255u, b1 = (v >> 8) & 255u, b2 = (v >> 16) & 255u, b3 = v >> 24; // the byte split (also part of the epilogue)
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const uint32_t wj = (s1 >> (8 * j)) & 255u;
                    s[j] = b0 * wj + (uint32_t)it; s[4 + j] = b1 * wj; s[8 + j] = b2 * wj; s[12 + j] = b3 * wj;
                }
                x[ch] = epilogue(s, c, q, c32, c32s, s0, s1);
            }
        }
        if (TEST == T_CT_SHOUP) {
#pragma unroll
            for (int i = 0; i < 2 * NCH; i += 2) {
                const uint32_t xr = csub(x[i], 2 * q);
                const uint32_t t = mul_shoup(x[i + 1], s0, s1, q);
                x[i] = xr + t;
                x[i + 1] = xr - t + 2 * q;
            }
        }
        if (TEST == T_CT_MONT) {
#pragma unroll
            for (int i = 0; i < 2 * NCH; i += 2) {
                const uint32_t xr = csub(x[i], 2 * q);
                const uint32_t t = redc((unsigned long long)x[i + 1] * s0, q, qni); // s0 plays w 2^32 mod q; result < 2q
                x[i] = xr + t;
                x[i + 1] = xr - t + 2 * q;
            }
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 2 * NCH; i++) acc ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <int TEST> static double run(const char *name, int sms, uint32_t *out, unsigned long long *cyc, double per_thread_units, bool last) {
    const int grid = sms * 2;
    const uint32_t q = 1073479681u;
    uint32_t inv = q;
    for (int i = 0; i < 5; i++) inv *= 2u - q * inv;
    k<TEST><<<grid, 512>>>(out, cyc, 12345u, 0x3a5c7e91u, q, 0u - inv);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<TEST><<<grid, 512>>>(out, cyc, 12345u, 0x3a5c7e91u, q, 0u - inv);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double total = per_thread_units * 512.0 * grid / (ms * 1e-3);
    printf("  \"%s\": {\"chip_per_s\": %.4e, \"ms\": %.4f}%s\n", name, total, ms, last ? "" : ",");
    return total;
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    uint32_t *out;
    unsigned long long *cyc;
    cudaMalloc(&out, (size_t)sms * 2 * 512 * 4);
    cudaMalloc(&cyc, (size_t)sms * 2 * 8);
    printf("{\n  \"device\": \"%s\", \"sms\": %d,\n", p.name, sms);
    const double ep = run<T_EPILOGUE>("mma_ntt_epilogue(recombine16+reduce+twiddle+split)", sms, out, cyc, (double)ITERS * NCH, false);
    const double bs = run<T_CT_SHOUP>("ct_butterfly(shoup)", sms, out, cyc, (double)ITERS * NCH, false);
    const double bm = run<T_CT_MONT>("ct_butterfly(montgomery)", sms, out, cyc, (double)ITERS * NCH, false);
    // per-element CUDA-core time of one 4096-point transform: 6 butterflies (classic) vs 2 epilogues (MMA route, MMAs / TMEM / staging free)
    printf("  \"per_element_ns_chipwide\": {\"classic_6_butterflies\": %.4e, \"mma_route_2_epilogues\": %.4e, \"montgomery_6_butterflies\": %.4e},\n",
           6.0 / bs * 1e9, 2.0 / ep * 1e9, 6.0 / bm * 1e9);
    printf("  \"mma_route_vs_classic\": %.3f,\n", (2.0 / ep) / (6.0 / bs));
    printf("  \"note\": \"ratio > 1: the CUDA-core epilogue of a two-step INT8-MMA NTT alone costs more than the whole butterfly NTT\"\n}\n");
    return 0;
}
