// pipe_peaks.cu -- issue-rate microbenchmark of the instructions the NTT120 / FFT64 kernels are made of (B200, sm_100a).
// "INT32/FP64 peaks are not in MEASURED_PEAKS.json: builder must commit a microbenchmark (IMAD.WIDE and DFMA
// issue rate) before quoting fractions."  Every test runs 1024 resident threads per SM (2 CTAs x 512, 64 registers each) on all SMs, each thread
// NCHAIN independent dependency chains, and reports the chip-wide rate from CUDA events (chip_per_s: what the roofline uses; the SM
// clock sags to ~1.4 GHz under these all-SM integer loops) and thread-level operations per clock per SM from clock64().
// (Plain add / min chains are folded by ptxas and are not reported.)
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_peaks pipe_peaks.cu && ./pipe_peaks > pipe_peaks.json
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define NCHAIN 8
#define ITERS 4096

enum { OP_IMAD_LO, OP_IMAD_HI, OP_IMAD_WIDE, OP_IADD, OP_IADD3, OP_UMIN, OP_SHF, OP_LOP3, OP_DFMA, OP_DADD, OP_DMUL, OP_FFMA,
       OP_SHOUP, OP_CT_BF, OP_GS_BF, OP_CT_BF_SIGN, OP_MIX_IMAD_IADD, OP_COUNT };
static const char *op_name[] = {"imad_lo_u32", "imad_hi_u32", "imad_wide_u32", "iadd_u32", "iadd3_u32", "umin_u32", "shf_r_u32", "lop3_u32",
                                "dfma", "dadd", "dmul", "ffma", "shoup_modmul(3 imad)", "ct_butterfly(harvey,shoup)", "gs_butterfly(shoup)",
                                "ct_butterfly(sign-bit csub)", "imad_lo+iadd alternating"};
// thread-level "operations" per chain step (for the composite tests: one modmul / one butterfly)
static const double op_unit[] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 2};

__device__ __forceinline__ uint32_t mul_shoup(uint32_t x, uint32_t w, uint32_t wp, uint32_t q) {
    uint32_t h = __umulhi(x, wp);
    return x * w - h * q;
}
__device__ __forceinline__ uint32_t csub(uint32_t x, uint32_t m) { return min(x, x - m); }

template <int OP> __global__ void __launch_bounds__(512, 2) k(uint32_t *out, unsigned long long *cycles, uint32_t s0, uint32_t s1, uint32_t q) {
    uint32_t x[NCHAIN], y[NCHAIN];
    double d[NCHAIN];
    float f[NCHAIN];
    unsigned long long w[NCHAIN];
#pragma unroll
    for (int i = 0; i < NCHAIN; i++) {
        x[i] = threadIdx.x * 2654435761u + i * s0;
        y[i] = x[i] ^ s1;
        d[i] = (double)x[i] * 1e-9;
        f[i] = (float)x[i] * 1e-9f;
        w[i] = x[i];
    }
    const double da = (double)s0 * 1e-3, db = (double)s1 * 1e-3;
    const float fa = (float)s0 * 1e-3f, fb = (float)s1 * 1e-3f;
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NCHAIN; i++) {
            if (OP == OP_IMAD_LO) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(s0), "r"(s1));
            if (OP == OP_IMAD_HI) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(s0));
            if (OP == OP_IMAD_WIDE) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x[i]), "r"(s0));
            if (OP == OP_IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(s0));
            if (OP == OP_IADD3) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(x[i]) : "r"(s0), "r"(y[i]));
            if (OP == OP_UMIN) asm volatile("min.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
            if (OP == OP_SHF) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(s0));
            if (OP == OP_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(s0));
            if (OP == OP_DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(da), "d"(db));
            if (OP == OP_DADD) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(da));
            if (OP == OP_DMUL) asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(da));
            if (OP == OP_FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb));
            if (OP == OP_SHOUP) x[i] = mul_shoup(x[i], s0, s1, q);
            if (OP == OP_MIX_IMAD_IADD) {
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(s0), "r"(s1));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(s0));
            }
        }
        if (OP == OP_CT_BF) {
#pragma unroll
            for (int i = 0; i < NCHAIN; i += 2) {
                uint32_t xr = csub(x[i], 2 * q);
                uint32_t t = mul_shoup(x[i + 1], s0, s1, q);
                x[i] = xr + t;
                x[i + 1] = xr - t + 2 * q;
            }
        }
        if (OP == OP_CT_BF_SIGN) {
#pragma unroll
            for (int i = 0; i < NCHAIN; i += 2) {
                uint32_t xr = x[i] - (x[i] >> 31) * (2 * q);
                uint32_t t = mul_shoup(x[i + 1], s0, s1, q);
                x[i] = xr + t;
                x[i + 1] = xr - t + 2 * q;
            }
        }
        if (OP == OP_GS_BF) {
#pragma unroll
            for (int i = 0; i < NCHAIN; i += 2) {
                uint32_t s = csub(x[i] + x[i + 1], 2 * q);
                uint32_t dd = x[i] - x[i + 1] + 2 * q;
                x[i] = s;
                x[i + 1] = mul_shoup(dd, s0, s1, q);
            }
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < NCHAIN; i++) acc ^= x[i] ^ y[i] ^ (uint32_t)__double_as_longlong(d[i]) ^ __float_as_uint(f[i]) ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

template <int OP> static void run(int sms, uint32_t *out, unsigned long long *cyc, unsigned long long *hcyc, bool last) {
    const int grid = sms * 2;
    k<OP><<<grid, 512>>>(out, cyc, 12345u, 6789u, 1073479681u);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<grid, 512>>>(out, cyc, 12345u, 6789u, 1073479681u);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(hcyc, cyc, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < grid; i++) avg += (double)hcyc[i];
    avg /= grid;
    const bool pairs = (OP == OP_CT_BF || OP == OP_GS_BF || OP == OP_CT_BF_SIGN);
    const double per_thread = (double)ITERS * (pairs ? NCHAIN / 2 : NCHAIN) * op_unit[OP];
    const double per_clk_sm = per_thread * 1024.0 / avg;
    const double total_per_s = per_thread * 512.0 * grid / (ms * 1e-3);
    printf("  \"%s\": {\"per_clk_per_sm\": %.2f, \"chip_per_s\": %.4e, \"ms\": %.4f}%s\n", op_name[OP], per_clk_sm, total_per_s, ms, last ? "" : ",");
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    uint32_t *out;
    unsigned long long *cyc, *hcyc = new unsigned long long[sms * 2];
    cudaMalloc(&out, (size_t)sms * 2 * 512 * 4);
    cudaMalloc(&cyc, (size_t)sms * 2 * 8);
    printf("{\n  \"device\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d, \"threads_per_sm\": 1024, \"chains_per_thread\": %d,\n", p.name, sms, p.clockRate, NCHAIN);
    run<OP_IMAD_LO>(sms, out, cyc, hcyc, false);
    run<OP_IMAD_HI>(sms, out, cyc, hcyc, false);
    run<OP_IMAD_WIDE>(sms, out, cyc, hcyc, false);
    run<OP_IADD3>(sms, out, cyc, hcyc, false);
    run<OP_SHF>(sms, out, cyc, hcyc, false);
    run<OP_LOP3>(sms, out, cyc, hcyc, false);
    run<OP_MIX_IMAD_IADD>(sms, out, cyc, hcyc, false);
    run<OP_DFMA>(sms, out, cyc, hcyc, false);
    run<OP_DADD>(sms, out, cyc, hcyc, false);
    run<OP_DMUL>(sms, out, cyc, hcyc, false);
    run<OP_FFMA>(sms, out, cyc, hcyc, false);
    run<OP_SHOUP>(sms, out, cyc, hcyc, false);
    run<OP_CT_BF>(sms, out, cyc, hcyc, false);
    run<OP_CT_BF_SIGN>(sms, out, cyc, hcyc, false);
    run<OP_GS_BF>(sms, out, cyc, hcyc, true);
    printf("}\n");
    return 0;
}
