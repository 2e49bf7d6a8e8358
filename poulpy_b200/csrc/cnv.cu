// cnv.cu -- bivariate convolution (HalImpl::cnv_*, poulpy-hal/src/oep/hal_impl.rs:670-754): the limb x limb products behind
// glwe_tensor_apply / CKKS multiplication.
//
// Reference: poulpy-cpu-ref/src/reference/ntt120/convolution.rs (prepare :66-236, apply_dft :256-335, by_const :361-410,
// pairwise :441-557) and reference/fft64/convolution.rs (:13-137, :199-249, :144-191, :256-334).
// CnvPVecL / CnvPVecR are opaque prepared layouts; here both are plain DFT limbs in the VecZnxDft layout of this backend
// (limb-major, column-minor, 16 B / 8 B per coefficient), so "prepare" is a forward transform with the last active limb masked.
//   res[k] = sum_{j = j_min}^{j_max - 1} a[k_abs - j] (.) b[j],   k_abs = k + min(cnv_offset, a.size + b.size - 1)
// per frequency (and prime).  The NTT120 kernel accumulates u32 x u32 products in u64 and reduces once per 16 terms; results
// are canonical residues, which is all the reference's CRT sees (arithmetic.rs:132).  (A register-tiled variant that loads every
// limb once per thread measured no faster at the CKKS shape -- the kernel is bound by streaming the operands, which L2 already
// de-duplicates across the output limbs of one ciphertext -- and was dropped.)
#include <stdlib.h>

#include "internal.h"
#include "ntt120.cuh"

using namespace n120;

struct CnvArgs {
    LimbSet res, a, a2, b, b2; // a2 / b2: second column of the pairwise form (base == nullptr: plain form)
    uint32_t n;                // ring degree
    int a_size, b_size, offset, min_size;
};

// one thread per uint4 (four consecutive frequencies of one prime plane); blockIdx.y = output limb, blockIdx.z = batch item
__global__ void __launch_bounds__(256) ntt120_cnv_kernel(CnvArgs p) {
    const uint32_t u = blockIdx.x * blockDim.x + threadIdx.x; // uint4 index inside a limb (n words per prime, 4 primes)
    if (u >= p.n) return;
    const PrimeRt pr(u / (p.n / 4));
    const uint32_t q = pr.q;
    const int k = blockIdx.y, k_abs = k + p.offset;
    const size_t bz = blockIdx.z;
    const int j_min = k_abs > p.a_size - 1 ? k_abs - (p.a_size - 1) : 0, j_max = min(k_abs + 1, p.b_size);
    unsigned long long acc[4] = {0, 0, 0, 0};
    int cnt = 0;
    for (int j = j_min; j < j_max; j++) {
        uint4 x = __ldg(reinterpret_cast<const uint4 *>(p.a.base + bz * p.a.batch_stride + (size_t)(k_abs - j) * p.a.limb_stride) + u);
        uint4 y = __ldg(reinterpret_cast<const uint4 *>(p.b.base + bz * p.b.batch_stride + (size_t)j * p.b.limb_stride) + u);
        if (p.a2.base) { // pairwise: (a_i + a_j) (.) (b_i + b_j), sums reduced to [0, q)
            const uint4 x2 = __ldg(reinterpret_cast<const uint4 *>(p.a2.base + bz * p.a2.batch_stride + (size_t)(k_abs - j) * p.a2.limb_stride) + u);
            const uint4 y2 = __ldg(reinterpret_cast<const uint4 *>(p.b2.base + bz * p.b2.batch_stride + (size_t)j * p.b2.limb_stride) + u);
            x = make_uint4(csub(x.x + x2.x, q), csub(x.y + x2.y, q), csub(x.z + x2.z, q), csub(x.w + x2.w, q));
            y = make_uint4(csub(y.x + y2.x, q), csub(y.y + y2.y, q), csub(y.z + y2.z, q), csub(y.w + y2.w, q));
        }
        acc[0] += (unsigned long long)x.x * y.x; acc[1] += (unsigned long long)x.y * y.y;
        acc[2] += (unsigned long long)x.z * y.z; acc[3] += (unsigned long long)x.w * y.w;
        if (++cnt == 16) { // 16 products < 2^60 each
            cnt = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) acc[i] = pr.reduce(acc[i]);
        }
    }
    uint4 *dst = reinterpret_cast<uint4 *>(p.res.base + bz * p.res.batch_stride + (size_t)k * p.res.limb_stride) + u;
    *dst = make_uint4(pr.reduce(acc[0]), pr.reduce(acc[1]), pr.reduce(acc[2]), pr.reduce(acc[3]));
}

// FFT64: one thread per complex frequency; sums in ascending j with the operation order of reim4_add_mul
// (reim4/arithmetic_ref.rs:223-232, no contraction across the real / imaginary updates beyond the FMA the butterflies also use)
__global__ void __launch_bounds__(256) fft64_cnv_kernel(CnvArgs p) {
    const uint32_t f = blockIdx.x * blockDim.x + threadIdx.x, m = p.n / 2;
    if (f >= m) return;
    const int k = blockIdx.y, k_abs = k + p.offset;
    const size_t bz = blockIdx.z;
    const int j_min = k_abs > p.a_size - 1 ? k_abs - (p.a_size - 1) : 0, j_max = min(k_abs + 1, p.b_size);
    double rr = 0.0, ri = 0.0;
    for (int j = j_min; j < j_max; j++) {
        const double *x = reinterpret_cast<const double *>(p.a.base + bz * p.a.batch_stride + (size_t)(k_abs - j) * p.a.limb_stride);
        const double *y = reinterpret_cast<const double *>(p.b.base + bz * p.b.batch_stride + (size_t)j * p.b.limb_stride);
        double ar = __ldg(x + f), ai = __ldg(x + f + m), br = __ldg(y + f), bi = __ldg(y + f + m);
        if (p.a2.base) {
            const double *x2 = reinterpret_cast<const double *>(p.a2.base + bz * p.a2.batch_stride + (size_t)(k_abs - j) * p.a2.limb_stride);
            const double *y2 = reinterpret_cast<const double *>(p.b2.base + bz * p.b2.batch_stride + (size_t)j * p.b2.limb_stride);
            ar += __ldg(x2 + f); ai += __ldg(x2 + f + m); br += __ldg(y2 + f); bi += __ldg(y2 + f + m);
        }
        rr += __dsub_rn(__dmul_rn(ar, br), __dmul_rn(ai, bi));
        ri += __dadd_rn(__dmul_rn(ar, bi), __dmul_rn(ai, br));
    }
    double *dst = reinterpret_cast<double *>(p.res.base + bz * p.res.batch_stride + (size_t)k * p.res.limb_stride);
    dst[f] = rr;
    dst[f + m] = ri;
}

int cnv_apply(pgb_module *m, LimbSet res, int res_size, LimbSet a, LimbSet a2, int a_size, LimbSet b, LimbSet b2, int b_size, uint64_t cnv_offset,
              uint32_t batch) {
    const uint64_t n = m->n, pb = prep_bytes(m);
    if (res_size == 0 || batch == 0) return PGB_OK;
    if (a_size == 0 || b_size == 0) return raw_limbs(m, true, res, res, n * pb, (uint32_t)res_size, batch);
    const int bound = a_size + b_size - 1;
    const int offset = (int)umin64(cnv_offset, (uint64_t)bound);
    // NTT120: min(res.size, bound + 1 - offset) (convolution.rs:283-285); FFT64: min(res.size, bound) (fft64/convolution.rs:224-226) --
    // the limbs in between are empty sums, i.e. zero either way
    const int min_size = res_size < bound + 1 - offset ? res_size : bound + 1 - offset;
    CnvArgs p = {res, a, a2, b, b2, (uint32_t)n, a_size, b_size, offset, min_size};
    if (min_size > 0) {
        ProfScope _ps(m, PROF_VMP);
        if (m->flavour == PGB_NTT120) ntt120_cnv_kernel<<<dim3(((uint32_t)n + 255) / 256, min_size, batch), 256, 0, m->stream>>>(p);
        else fft64_cnv_kernel<<<dim3(((uint32_t)(n / 2) + 255) / 256, min_size, batch), 256, 0, m->stream>>>(p);
        PGB_CHECK_CUDA(cudaGetLastError());
    }
    LimbSet z = res;
    z.base += (size_t)min_size * res.limb_stride;
    return raw_limbs(m, true, z, z, n * pb, (uint32_t)(res_size - min_size), batch);
}

// ---- by_const: coefficient domain, i128 (NTT120 big) or wrapping i64 (FFT64 big) accumulators -------------------------------
struct CnvConstArgs {
    LimbSet res, a;
    const long long *b; // device copy of the constant limbs
    uint32_t n;
    int a_size, b_size, offset;
    int big_is_i128;
};
__global__ void __launch_bounds__(256) cnv_by_const_kernel(CnvConstArgs p) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n) return;
    const int k = blockIdx.y, k_abs = k + p.offset;
    const size_t bz = blockIdx.z;
    const int j_min = k_abs > p.a_size - 1 ? k_abs - (p.a_size - 1) : 0, j_max = min(k_abs + 1, p.b_size);
    u128 acc = 0;
    for (int j = j_min; j < j_max; j++) {
        const long long x = reinterpret_cast<const long long *>(p.a.base + bz * p.a.batch_stride + (size_t)(k_abs - j) * p.a.limb_stride)[i];
        acc += (u128)((i128)x * (i128)p.b[j]);
    }
    char *dst = p.res.base + bz * p.res.batch_stride + (size_t)k * p.res.limb_stride;
    if (p.big_is_i128) reinterpret_cast<i128 *>(dst)[i] = (i128)acc;
    else reinterpret_cast<long long *>(dst)[i] = (long long)(unsigned long long)acc; // low 64 bits = wrapping i64 arithmetic
}
int cnv_by_const(pgb_module *m, LimbSet res, int res_size, LimbSet a, int a_size, const long long *b_dev, int b_size, uint64_t cnv_offset,
                 uint32_t batch) {
    const uint64_t n = m->n, bb = big_bytes(m);
    if (res_size == 0 || batch == 0) return PGB_OK;
    if (a_size == 0 || b_size == 0) return raw_limbs(m, true, res, res, n * bb, (uint32_t)res_size, batch);
    const int bound = a_size + b_size - 1;
    const int offset = (int)umin64(cnv_offset, (uint64_t)bound);
    const int min_size = res_size < bound + 1 - offset ? res_size : bound + 1 - offset;
    if (min_size > 0) {
        ProfScope _ps(m, PROF_ELEMENTWISE);
        CnvConstArgs p = {res, a, b_dev, (uint32_t)n, a_size, b_size, offset, m->flavour == PGB_NTT120 ? 1 : 0};
        cnv_by_const_kernel<<<dim3(((uint32_t)n + 255) / 256, min_size, batch), 256, 0, m->stream>>>(p);
        PGB_CHECK_CUDA(cudaGetLastError());
    }
    LimbSet z = res;
    z.base += (size_t)min_size * res.limb_stride;
    return raw_limbs(m, true, z, z, n * bb, (uint32_t)(res_size - min_size), batch);
}
