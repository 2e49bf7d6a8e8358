// common.cuh -- shared host-side plumbing of libpoulpy_b200.so (module handle, error model, launch helpers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <string>

#include "../../include/poulpy_b200.h"

typedef unsigned __int128 u128;
typedef __int128 i128;

// Thread-local last error (pgb_last_error); the Rust shim turns a non-zero status into panic!.
void pgb_set_error(const char *fmt, ...);

#define PGB_CHECK_CUDA(expr)                                                                          \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) {                                                                      \
            pgb_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__,      \
                          cudaGetErrorString(_e));                                                    \
            return PGB_ERR_CUDA;                                                                      \
        }                                                                                             \
    } while (0)

#define PGB_REQUIRE(cond, ...)          \
    do {                                \
        if (!(cond)) {                  \
            pgb_set_error(__VA_ARGS__); \
            return PGB_ERR_SHAPE;       \
        }                               \
    } while (0)

#define PGB_TRY(expr)             \
    do {                          \
        int _s = (expr);          \
        if (_s != PGB_OK) return _s; \
    } while (0)

// Per-n constants of the NTT120 flavour handed to kernels by value.
struct Ntt120Consts {
    uint32_t crt_ninv[4];    // CRT_CST[k] * n^-1 mod Q[k]   (folds the 1/n of the inverse NTT into the CRT step)
    uint32_t crt_ninv_sh[4]; // Shoup companion floor(crt_ninv * 2^32 / Q[k])
};

struct pgb_module {
    uint64_t n;
    int log_n;
    int flavour;
    int device;
    cudaStream_t stream;
    bool own_stream;
    cudaStream_t aux_stream[2]; // H2D / D2H staging streams of the host front ends
    cudaEvent_t ev[8];
    uint64_t launches;
    // optional per-category kernel timing with CUDA events on the launching stream (pgb_profile_*)
    bool prof_on;
    struct ProfState *prof;
    // NTT120: twiddles (w, floor(w*2^32/q)) in block-twiddle (bit-reversed) order, [4][n] each direction
    uint2 *ntt_fwd, *ntt_inv;
    // the 15 twiddles of tree node n/16 + t and its descendants, per prime and thread, as [4][8][n/16] uint4 (slot 0: node, 1: children,
    // 2-3: grandchildren, 4-7: great-grandchildren): the last radix-16 pass of the gadget kernel loads them coalesced (null for n < 32)
    uint4 *ntt_last16_f, *ntt_last16_i;
    Ntt120Consts nc;
    uint2 tw_top_f[4][16], tw_top_i[4][16]; // host copy of block twiddles 1..15 per prime (kernel-parameter twiddles of the gadget kernel)
    // FFT64: complex twiddles in block-twiddle order, [m] each direction
    double2 *fft_fwd, *fft_inv;
    double2 *fft_last_f, *fft_last_i; // last-pass twiddles per thread, [7][m/8] (fft64.cuh: load_tw7); null for m < 8
    // lazily grown device workspace for host front ends
    void *ws;
    size_t ws_len;
    void *pinned[4];
    size_t pinned_len;
    // L2-resident carry scratch of the fused normalising kernels (16 B per coefficient per resident CTA)
    void *carry_ws;
    size_t carry_len;
    // collapsed key, key coefficient scratch and per-ciphertext guard flags of the collapsed-key fast path
    void *aux_ws;
    size_t aux_len;
    // route / tuning knobs (pgb_module_set_option); the PGB_* environment variables only seed them when the module is created, so no
    // hot-path call ever reads the environment
    int64_t opt[PGB_OPT_COUNT];
    // cached per-key forms of pinned keys (key_cache.cu; pgb_gadget_key_pin / _unpin)
    struct KeyCache *key_cache;
};
static inline bool opt_on(const pgb_module *m, int o) { return m->opt[o] != 0; }

// kernel categories of the profiler
enum { PROF_DFT_FWD = 0, PROF_DFT_INV = 1, PROF_VMP = 2, PROF_NORMALIZE = 3, PROF_ELEMENTWISE = 4, PROF_OTHER = 5, PROF_GADGET = 6, PROF_NCAT = 7 };
void prof_begin(pgb_module *m, int cat);
void prof_end(pgb_module *m);
struct ProfScope {
    pgb_module *m;
    // every kernel launch goes through here: the module's device is made current first (a kernel must be launched with its stream's device
    // current; the caller may drive several modules on several devices from one thread, or torch may have switched devices in between)
    ProfScope(pgb_module *mm, int cat) : m(mm) { cudaSetDevice(m->device); m->launches++; if (m->prof_on) prof_begin(m, cat); }
    ~ProfScope() { if (m->prof_on) prof_end(m); }
};

// A strided set of limbs ("jobs"): job j of batch item b starts at base + b*batch_stride + j*limb_stride (bytes).
struct LimbSet {
    char *base;
    uint64_t limb_stride;
    uint64_t batch_stride;
};

static inline int ilog2_u64(uint64_t x) {
    int l = 0;
    while ((1ull << l) < x) l++;
    return l;
}

static inline uint64_t umin64(uint64_t a, uint64_t b) { return a < b ? a : b; }
static inline uint64_t div_ceil64(uint64_t a, uint64_t b) { return (a + b - 1) / b; }

// byte offset of limb (col, j) inside a limb-major / column-minor container (layouts/znx_base.rs:74)
static inline uint64_t limb_off(uint64_t n, uint64_t cols, uint64_t col, uint64_t j, uint64_t scalar_bytes) {
    return n * (j * cols + col) * scalar_bytes;
}
