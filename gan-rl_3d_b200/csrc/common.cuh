// common.cuh -- shared host/device helpers for librlg_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/rlg_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "librlg_b200 is written for sm_100a (B200) only"
#endif

namespace rlg {

// thread-local last-error text (api.cu)
void set_error(const char *fmt, ...);
int  fail(int code, const char *fmt, ...);
int  check_launch(const char *what);

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int sm_count();

typedef unsigned long long u64;

// ---- programmatic dependent launch (PDL): a kernel launched with launch_pdl() may start while its predecessor in the
// stream is still running; it must call pdl_wait() before touching anything the predecessor (or anything before it)
// wrote.  Every kernel calls pdl_launch_dependents() early, so its successor's CTAs can take over SM slots as they
// free up and run their independent prologue under this kernel's tail.  Outside a PDL launch both are no-ops.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
#ifdef RLG_NO_PDL      // A/B build (tools/): every kernel launched plainly
    attr[0].val.programmaticStreamSerializationAllowed = 0;
#else
    attr[0].val.programmaticStreamSerializationAllowed = 1;
#endif
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// ---- Chamfer forward
static constexpr int kGroup = 32;           // candidates per group == lanes per warp
static constexpr u64 kKeyInit = ~0ull;      // workspace state on entry and on exit of every forward

// ---- Chamfer forward, filter-and-refine path (chamfer_filter.cu).  Every array starts and ends all-ones.
struct FwdWs {
    u64 *rowkey, *colkey;          // (B,N), (B,M): filter value bits << 32 | winning candidate group
    unsigned *rowsec, *colsec;     // (B,N), (B,M): smallest filter value of any OTHER group
    unsigned *nrm;                 // (2,B): bitwise complement of the largest |p|^2 of pc1[b] / pc2[b]
    // tensor-core sweep only (plain stores, no all-ones invariant): the group holding the second-smallest group minimum
    // and the third-smallest group minimum -- an ambiguous point is refined on two groups instead of the whole cloud
    unsigned *rowsg, *colsg;       // (B,N), (B,M)
    unsigned *rowth, *colth;       // (B,N), (B,M)
};
size_t finalize2_ws_bytes(int B, int N, int M);
int launch_filter(const float *pc1, const float *pc2, int B, int N, int M, int variant, const FwdWs &w,
                  int *rows_per_lane, cudaStream_t st);
// tensor-core pair sweep with refinement, ambiguous points, means and loss fused in (chamfer_tcsweep.cu): one launch
size_t tcsweep_ws_bytes(int B, int N, int M);
int launch_tcsweep(const float *pc1, const float *pc2, int B, int N, int M, const FwdWs &w, void *fin_ws,
                   float *d1, float *d2, int32_t *i1, int32_t *i2, float *mean1, float *mean2, float *loss, float w1,
                   float w2, float *zero1, float *zero2, bool filter_only, bool force_top3, int reserve_sms, cudaStream_t st);
#ifdef RLG_EXPERIMENTS
// experiments build only (build.py --experiments): flag bits of rlg_chamfer_fwd that select measurement variants
#define RLG_X_CHAMFER_TENSOR_V1   128u     // first-generation tensor sweep (chamfer_tcfilter.cu) + separate finalize
// first-generation tensor-core pair sweep (chamfer_tcfilter.cu): the finalize is told its group layout by
// rows_per_lane <= 0 (0: the runner-up group and the third value are reported too, -1: not)
int launch_tcfilter(const float *pc1, const float *pc2, int B, int N, int M, const FwdWs &w, int *rows_per_lane, cudaStream_t st);
#endif
int launch_finalize2(const float *pc1, const float *pc2, int B, int N, int M, int rows_per_lane, const FwdWs &w,
                     void *fin_ws, float *d1, float *d2, int32_t *i1, int32_t *i2, float *mean1, float *mean2,
                     float *loss, float w1, float w2, float *zero1, float *zero2, cudaStream_t st);

// ---- one tcgen05 layer GEMM of the encoder (encoder_layers.cu), shared with the train-mode path (encoder_train.cu)
struct GemmCall {
    int B, N, K, C_out, pieces;       // X is (B, N, K), W is (C_out, K); pieces: 1 = bf16, 2 = fp16 hi + lo
    const void *x0, *x1, *w0, *w1;    // operand piece arrays (2-byte elements, row-major, K contiguous)
    const float *bias;                // fp32 (C_out), nullable with epi 2
    float out_scale;                  // multiplies the accumulator (undoes power-of-two operand scales)
    const float *dscale0, *dscale1;   // optional DEVICE scalars multiplied into out_scale
    int epi;                          // 0: ReLU -> next operand pieces y0/y1; 1: ReLU + max over points -> pooled; 2: raw fp32 -> y0
    void *y0, *y1;
    float *pooled;
};
int launch_layer_gemm(const GemmCall &g, int sms, cudaStream_t st);

// ---- packed fp32x2 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2) ------------------------------
// One instruction issues two IEEE fp32 operations, halving the issue-slot cost of the distance math.
__device__ __forceinline__ u64 pack2(float lo, float hi) {
    u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r;
}
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
    u64 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
// 3-input min (FMNMX3).  NaN operands are dropped, like fminf.
__device__ __forceinline__ float min3(float a, float b, float c) {
    float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
}
// 2-input min kept as its own instruction (ptxas otherwise re-fuses fminf(fminf(a,b),c) into FMNMX3)
__device__ __forceinline__ float min2(float a, float b) {
    float r; asm("min.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r;
}

// Squared distance in the reference's direct-mode operation order (see include/rlg_b200.h).
__device__ __forceinline__ float sqdist(float px, float py, float pz, float qx, float qy, float qz) {
    float d0 = __fsub_rn(px, qx), d1 = __fsub_rn(py, qy), d2 = __fsub_rn(pz, qz);
    float t = __fmul_rn(d0, d0);
    t = __fmaf_rn(d1, d1, t);
    t = __fmaf_rn(d2, d2, t);
    return t;
}

// The reference takes torch.min over the SQRT-ED distance matrix (utils/losses.py:29-33), and sqrtf maps up to three
// adjacent fp32 values onto one: candidates whose squared distances differ only in the last bits tie there and the lowest
// index wins.  Largest fp32 h >= m with sqrtf(h) == sqrtf(m) = s (m finite, >= 0); the exact argmin under the reference's
// rule is the lowest index j with t_j <= h.
__device__ __forceinline__ float sqrt_window_top(float m, float s) {
    float h = m;
#pragma unroll 1
    for (int step = 0; step < 3; ++step) {
        const float n = __uint_as_float(__float_as_uint(h) + 1u);
        if (!(__fsqrt_rn(n) == s)) break;
        h = n;
    }
    return h;
}

}  // namespace rlg
