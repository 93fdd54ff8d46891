// tools/ubench.cu -- FP32-pipe microbenchmarks for B200 (sm_100a): what the issue slots, the FMA pipe and
// the ALU pipe can sustain for the instruction kinds the Chamfer tile kernel uses.  Standalone binary:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench tools/ubench.cu && build/ubench
// Prints cycles per warp-instruction per SM sub-partition (assuming the max SM clock) for each mix.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 sub2(u64 a, u64 b) { u64 r; asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float min3(float a, float b, float c) { float r; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float min2(float a, float b) { float r; asm volatile("min.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float add1(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float mul1(float a, float b) { float r; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

#define NCH 8
#define DEF_KERNEL(name, NINSTR, ...)                                                       \
    __global__ void __launch_bounds__(128) name(float *out, int iters, float fa, float fb) { \
        float s[NCH], q[NCH]; u64 v[NCH], w[NCH];                                             \
        for (int k = 0; k < NCH; ++k) { s[k] = fa * (threadIdx.x + k); q[k] = fb + k; v[k] = pack2(s[k], q[k]); w[k] = pack2(q[k], s[k]); } \
        const u64 pa = pack2(fa, fb), pb = pack2(fb, fa);                                    \
        for (int it = 0; it < iters; ++it) {                                                 \
            _Pragma("unroll") for (int u = 0; u < 8; ++u) {                                  \
                _Pragma("unroll") for (int k = 0; k < NCH; ++k) { __VA_ARGS__ }                     \
            }                                                                                \
        }                                                                                    \
        float acc = 0.f;                                                                     \
        for (int k = 0; k < NCH; ++k) { float lo, hi; unpack2(v[k], lo, hi); acc += lo + hi + s[k] + q[k]; unpack2(w[k], lo, hi); acc += lo + hi; } \
        if (acc == 123.456f) out[0] = acc;                                                   \
    }                                                                                        \
    static const int name##_n = NINSTR;

DEF_KERNEL(k_ffma, 1, s[k] = fma1(s[k], fa, fb);)
DEF_KERNEL(k_ffma_3reg, 1, s[k] = fma1(s[k], q[k], q[(k + 1) % NCH]);)
DEF_KERNEL(k_fadd, 1, s[k] = add1(s[k], fa);)
DEF_KERNEL(k_fmul, 1, s[k] = mul1(s[k], fa);)
DEF_KERNEL(k_ffma2, 1, v[k] = fma2(v[k], pa, pb);)
DEF_KERNEL(k_ffma2_3reg, 1, v[k] = fma2(v[k], w[k], pa);)
DEF_KERNEL(k_ffma2_sq, 1, v[k] = fma2(w[k], w[k], v[k]);)
DEF_KERNEL(k_fadd2, 1, v[k] = sub2(v[k], pa);)
DEF_KERNEL(k_fadd2_bc, 1, v[k] = sub2(pack2(s[k], s[k]), v[k]);)
DEF_KERNEL(k_fmul2, 1, v[k] = mul2(v[k], pa);)
DEF_KERNEL(k_fmul2_sq, 1, v[k] = mul2(v[k], v[k]);)
DEF_KERNEL(k_fmnmx, 1, s[k] = min2(s[k], q[k]);)
DEF_KERNEL(k_fmnmx3, 1, s[k] = min3(s[k], q[k], q[(k + 1) % NCH]);)
DEF_KERNEL(k_ffma2_min3, 2, v[k] = fma2(v[k], pa, pb); s[k] = min3(s[k], q[k], q[(k + 1) % NCH]);)
DEF_KERNEL(k_ffma2_min2, 2, v[k] = fma2(v[k], pa, pb); s[k] = min2(s[k], q[k]);)
DEF_KERNEL(k_ffma_min2, 2, s[k] = fma1(s[k], fa, fb); q[k] = min2(q[k], fa);)
DEF_KERNEL(k_ffma2x3_min3, 4, v[k] = fma2(v[k], pa, pb); w[k] = fma2(w[k], pa, pb); v[k] = fma2(v[k], pb, pa); s[k] = min3(s[k], q[k], q[(k + 1) % NCH]);)
// direct distance for (row s[k], one column pair): 3 FADD2(bcast) + FMUL2 + 2 FFMA2 (+ FMNMX3), loop-carried via s[k]
DEF_KERNEL(k_dist2, 6, { u64 px = pack2(s[k], s[k]); u64 d0 = sub2(px, pa); u64 d1 = sub2(px, pb); u64 d2 = sub2(px, w[k]);
                      u64 t = mul2(d0, d0); t = fma2(d1, d1, t); t = fma2(d2, d2, t); float lo, hi; unpack2(t, lo, hi); s[k] = lo; q[k] = hi; })
DEF_KERNEL(k_dist2_min, 7, { u64 px = pack2(s[k], s[k]); u64 d0 = sub2(px, pa); u64 d1 = sub2(px, pb); u64 d2 = sub2(px, w[k]);
                      u64 t = mul2(d0, d0); t = fma2(d1, d1, t); t = fma2(d2, d2, t); float lo, hi; unpack2(t, lo, hi); s[k] = min3(q[k], lo, hi); })
DEF_KERNEL(k_dist2_min2, 8, { u64 px = pack2(s[k], s[k]); u64 d0 = sub2(px, pa); u64 d1 = sub2(px, pb); u64 d2 = sub2(px, w[k]);
                      u64 t = mul2(d0, d0); t = fma2(d1, d1, t); t = fma2(d2, d2, t); float lo, hi; unpack2(t, lo, hi); s[k] = min3(q[k], lo, hi); q[k] = min3(q[k], hi, lo); })
// scalar version: 3 FADD + FMUL + 2 FFMA + FMNMX per pair
DEF_KERNEL(k_dist1_min, 7, { float d0 = add1(s[k], -fa); float d1 = add1(s[k], -fb); float d2 = add1(s[k], fa);
                      float t = mul1(d0, d0); t = fma1(d1, d1, t); t = fma1(d2, d2, t); s[k] = min2(q[k], t); })
// expansion form: 3 FFMA2 per column pair (+ min3)
DEF_KERNEL(k_exp2_min, 4, { u64 px = pack2(s[k], s[k]); u64 t = fma2(px, pa, w[k]); t = fma2(px, pb, t); t = fma2(px, w[(k + 1) % NCH], t);
                      float lo, hi; unpack2(t, lo, hi); s[k] = min3(q[k], lo, hi); })

__global__ void __launch_bounds__(128) k_redux(float *out, int iters, float fa, float fb) {
    unsigned s[NCH];
    for (int k = 0; k < NCH; ++k) s[k] = threadIdx.x * 77u + k;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int k = 0; k < NCH; ++k) s[k] = __reduce_min_sync(0xffffffffu, s[k] + threadIdx.x);
        }
    }
    unsigned acc = 0; for (int k = 0; k < NCH; ++k) acc += s[k];
    if (acc == 123456u) out[0] = (float)acc;
}
static const int k_redux_n = 2;   // IADD + CREDUX
__global__ void __launch_bounds__(128) k_vote(float *out, int iters, float fa, float fb) {
    unsigned s[NCH];
    for (int k = 0; k < NCH; ++k) s[k] = threadIdx.x * 77u + k;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int k = 0; k < NCH; ++k) s[k] = __ballot_sync(0xffffffffu, s[k] > threadIdx.x + it);
        }
    }
    unsigned acc = 0; for (int k = 0; k < NCH; ++k) acc += s[k];
    if (acc == 123456u) out[0] = (float)acc;
}
static const int k_vote_n = 2;    // ISETP + VOTE

template <typename F>
static void run(const char *name, F kernel, int n_instr, int sms, double mhz, float *scratch, int warps_per_smsp) {
    const int iters = 4000;
    const int grid = sms * warps_per_smsp;     // 128-thread CTAs: 4 warps = 1 per SMSP each
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kernel<<<grid, 128>>>(scratch, iters, 1.0000001f, 1e-9f);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        kernel<<<grid, 128>>>(scratch, iters, 1.0000001f, 1e-9f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    // per SMSP: warps_per_smsp warps each issuing iters*8*NCH*n_instr warp-instructions
    const double winstr = (double)warps_per_smsp * iters * 8 * NCH * n_instr;
    const double cycles = best * 1e-3 * mhz * 1e6;
    printf("%-16s warps/SMSP=%d  %8.3f ms  %6.3f cycles per warp-instr per SMSP  (%d instr/body)\n", name,
           warps_per_smsp, best, cycles / winstr, n_instr);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(e));
}

int main() {
    int dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    const double mhz = khz / 1000.0;
    printf("SMs=%d clock=%.0f MHz\n", sms, mhz);
    float *scratch; cudaMalloc(&scratch, 256);
#define RUN(k, w) run(#k, k, k##_n, sms, mhz, scratch, w)
    for (int w : {4, 8}) {
        RUN(k_ffma, w); RUN(k_ffma_3reg, w); RUN(k_fadd, w); RUN(k_fmul, w);
        RUN(k_ffma2, w); RUN(k_ffma2_3reg, w); RUN(k_ffma2_sq, w); RUN(k_fadd2, w); RUN(k_fadd2_bc, w); RUN(k_fmul2, w); RUN(k_fmul2_sq, w);
        RUN(k_fmnmx, w); RUN(k_fmnmx3, w);
        RUN(k_ffma2_min3, w); RUN(k_ffma2_min2, w); RUN(k_ffma_min2, w); RUN(k_ffma2x3_min3, w);
        RUN(k_dist2, w); RUN(k_dist2_min, w); RUN(k_dist2_min2, w); RUN(k_dist1_min, w); RUN(k_exp2_min, w);
        RUN(k_redux, w); RUN(k_vote, w);
    }
    return 0;
}
