// tools/ubench2.cu -- which instruction mix sustains the Chamfer pair sweep on B200 (sm_100a)?
// Follow-up of tools/ubench.cu: (1) how the min instructions cost depending on where their sources live,
// (2) register-resident models of the sweep (8 row pairs per lane x 2 columns per body, operands loop-invariant)
// for several min / packing schemes.  Prints FP32-pipe cycles per point pair per lane (the floor is 4: 3 FMA + 1 ADD).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench2 tools/ubench2.cu && build/ubench2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float min3(float a, float b, float c) { float r; asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float min2(float a, float b) { float r; asm volatile("min.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float add1(float a, float b) { float r; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ int imin2(int a, int b) { int r; asm volatile("min.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }

// ---------------------------------------------------------------- (1) single-instruction probes
#define NCH 8
#define PROBE(name, NINSTR, ...)                                                                          \
    __global__ void __launch_bounds__(128) name(float *out, int iters, float fa, float fb) {              \
        float s[NCH], q[NCH]; u64 v[NCH]; int a[NCH], b[NCH], c[NCH];                                      \
        for (int k = 0; k < NCH; ++k) { s[k] = fa * (threadIdx.x + k); q[k] = fb + k; v[k] = pack2(s[k] + 1.f, q[k] + 2.f); \
                                        a[k] = threadIdx.x * 3 + k; b[k] = threadIdx.x ^ (k * 77); c[k] = k - (int)threadIdx.x; } \
        const u64 pa = pack2(fa, fb), pb = pack2(fb, fa);                                                 \
        for (int it = 0; it < iters; ++it) {                                                              \
            _Pragma("unroll") for (int u = 0; u < 8; ++u) {                                               \
                _Pragma("unroll") for (int k = 0; k < NCH; ++k) { __VA_ARGS__ }                           \
            }                                                                                             \
        }                                                                                                 \
        float acc = 0.f;                                                                                  \
        for (int k = 0; k < NCH; ++k) { float lo, hi; unpack2(v[k], lo, hi); acc += lo + hi + s[k] + q[k] + a[k] + b[k] + c[k]; } \
        if (acc == 123.456f) out[0] = acc + (float)(pa + pb);                                             \
    }                                                                                                     \
    static const int name##_n = NINSTR;

PROBE(p_min3_pair, 1, { float lo, hi; unpack2(v[k], lo, hi); s[k] = min3(s[k], lo, hi); })
PROBE(p_min3_3reg, 1, s[k] = min3(s[k], q[k], q[(k + 1) % NCH]);)
PROBE(p_min3_2reg, 1, s[k] = min3(s[k], q[k], q[k]);)
PROBE(p_min3_const, 1, s[k] = min3(s[k], q[k], fa);)
PROBE(p_min2, 1, s[k] = min2(s[k], q[k]);)
PROBE(p_imin2, 1, a[k] = imin2(a[k], b[k]);)
PROBE(p_imin3, 1, a[k] = __vimin3_s32(a[k], b[k], c[k]);)
PROBE(p_imin3_pair, 1, { float lo, hi; unpack2(v[k], lo, hi); a[k] = __vimin3_s32(a[k], __float_as_int(lo), __float_as_int(hi)); })
PROBE(p_viaddmin, 1, a[k] = __viaddmin_s32(b[k], c[k], a[k]);)
PROBE(p_ffma2_min3pair, 2, { v[k] = fma2(v[k], pa, pb); float lo, hi; unpack2(v[k], lo, hi); s[k] = min3(s[k], lo, hi); })
PROBE(p_ffma2_2min2, 3, { v[k] = fma2(v[k], pa, pb); float lo, hi; unpack2(v[k], lo, hi); s[k] = min2(s[k], lo); q[k] = min2(q[k], hi); })
PROBE(p_ffma2_imin3pair, 2, { v[k] = fma2(v[k], pa, pb); float lo, hi; unpack2(v[k], lo, hi); a[k] = __vimin3_s32(a[k], __float_as_int(lo), __float_as_int(hi)); })
PROBE(p_ffma_min2, 2, { s[k] = fma1(s[k], fa, fb); q[k] = min2(q[k], s[k]); })
PROBE(p_2ffma_min3, 3, { s[k] = fma1(s[k], fa, fb); float t = fma1(s[k], fb, fa); q[k] = min3(q[k], s[k], t); })

// ---------------------------------------------------------------- (2) register-resident sweep models
// 8 row pairs per lane (16 rows), body = one column pair (A,B) -> 32 point pairs per lane per body.
// MODE 0: packed FFMA2 x3 + FADD2; rows min3 over the two columns; columns min3 over the row pair   (the kernel today)
// MODE 1: packed; rows 2 x min2; columns min3 over the row pair
// MODE 2: packed; all min2
// MODE 3: packed; FMA + ADD only, one min3 per body to keep the results alive (upper bound)
// MODE 4: scalar FFMA x3 + FADD per pair; rows min3 over the two columns; columns min3 over two rows
// MODE 5: scalar; all min2
// MODE 6: packed; integer mins on the bit patterns (values are positive here): rows VIMNMX3, columns VIMNMX3
// MODE 7: packed, columns packed instead of rows: x duplicated, rows min3 on the pair, columns 2 x min2
template <int MODE>
__global__ void __launch_bounds__(128) sweep_model(float *out, int iters, float fa, float fb) {
    constexpr int P = 8;
    u64 xp0[P], xp1[P], xp2[P], nxp[P];
    float rowmin[2 * P];
    for (int p = 0; p < P; ++p) {
        const float a = fa * (threadIdx.x + p), b = fb * (p + 1);
        xp0[p] = pack2(a, b); xp1[p] = pack2(b + 1.f, a - 1.f); xp2[p] = pack2(a * 0.5f, b * 0.25f); nxp[p] = pack2(a * a, b * b);
        rowmin[2 * p] = 1e30f; rowmin[2 * p + 1] = 1e30f;
    }
    // loop-invariant column operands (the real kernel reloads them from shared memory every body)
    u64 A0 = pack2(fa, fa), A1 = pack2(fb, fb), A2 = pack2(fa + fb, fa + fb), A3 = pack2(fa * fb, fa * fb);
    u64 B0 = pack2(fb, fb), B1 = pack2(fa, fa), B2 = pack2(fa - fb, fa - fb), B3 = pack2(fa * fa, fa * fa);
    float cA = 1e30f, cB = 1e30f;
    int icA = 0x7fffffff, icB = 0x7fffffff;
    int irow[2 * P];
    for (int r = 0; r < 2 * P; ++r) irow[r] = 0x7fffffff;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int p = 0; p < P; ++p) {
            if (MODE <= 3 || MODE == 6 || MODE == 7) {
                const u64 gA = fma2(xp0[p], A0, fma2(xp1[p], A1, fma2(xp2[p], A2, A3)));
                const u64 gB = fma2(xp0[p], B0, fma2(xp1[p], B1, fma2(xp2[p], B2, B3)));
                const u64 tA = add2(gA, nxp[p]);
                const u64 tB = add2(gB, nxp[p]);
                float gAl, gAh, gBl, gBh, tAl, tAh, tBl, tBh;
                unpack2(gA, gAl, gAh); unpack2(gB, gBl, gBh); unpack2(tA, tAl, tAh); unpack2(tB, tBl, tBh);
                if (MODE == 0) {
                    rowmin[2 * p] = min3(rowmin[2 * p], gAl, gBl); rowmin[2 * p + 1] = min3(rowmin[2 * p + 1], gAh, gBh);
                    cA = min3(cA, tAl, tAh); cB = min3(cB, tBl, tBh);
                } else if (MODE == 1) {
                    rowmin[2 * p] = min2(min2(rowmin[2 * p], gAl), gBl); rowmin[2 * p + 1] = min2(min2(rowmin[2 * p + 1], gAh), gBh);
                    cA = min3(cA, tAl, tAh); cB = min3(cB, tBl, tBh);
                } else if (MODE == 2) {
                    rowmin[2 * p] = min2(min2(rowmin[2 * p], gAl), gBl); rowmin[2 * p + 1] = min2(min2(rowmin[2 * p + 1], gAh), gBh);
                    cA = min2(min2(cA, tAl), tAh); cB = min2(min2(cB, tBl), tBh);
                } else if (MODE == 3) {
                    if (p == 0) { cA = min3(cA, tAl, tAh); cB = min3(cB, tBl, tBh); rowmin[0] = min3(rowmin[0], gAl, gBh); }
                    else { xp0[p] = add2(tA, tB); }          // keeps the chain alive at FMA-pipe cost only... one extra ADD
                } else if (MODE == 6) {
                    irow[2 * p] = __vimin3_s32(irow[2 * p], __float_as_int(gAl), __float_as_int(gBl));
                    irow[2 * p + 1] = __vimin3_s32(irow[2 * p + 1], __float_as_int(gAh), __float_as_int(gBh));
                    icA = __vimin3_s32(icA, __float_as_int(tAl), __float_as_int(tAh));
                    icB = __vimin3_s32(icB, __float_as_int(tBl), __float_as_int(tBh));
                } else {   // MODE 7: pretend the pair is two columns of one row: rows take the pair, columns split
                    rowmin[2 * p] = min3(rowmin[2 * p], gAl, gAh); rowmin[2 * p + 1] = min3(rowmin[2 * p + 1], gBl, gBh);
                    cA = min2(cA, tAl); cB = min2(cB, tAh); cA = min2(cA, tBl); cB = min2(cB, tBh);
                }
            } else {
                float x0l, x0h, x1l, x1h, x2l, x2h, nl, nh, a0, a1, a2, a3, b0, b1, b2, b3, d;
                unpack2(xp0[p], x0l, x0h); unpack2(xp1[p], x1l, x1h); unpack2(xp2[p], x2l, x2h); unpack2(nxp[p], nl, nh);
                unpack2(A0, a0, d); unpack2(A1, a1, d); unpack2(A2, a2, d); unpack2(A3, a3, d);
                unpack2(B0, b0, d); unpack2(B1, b1, d); unpack2(B2, b2, d); unpack2(B3, b3, d);
                const float gAl = fma1(x0l, a0, fma1(x1l, a1, fma1(x2l, a2, a3)));
                const float gAh = fma1(x0h, a0, fma1(x1h, a1, fma1(x2h, a2, a3)));
                const float gBl = fma1(x0l, b0, fma1(x1l, b1, fma1(x2l, b2, b3)));
                const float gBh = fma1(x0h, b0, fma1(x1h, b1, fma1(x2h, b2, b3)));
                const float tAl = add1(gAl, nl), tAh = add1(gAh, nh), tBl = add1(gBl, nl), tBh = add1(gBh, nh);
                if (MODE == 4) {
                    rowmin[2 * p] = min3(rowmin[2 * p], gAl, gBl); rowmin[2 * p + 1] = min3(rowmin[2 * p + 1], gAh, gBh);
                    cA = min3(cA, tAl, tAh); cB = min3(cB, tBl, tBh);
                } else {
                    rowmin[2 * p] = min2(min2(rowmin[2 * p], gAl), gBl); rowmin[2 * p + 1] = min2(min2(rowmin[2 * p + 1], gAh), gBh);
                    cA = min2(min2(cA, tAl), tAh); cB = min2(min2(cB, tBl), tBh);
                }
            }
        }
        // perturb the column operands a little so nothing is loop-invariant for the compiler (2 instructions per body)
        A3 = add2(A3, B0);
        B3 = add2(B3, A0);
    }
    float acc = cA + cB + (float)icA + (float)icB;
    for (int r = 0; r < 2 * P; ++r) acc += rowmin[r] + (float)irow[r];
    for (int p = 0; p < P; ++p) { float lo, hi; unpack2(xp0[p], lo, hi); acc += lo + hi; }
    if (acc == 123.456f) out[0] = acc;
}

static double g_mhz;
static int g_sms;
static float *g_scratch;

template <typename F>
static float time_kernel(F kernel, int grid, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    kernel<<<grid, 128>>>(g_scratch, iters, 1.0000001f, 1e-9f);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        kernel<<<grid, 128>>>(g_scratch, iters, 1.0000001f, 1e-9f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(e));
    return best;
}

template <typename F>
static void run_probe(const char *name, F kernel, int n_instr, int w) {
    const int iters = 4000;
    const float ms = time_kernel(kernel, g_sms * w, iters);
    const double winstr = (double)w * iters * 8 * NCH * n_instr;
    printf("%-20s warps/SMSP=%d  %6.3f cycles per warp-instr per SMSP  (%d instr/body, %.2f cycles/body)\n", name, w,
           ms * 1e-3 * g_mhz * 1e6 / winstr, n_instr, ms * 1e-3 * g_mhz * 1e6 / winstr * n_instr);
}

template <int MODE>
static void run_model(const char *what, int w) {
    const int iters = 20000;
    const float ms = time_kernel(sweep_model<MODE>, g_sms * w, iters);
    // per SMSP: w warps x iters bodies x 32 pairs per lane
    const double pairs = (double)w * iters * 32.0;
    const double cyc = ms * 1e-3 * g_mhz * 1e6 / pairs;
    printf("model %d %-58s warps/SMSP=%d  %5.2f cycles/pair  (%.0f%% of the FP32 pipe)\n", MODE, what, w, cyc, 400.0 / cyc);
}

int main() {
    int dev = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    g_mhz = khz / 1000.0;
    printf("SMs=%d clock=%.0f MHz\n", g_sms, g_mhz);
    cudaMalloc(&g_scratch, 256);
#define RUN(k, w) run_probe(#k, k, k##_n, w)
    for (int w : {2, 4}) {
        RUN(p_min3_pair, w); RUN(p_min3_3reg, w); RUN(p_min3_2reg, w); RUN(p_min3_const, w); RUN(p_min2, w);
        RUN(p_imin2, w); RUN(p_imin3, w); RUN(p_imin3_pair, w); RUN(p_viaddmin, w);
        RUN(p_ffma2_min3pair, w); RUN(p_ffma2_2min2, w); RUN(p_ffma2_imin3pair, w); RUN(p_ffma_min2, w); RUN(p_2ffma_min3, w);
    }
    for (int w : {1, 2, 3, 4}) {
        run_model<3>("packed, FMA+ADD only (bound)", w);
        run_model<0>("packed, rows min3(2 cols), cols min3(row pair)  [today]", w);
        run_model<1>("packed, rows 2 x min2, cols min3(row pair)", w);
        run_model<2>("packed, all min2", w);
        run_model<6>("packed, integer min3 for both", w);
        run_model<7>("packed, rows min3(pair), cols 2 x min2", w);
        run_model<4>("scalar FFMA, min3 for both", w);
        run_model<5>("scalar FFMA, all min2", w);
    }
    return 0;
}
