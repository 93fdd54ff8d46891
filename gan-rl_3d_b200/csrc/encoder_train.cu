// encoder_train.cu -- the PointNet encoder trunk WITH its BatchNorm layers, forward and backward: what the autoencoder
// training step runs (train_rl_gan_net.py:220-249 -> models/autoencoder.py:32-47,65-71 with model.train()).
//
// Per block  z = Conv1d_k1(a_prev) ; y = BatchNorm1d(z) ; a = ReLU(y) ; the last block is followed by max over the points.
//
//   forward   layer 0 (3 -> C0)  CUDA cores, fp32
//             layer l >= 1       tcgen05 GEMM of encoder_layers.cu (fp16 hi+lo operands = fp32 grade, raw fp32 epilogue)
//             BatchNorm          one pass of shifted sums per channel (float64 accumulators) -> mean / biased variance ->
//                                running-stat update -> one pass that normalises, applies ReLU and writes the next GEMM's
//                                operand pieces (or, for the last layer, reduces max + arg-max over the points)
//   backward  BatchNorm          dz = gamma*invstd * (dy - mean(dy) - zhat * mean(dy*zhat)),  dy = da * [y > 0]
//                                (one pass for the two sums, one pass that writes dz as scaled operand pieces)
//             input gradient     da_prev = dz . W          the same GEMM kernel with the transposed weight image
//             weight gradient    dW = dz^T . a_prev        encoder_wgrad_kernel below: both operands are read straight from
//                                the point-major piece arrays as MN-MAJOR tcgen05 operands (TMA boxes of 64 channels x 64
//                                points), split over the points across the SMs, partial tiles reduced in a fixed order
//             layer 0            dW0 = dz0^T . x on the CUDA cores (K = 3)
//
// Operand scaling: fp16 pieces need |v| <= 65504 and lose their low part below 2^-14, so weights and gradients are scaled
// by a power of two found ON THE DEVICE (no host synchronisation: training changes the weights every step) -- weights by
// their max, dz by a bound computed from the sums of its own formula -- and the GEMM epilogues undo the scale exactly.
// The tensor core's fp32 accumulation truncates; long sums are cut into short chains (the forward GEMM's four
// accumulators; the weight-gradient kernel drains its accumulators into registers every 192 points and adds there in
// round-to-nearest).
#include "common.cuh"
#include "tcgen05.cuh"

namespace rlg {

static constexpr int kTMaxLayers = 8;
static constexpr int kTMaxC = 256;
enum { EPI_RAW_ = 2 };

// ---- buffer layouts ---------------------------------------------------------------------------------------------------
struct SavedLayout {
    size_t z[kTMaxLayers];        // fp32 [P x C_l]   pre-BatchNorm outputs
    size_t ahi[kTMaxLayers];      // fp16 [P x C_l]   activation pieces (l < L-1)
    size_t alo[kTMaxLayers];
    size_t bnp[kTMaxLayers];      // fp32 [4 x C_l]   scale = gamma*invstd, shift = beta - mean*scale, mean*invstd, invstd
    size_t zmax[kTMaxLayers];     // u32  [C_l]       bit pattern of max |zhat| per channel
    size_t keys;                  // u64  [B x C_last] (value bits << 32 | ~point index) of the max-pool
    size_t wscale;                // f32  [L][2]      weight scale and its inverse (device scalars)
    size_t wp[kTMaxLayers][4];    // fp16 pieces of W_l: hi, lo (c_out x c_in) and of W_l^T: hi, lo (c_in x c_out); l >= 1
    size_t total;
};
static SavedLayout saved_layout(long long P, int B, const rlg_bn_layer *layers, int L) {
    SavedLayout s;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
    for (int l = 0; l < L; ++l) {
        const size_t C = (size_t)layers[l].c_out;
        s.z[l] = take((size_t)P * C * 4);
        s.ahi[l] = l < L - 1 ? take((size_t)P * C * 2) : 0;
        s.alo[l] = l < L - 1 ? take((size_t)P * C * 2) : 0;
        s.bnp[l] = take(4 * C * 4);
        s.zmax[l] = take(C * 4);
    }
    s.keys = take((size_t)B * layers[L - 1].c_out * 8);
    s.wscale = take((size_t)L * 2 * 4);
    for (int l = 0; l < L; ++l)
        for (int k = 0; k < 4; ++k) s.wp[l][k] = l >= 1 ? take((size_t)layers[l].c_out * layers[l].c_in * 2) : 0;
    s.total = off;
    return s;
}

struct WsLayout {
    // zero-filled at the start of every call: [0, zero_bytes)
    size_t stat;                  // f64 [L][2][256]   forward: shifted sum, shifted sum of squares; backward: sum dy, sum dy*zhat
    size_t mxdy;                  // u32 [L][256]      backward: bit pattern of max |dy| per channel
    size_t w0acc;                 // f64 [256 x 3]     layer-0 weight gradient accumulators
    size_t wamax;                 // u32 [L]           forward: bit pattern of max |w| per layer
    size_t zero_bytes;
    size_t dscale;                // f32 [L][2]        dz scale and its inverse
    size_t m12;                   // f32 [L][2][256]   mean(dy), mean(dy*zhat)
    size_t da;                    // fp32 [P x Cmax]   gradient w.r.t. a layer's activations
    size_t dzhi, dzlo;            // fp16 [P x Cmax]
    size_t partial;               // fp32 [sms][128 x 128] weight-gradient partial tiles
    size_t total;
};
static WsLayout ws_layout(long long P, const rlg_bn_layer *layers, int L, int sms) {
    WsLayout w;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
    w.stat = take((size_t)L * 2 * kTMaxC * 8);
    w.mxdy = take((size_t)L * kTMaxC * 4);
    w.w0acc = take((size_t)kTMaxC * 3 * 8);
    w.wamax = take((size_t)kTMaxLayers * 4);
    w.zero_bytes = off;
    w.dscale = take((size_t)L * 2 * 4);
    w.m12 = take((size_t)L * 2 * kTMaxC * 4);
    int cmax = 0;
    for (int l = 0; l < L; ++l) cmax = layers[l].c_out > cmax ? layers[l].c_out : cmax;
    w.da = take((size_t)P * cmax * 4);
    w.dzhi = take((size_t)P * cmax * 2);
    w.dzlo = take((size_t)P * cmax * 2);
    w.partial = take((size_t)(sms > 0 ? sms : 1) * 128 * 128 * 4);
    w.total = off;
    return w;
}

// ---- per-channel reductions over the points: 256 threads, a thread owns 4 consecutive channels of every RPI-th row ------
struct ColGeom {
    int tpr, rpi, tx, ty;
    bool active;
    __device__ __forceinline__ ColGeom(int C) {
        tpr = C >> 2;
        rpi = 256 / tpr;
        tx = threadIdx.x % tpr;
        ty = threadIdx.x / tpr;
        active = ty < rpi;
    }
};
// red: shared [Q][rpi][C]; after the call threads c < C hold the column totals of quantity q in out[q] (as double)
template <int Q>
__device__ __forceinline__ void col_reduce_sum(const float (&v)[Q][4], const ColGeom &g, int C, float *red, double (&out)[Q]) {
    if (g.active) {
#pragma unroll
        for (int q = 0; q < Q; ++q)
#pragma unroll
            for (int k = 0; k < 4; ++k) red[(q * g.rpi + g.ty) * C + 4 * g.tx + k] = v[q][k];
    }
    __syncthreads();
    if ((int)threadIdx.x < C) {
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            double s = 0.0;
            for (int r = 0; r < g.rpi; ++r) s += (double)red[(q * g.rpi + r) * C + threadIdx.x];
            out[q] = s;
        }
    }
    __syncthreads();
}
__device__ __forceinline__ void col_reduce_max(const float (&v)[4], const ColGeom &g, int C, float *red, float &out) {
    if (g.active) {
#pragma unroll
        for (int k = 0; k < 4; ++k) red[g.ty * C + 4 * g.tx + k] = v[k];
    }
    __syncthreads();
    if ((int)threadIdx.x < C) {
        float m = 0.0f;
        for (int r = 0; r < g.rpi; ++r) m = fmaxf(m, red[r * C + threadIdx.x]);
        out = m;
    }
    __syncthreads();
}

// ---- forward kernels --------------------------------------------------------------------------------------------------
// layer 0: z0[p][c] = b[c] + x[p] . w[c]   (one thread per point and 8 channels)
__global__ void __launch_bounds__(256) train_layer0_kernel(const float *__restrict__ x, long long P, int C, const float *__restrict__ w,
                                                          const float *__restrict__ bias, float *__restrict__ z) {
    const int chunks = C / 8;
    const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
    if (e >= P * chunks) return;
    const long long p = e / chunks;
    const int c0 = (int)(e - p * chunks) * 8;
    const float px = __ldg(x + 3 * p), py = __ldg(x + 3 * p + 1), pz = __ldg(x + 3 * p + 2);
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float *wr = w + 3 * (c0 + k);
        v[k] = fmaf(pz, __ldg(wr + 2), fmaf(py, __ldg(wr + 1), fmaf(px, __ldg(wr), bias ? __ldg(bias + c0 + k) : 0.0f)));
    }
    float4 *dst = reinterpret_cast<float4 *>(z + p * C + c0);
    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
    dst[1] = make_float4(v[4], v[5], v[6], v[7]);
}

// Row loops of the per-channel kernels: a thread visits the rows  first, first + stride, ...  four at a time, so four
// independent loads are in flight per thread (one per iteration left these kernels latency-bound at a sixth of HBM speed)
template <typename Load, typename Use>
__device__ __forceinline__ void rows4(long long first, long long stride, long long P, Load load, Use use) {
    for (long long row0 = first; row0 < P; row0 += 4 * stride) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long row = row0 + (long long)u * stride;
            load(u, row, row < P);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const long long row = row0 + (long long)u * stride;
            if (row < P) use(u, row);
        }
    }
}

// shifted sums per channel: acc[c] += sum_p (z - K_c), acc[256 + c] += sum_p (z - K_c)^2 with K_c = z[0][c] (any sample of
// the channel keeps the subtraction in the variance formula harmless)
__global__ void __launch_bounds__(256) bn_stats_kernel(const float *__restrict__ z, long long P, int C, double *__restrict__ acc) {
    __shared__ float red[2 * 1024];
    const ColGeom g(C);
    float v[2][4] = {};
    if (g.active) {
        const float4 k4 = __ldg(reinterpret_cast<const float4 *>(z) + g.tx);
        float4 t[4];
        rows4((long long)blockIdx.x * g.rpi + g.ty, (long long)gridDim.x * g.rpi, P,
                  [&](int u, long long row, bool ok) { t[u] = ok ? __ldg(reinterpret_cast<const float4 *>(z + row * C) + g.tx) : k4; },
                  [&](int u, long long row) {
                      const float d0 = t[u].x - k4.x, d1 = t[u].y - k4.y, d2 = t[u].z - k4.z, d3 = t[u].w - k4.w;
                      v[0][0] += d0; v[0][1] += d1; v[0][2] += d2; v[0][3] += d3;
                      v[1][0] = fmaf(d0, d0, v[1][0]); v[1][1] = fmaf(d1, d1, v[1][1]);
                      v[1][2] = fmaf(d2, d2, v[1][2]); v[1][3] = fmaf(d3, d3, v[1][3]);
                  });
    }
    double out[2];
    col_reduce_sum<2>(v, g, C, red, out);
    if ((int)threadIdx.x < C) {
        atomicAdd(acc + threadIdx.x, out[0]);
        atomicAdd(acc + kTMaxC + threadIdx.x, out[1]);
    }
}

// What BatchNorm makes of a layer's statistics, computed by EVERY CTA of the consuming kernel into shared memory (C <= 256
// channels: cheaper than a launch of its own); `writer` (one CTA) also stores it for the backward and updates the running
// statistics in train mode.  s_bnp[0..C) = gamma*invstd, [C..2C) = beta - mean*gamma*invstd, [2C..3C) = mean*invstd, [3C..4C) = invstd.
struct BnArgs {
    const double *acc;            // shifted sums of bn_stats_kernel (train mode)
    const float *gamma, *beta;
    float *rmean, *rvar;
    float eps, momentum;
    int batch_stats;
    float *bnp;                   // out (saved): [4][C]
};
__device__ __forceinline__ void bn_prologue(const float *__restrict__ z, long long P, int C, const BnArgs &a, bool writer, float *s_bnp) {
    const int c = threadIdx.x;
    if (c < C) {
        double mean, var;
        if (a.batch_stats) {
            const double n = (double)P, s1 = a.acc[c], s2 = a.acc[kTMaxC + c];
            const double dm = s1 / n;
            mean = (double)z[c] + dm;
            var = s2 / n - dm * dm;
            if (var < 0.0) var = 0.0;
            if (writer) {
                if (a.rmean) a.rmean[c] = (float)((1.0 - (double)a.momentum) * (double)a.rmean[c] + (double)a.momentum * mean);
                if (a.rvar) a.rvar[c] = (float)((1.0 - (double)a.momentum) * (double)a.rvar[c] + (double)a.momentum * var * (n / (n - 1.0)));
            }
        } else {
            mean = (double)a.rmean[c];
            var = (double)a.rvar[c];
        }
        const double invstd = 1.0 / sqrt(var + (double)a.eps);
        const double ga = a.gamma ? (double)a.gamma[c] : 1.0, be = a.beta ? (double)a.beta[c] : 0.0;
        const float f0 = (float)(ga * invstd), f1 = (float)(be - mean * ga * invstd), f2 = (float)(mean * invstd), f3 = (float)invstd;
        s_bnp[c] = f0; s_bnp[C + c] = f1; s_bnp[2 * C + c] = f2; s_bnp[3 * C + c] = f3;
        if (writer) { a.bnp[c] = f0; a.bnp[C + c] = f1; a.bnp[2 * C + c] = f2; a.bnp[3 * C + c] = f3; }
    }
    __syncthreads();
}

// a = ReLU(z * scale + shift) -> fp16 hi+lo operand rows of the next GEMM; max |zhat| per channel for the backward's scale
__global__ void __launch_bounds__(256) bn_act_kernel(const float *__restrict__ z, long long P, int C, BnArgs bn,
                                                    unsigned short *__restrict__ ahi, unsigned short *__restrict__ alo,
                                                    unsigned *__restrict__ zmax) {
    __shared__ float red[1024];
    __shared__ __align__(16) float s_bnp[4 * kTMaxC];
    bn_prologue(z, P, C, bn, blockIdx.x == 0, s_bnp);
    const ColGeom g(C);
    float zm[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (g.active) {
        const float4 sc = reinterpret_cast<const float4 *>(s_bnp)[g.tx], sh = reinterpret_cast<const float4 *>(s_bnp + C)[g.tx];
        const float4 mi = reinterpret_cast<const float4 *>(s_bnp + 2 * C)[g.tx], is = reinterpret_cast<const float4 *>(s_bnp + 3 * C)[g.tx];
        float4 t[4];
        rows4((long long)blockIdx.x * g.rpi + g.ty, (long long)gridDim.x * g.rpi, P,
                  [&](int u, long long row, bool ok) { t[u] = ok ? __ldg(reinterpret_cast<const float4 *>(z + row * C) + g.tx) : make_float4(0.f, 0.f, 0.f, 0.f); },
                  [&](int u, long long row) {
                      const float4 tt = t[u];
                      const float a0 = fmaxf(fmaf(tt.x, sc.x, sh.x), 0.0f), a1 = fmaxf(fmaf(tt.y, sc.y, sh.y), 0.0f);
                      const float a2 = fmaxf(fmaf(tt.z, sc.z, sh.z), 0.0f), a3 = fmaxf(fmaf(tt.w, sc.w, sh.w), 0.0f);
                      zm[0] = fmaxf(zm[0], fabsf(fmaf(tt.x, is.x, -mi.x))); zm[1] = fmaxf(zm[1], fabsf(fmaf(tt.y, is.y, -mi.y)));
                      zm[2] = fmaxf(zm[2], fabsf(fmaf(tt.z, is.z, -mi.z))); zm[3] = fmaxf(zm[3], fabsf(fmaf(tt.w, is.w, -mi.w)));
                      uint32_t h0, l0, h1, l1;
                      split_f16x2(a0, a1, h0, l0);
                      split_f16x2(a2, a3, h1, l1);
                      const size_t o = ((size_t)row * C + 4 * g.tx);
                      *reinterpret_cast<uint2 *>(ahi + o) = make_uint2(h0, h1);
                      *reinterpret_cast<uint2 *>(alo + o) = make_uint2(l0, l1);
                  });
    }
    float m;
    col_reduce_max(zm, g, C, red, m);
    if ((int)threadIdx.x < C && m > 0.0f) atomicMax(zmax + threadIdx.x, __float_as_uint(m));
}

// last layer: a = ReLU(BatchNorm(z)), max and arg-max over the points of a cloud.  grid (chunks of points, B)
__global__ void __launch_bounds__(256) bn_pool_kernel(const float *__restrict__ z, long long P, int N, int C, int rows_per_cta, BnArgs bn,
                                                     u64 *__restrict__ keys, unsigned *__restrict__ zmax) {
    __shared__ u64 kred[1024];
    __shared__ float red[1024];
    __shared__ __align__(16) float s_bnp[4 * kTMaxC];
    bn_prologue(z, P, C, bn, blockIdx.x == 0 && blockIdx.y == 0, s_bnp);
    const ColGeom g(C);
    const int b = blockIdx.y;
    const int n_begin = blockIdx.x * rows_per_cta, n_end = min(N, n_begin + rows_per_cta);
    float zm[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    u64 best[4] = {0, 0, 0, 0};
    if (g.active) {
        const float4 sc = reinterpret_cast<const float4 *>(s_bnp)[g.tx], sh = reinterpret_cast<const float4 *>(s_bnp + C)[g.tx];
        const float4 mi = reinterpret_cast<const float4 *>(s_bnp + 2 * C)[g.tx], is = reinterpret_cast<const float4 *>(s_bnp + 3 * C)[g.tx];
        const float *zb = z + (size_t)b * N * C;
        float4 t[4];
        rows4((long long)n_begin + g.ty, (long long)g.rpi, (long long)n_end,
                  [&](int u, long long row, bool ok) { t[u] = ok ? __ldg(reinterpret_cast<const float4 *>(zb + row * C) + g.tx) : make_float4(0.f, 0.f, 0.f, 0.f); },
                  [&](int u, long long row) {
                      const float4 tt = t[u];
                      const float a[4] = {fmaxf(fmaf(tt.x, sc.x, sh.x), 0.0f), fmaxf(fmaf(tt.y, sc.y, sh.y), 0.0f),
                                          fmaxf(fmaf(tt.z, sc.z, sh.z), 0.0f), fmaxf(fmaf(tt.w, sc.w, sh.w), 0.0f)};
                      zm[0] = fmaxf(zm[0], fabsf(fmaf(tt.x, is.x, -mi.x))); zm[1] = fmaxf(zm[1], fabsf(fmaf(tt.y, is.y, -mi.y)));
                      zm[2] = fmaxf(zm[2], fabsf(fmaf(tt.z, is.z, -mi.z))); zm[3] = fmaxf(zm[3], fabsf(fmaf(tt.w, is.w, -mi.w)));
                      const u64 low = (u64)(0xFFFFFFFFu - (unsigned)row);
                      for (int k = 0; k < 4; ++k) {
                          const u64 key = ((u64)__float_as_uint(a[k]) << 32) | low;    // a >= 0: the bit pattern orders like the value
                          best[k] = key > best[k] ? key : best[k];
                      }
                  });
#pragma unroll
        for (int k = 0; k < 4; ++k) kred[g.ty * C + 4 * g.tx + k] = best[k];
    }
    __syncthreads();
    if ((int)threadIdx.x < C) {
        u64 m = 0;
        for (int r = 0; r < g.rpi; ++r) m = kred[r * C + threadIdx.x] > m ? kred[r * C + threadIdx.x] : m;
        if (n_begin < n_end) atomicMax(keys + (size_t)b * C + threadIdx.x, m);
    }
    __syncthreads();
    float m;
    col_reduce_max(zm, g, C, red, m);
    if ((int)threadIdx.x < C && m > 0.0f) atomicMax(zmax + threadIdx.x, __float_as_uint(m));
}

__global__ void __launch_bounds__(256) pool_decode_kernel(const u64 *__restrict__ keys, int n, float *__restrict__ pooled) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) pooled[i] = __uint_as_float((unsigned)(keys[i] >> 32));
}

// weights of the layers >= 1 -> fp16 hi+lo pieces scaled by a power of two (max |w| * scale in [2^13, 2^14)), in both
// orientations: (c_out x c_in) for the forward GEMM and (c_in x c_out) for the input-gradient GEMM.  Two launches over
// (chunk, layer): the per-layer max, then the split; the backward reuses the forward's pieces (they live in `saved`).
static constexpr int kPackChunks = 16;
struct PackArgs {
    const float *w[kTMaxLayers];
    unsigned short *p[kTMaxLayers][4];
    int c_out[kTMaxLayers], c_in[kTMaxLayers];
    unsigned *wamax;
    float *wscale;
};
__global__ void __launch_bounds__(256) train_wamax_kernel(PackArgs a) {
    __shared__ float red[8];
    const int l = blockIdx.y + 1;
    const int n = a.c_out[l] * a.c_in[l];
    const float *w = a.w[l];
    float m = 0.0f;
    for (int e = blockIdx.x * 256 + threadIdx.x; e < n; e += kPackChunks * 256) m = fmaxf(m, fabsf(__ldg(w + e)));
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        float mm = 0.0f;
        for (int k = 0; k < 8; ++k) mm = fmaxf(mm, red[k]);
        if (mm > 0.0f) atomicMax(a.wamax + l, __float_as_uint(mm));
    }
}
__global__ void __launch_bounds__(256) train_pack_kernel(PackArgs a) {
    const int l = blockIdx.y + 1;
    const int co = a.c_out[l], ci = a.c_in[l], n = co * ci;
    const float *w = a.w[l];
    const float mm = __uint_as_float(a.wamax[l]);
    int ex = 0;
    if (mm > 0.0f && mm < 3.0e38f) ex = 14 - (ilogbf(mm) + 1);           // mm * 2^ex in [2^13, 2^14)
    ex = max(-100, min(100, ex));
    const float sc = ldexpf(1.0f, ex);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.wscale[2 * l] = sc;
        a.wscale[2 * l + 1] = ldexpf(1.0f, -ex);
    }
    for (int e = blockIdx.x * 256 + threadIdx.x; e < n; e += kPackChunks * 256) {
        const float v = __ldg(w + e) * sc;
        const __half h = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(h));
        const unsigned short hb = *reinterpret_cast<const unsigned short *>(&h), lb = *reinterpret_cast<const unsigned short *>(&lo);
        const int r = e / ci, c = e - r * ci;
        a.p[l][0][e] = hb;
        a.p[l][1][e] = lb;
        a.p[l][2][c * co + r] = hb;
        a.p[l][3][c * co + r] = lb;
    }
}

// ---- backward kernels -------------------------------------------------------------------------------------------------
// last layer: dy is non-zero only at the arg-max point of each (cloud, channel) with a positive maximum
__global__ void __launch_bounds__(256) pool_bwd_stats_kernel(const float *__restrict__ z, const float *__restrict__ gp, const u64 *__restrict__ keys,
                                                            int B, int N, int C, const float *__restrict__ bnp, double *__restrict__ acc,
                                                            unsigned *__restrict__ mxdy) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= B * C) return;
    const int b = i / C, c = i - b * C;
    const u64 key = keys[i];
    if (__uint_as_float((unsigned)(key >> 32)) > 0.0f) {
        const int n = (int)(0xFFFFFFFFu - (unsigned)key);
        const float g = gp[i];
        const float zh = fmaf(z[((size_t)b * N + n) * C + c], bnp[3 * C + c], -bnp[2 * C + c]);
        atomicAdd(acc + c, (double)g);
        atomicAdd(acc + kTMaxC + c, (double)g * (double)zh);
        atomicMax(mxdy + c, __float_as_uint(fabsf(g)));
    }
}

// hidden layers: dy = da * [y > 0];  acc[c] += sum dy, acc[256 + c] += sum dy * zhat, mxdy[c] = max |dy|
__global__ void __launch_bounds__(256) bn_bwd_stats_kernel(const float *__restrict__ z, const float *__restrict__ da, long long P, int C,
                                                          const float *__restrict__ bnp, double *__restrict__ acc, unsigned *__restrict__ mxdy) {
    __shared__ float red[2 * 1024];
    const ColGeom g(C);
    float v[2][4] = {};
    float mx[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (g.active) {
        const float4 sc = __ldg(reinterpret_cast<const float4 *>(bnp) + g.tx), sh = __ldg(reinterpret_cast<const float4 *>(bnp + C) + g.tx);
        const float4 mi = __ldg(reinterpret_cast<const float4 *>(bnp + 2 * C) + g.tx), is = __ldg(reinterpret_cast<const float4 *>(bnp + 3 * C) + g.tx);
        float4 t[4], d[4];
        rows4((long long)blockIdx.x * g.rpi + g.ty, (long long)gridDim.x * g.rpi, P,
              [&](int u, long long row, bool ok) {
                  t[u] = ok ? __ldg(reinterpret_cast<const float4 *>(z + row * C) + g.tx) : make_float4(0.f, 0.f, 0.f, 0.f);
                  d[u] = ok ? __ldg(reinterpret_cast<const float4 *>(da + row * C) + g.tx) : make_float4(0.f, 0.f, 0.f, 0.f);
              },
              [&](int u, long long row) {
                  const float4 tt = t[u], dd = d[u];
                  const float dy0 = fmaf(tt.x, sc.x, sh.x) > 0.0f ? dd.x : 0.0f, dy1 = fmaf(tt.y, sc.y, sh.y) > 0.0f ? dd.y : 0.0f;
                  const float dy2 = fmaf(tt.z, sc.z, sh.z) > 0.0f ? dd.z : 0.0f, dy3 = fmaf(tt.w, sc.w, sh.w) > 0.0f ? dd.w : 0.0f;
                  v[0][0] += dy0; v[0][1] += dy1; v[0][2] += dy2; v[0][3] += dy3;
                  v[1][0] = fmaf(dy0, fmaf(tt.x, is.x, -mi.x), v[1][0]); v[1][1] = fmaf(dy1, fmaf(tt.y, is.y, -mi.y), v[1][1]);
                  v[1][2] = fmaf(dy2, fmaf(tt.z, is.z, -mi.z), v[1][2]); v[1][3] = fmaf(dy3, fmaf(tt.w, is.w, -mi.w), v[1][3]);
                  mx[0] = fmaxf(mx[0], fabsf(dy0)); mx[1] = fmaxf(mx[1], fabsf(dy1));
                  mx[2] = fmaxf(mx[2], fabsf(dy2)); mx[3] = fmaxf(mx[3], fabsf(dy3));
              });
    }
    double out[2];
    col_reduce_sum<2>(v, g, C, red, out);
    if ((int)threadIdx.x < C) {
        atomicAdd(acc + threadIdx.x, out[0]);
        atomicAdd(acc + kTMaxC + threadIdx.x, out[1]);
    }
    float m;
    col_reduce_max(mx, g, C, red, m);
    if ((int)threadIdx.x < C && m > 0.0f) atomicMax(mxdy + threadIdx.x, __float_as_uint(m));
}

// The two means of the BatchNorm backward, the parameter gradients they are, and the power-of-two scale of dz, computed by
// EVERY CTA of the consuming kernel into shared memory (s_m[0..C) = mean(dy), s_m[256..256+C) = mean(dy * zhat), *s_scale);
// `writer` (one CTA) stores them for the kernels that follow and writes dgamma / dbeta / dbias.
struct BwdArgs {
    const double *acc;
    const unsigned *mxdy, *zmax;
    const float *bnp;
    int batch_stats;
    float *m12, *dscale;          // out: [2][256], [2] = scale, 1 / scale
    float *dgamma, *dbeta, *dbias;
};
__device__ __forceinline__ void bwd_prologue(long long P, int C, const BwdArgs &a, bool writer, float *s_m, float *s_scale) {
    __shared__ float red[8];
    const int c = threadIdx.x;
    float bound = 0.0f;
    if (c < C) {
        const double s1 = a.acc[c], s2 = a.acc[kTMaxC + c];
        const float m1 = a.batch_stats ? (float)(s1 / (double)P) : 0.0f, m2 = a.batch_stats ? (float)(s2 / (double)P) : 0.0f;
        s_m[c] = m1;
        s_m[kTMaxC + c] = m2;
        const float scale = __ldg(a.bnp + c);
        if (writer) {
            a.m12[c] = m1;
            a.m12[kTMaxC + c] = m2;
            if (a.dgamma) a.dgamma[c] = (float)s2;
            if (a.dbeta) a.dbeta[c] = (float)s1;
            if (a.dbias) a.dbias[c] = a.batch_stats ? 0.0f : (float)((double)scale * s1);   // sum_p dz (zero through batch statistics)
        }
        bound = fabsf(scale) * (__uint_as_float(a.mxdy[c]) + fabsf(m1) + __uint_as_float(a.zmax[c]) * fabsf(m2));
    }
    for (int o = 16; o > 0; o >>= 1) bound = fmaxf(bound, __shfl_xor_sync(0xffffffffu, bound, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = bound;
    __syncthreads();
    if (threadIdx.x == 0) {
        float mm = 0.0f;
        for (int k = 0; k < 8; ++k) mm = fmaxf(mm, red[k]);
        int e = 0;
        if (mm > 0.0f && mm < 3.0e38f) e = 14 - (ilogbf(mm) + 1);
        e = max(-100, min(100, e));
        *s_scale = ldexpf(1.0f, e);
        if (writer) {
            a.dscale[0] = ldexpf(1.0f, e);
            a.dscale[1] = ldexpf(1.0f, -e);
        }
    }
    __syncthreads();
}

// dz = scale * (dy - m1 - zhat * m2), written as fp16 hi+lo operand rows multiplied by the power-of-two scale s.
// LAST: the dense part only (dy = 0): the few (cloud, channel) arg-max entries are patched by pool_bwd_patch_kernel.
template <bool LAST>
__global__ void __launch_bounds__(256) bn_bwd_dz_kernel(const float *__restrict__ z, const float *__restrict__ da, long long P, int C,
                                                       BwdArgs bw, unsigned short *__restrict__ dzhi, unsigned short *__restrict__ dzlo) {
    __shared__ __align__(16) float s_m[2 * kTMaxC];
    __shared__ float s_scale;
    bwd_prologue(P, C, bw, blockIdx.x == 0, s_m, &s_scale);
    const ColGeom g(C);
    if (!g.active) return;
    const float s = s_scale;
    const float4 sc = __ldg(reinterpret_cast<const float4 *>(bw.bnp) + g.tx), sh = __ldg(reinterpret_cast<const float4 *>(bw.bnp + C) + g.tx);
    const float4 mi = __ldg(reinterpret_cast<const float4 *>(bw.bnp + 2 * C) + g.tx), is = __ldg(reinterpret_cast<const float4 *>(bw.bnp + 3 * C) + g.tx);
    const float4 m1 = reinterpret_cast<const float4 *>(s_m)[g.tx], m2 = reinterpret_cast<const float4 *>(s_m + kTMaxC)[g.tx];
    const float4 ss = make_float4(sc.x * s, sc.y * s, sc.z * s, sc.w * s);
    float4 t[4], d[4];
    rows4((long long)blockIdx.x * g.rpi + g.ty, (long long)gridDim.x * g.rpi, P,
          [&](int u, long long row, bool ok) {
              t[u] = ok ? __ldg(reinterpret_cast<const float4 *>(z + row * C) + g.tx) : make_float4(0.f, 0.f, 0.f, 0.f);
              if (!LAST) d[u] = ok ? __ldg(reinterpret_cast<const float4 *>(da + row * C) + g.tx) : make_float4(0.f, 0.f, 0.f, 0.f);
          },
          [&](int u, long long row) {
              const float4 tt = t[u];
              float dy0 = 0.0f, dy1 = 0.0f, dy2 = 0.0f, dy3 = 0.0f;
              if (!LAST) {
                  const float4 dd = d[u];
                  dy0 = fmaf(tt.x, sc.x, sh.x) > 0.0f ? dd.x : 0.0f; dy1 = fmaf(tt.y, sc.y, sh.y) > 0.0f ? dd.y : 0.0f;
                  dy2 = fmaf(tt.z, sc.z, sh.z) > 0.0f ? dd.z : 0.0f; dy3 = fmaf(tt.w, sc.w, sh.w) > 0.0f ? dd.w : 0.0f;
              }
              const float v0 = ss.x * (dy0 - m1.x - fmaf(tt.x, is.x, -mi.x) * m2.x), v1 = ss.y * (dy1 - m1.y - fmaf(tt.y, is.y, -mi.y) * m2.y);
              const float v2 = ss.z * (dy2 - m1.z - fmaf(tt.z, is.z, -mi.z) * m2.z), v3 = ss.w * (dy3 - m1.w - fmaf(tt.w, is.w, -mi.w) * m2.w);
              uint32_t h0, l0, h1, l1;
              split_f16x2(v0, v1, h0, l0);
              split_f16x2(v2, v3, h1, l1);
              const size_t o = (size_t)row * C + 4 * g.tx;
              *reinterpret_cast<uint2 *>(dzhi + o) = make_uint2(h0, h1);
              *reinterpret_cast<uint2 *>(dzlo + o) = make_uint2(l0, l1);
          });
}

// last layer: the (cloud, channel) entries at the arg-max points carry the pooled upstream gradient
__global__ void __launch_bounds__(256) pool_bwd_patch_kernel(const float *__restrict__ z, const float *__restrict__ gp, const u64 *__restrict__ keys,
                                                            int B, int N, int C, const float *__restrict__ bnp, const float *__restrict__ m12,
                                                            const float *__restrict__ dscale, unsigned short *__restrict__ dzhi,
                                                            unsigned short *__restrict__ dzlo) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= B * C) return;
    const int b = i / C, c = i - b * C;
    const u64 key = keys[i];
    if (!(__uint_as_float((unsigned)(key >> 32)) > 0.0f)) return;
    const int n = (int)(0xFFFFFFFFu - (unsigned)key);
    const size_t o = ((size_t)b * N + n) * C + c;
    const float zh = fmaf(z[o], bnp[3 * C + c], -bnp[2 * C + c]);
    const float v = (bnp[c] * dscale[0]) * (gp[i] - m12[c] - zh * m12[kTMaxC + c]);
    const __half h = __float2half_rn(fminf(fmaxf(v, -65504.0f), 65504.0f));
    const __half lo = __float2half_rn(v - __half2float(h));
    dzhi[o] = *reinterpret_cast<const unsigned short *>(&h);
    dzlo[o] = *reinterpret_cast<const unsigned short *>(&lo);
}

// layer 0: dW0[c][k] = sum_p dz0[p][c] * x[p][k] on the CUDA cores (K = 3), dz0 formed on the fly
__global__ void __launch_bounds__(256) layer0_wgrad_kernel(const float *__restrict__ z, const float *__restrict__ da, const float *__restrict__ x,
                                                          long long P, int C, BwdArgs bw, double *__restrict__ w0acc) {
    __shared__ float red[3 * 1024];
    __shared__ __align__(16) float s_m[2 * kTMaxC];
    __shared__ float s_scale;
    bwd_prologue(P, C, bw, blockIdx.x == 0, s_m, &s_scale);
    const ColGeom g(C);
    float v[3][4] = {};
    if (g.active) {
        const float *bnp = bw.bnp;
        const float4 sc = __ldg(reinterpret_cast<const float4 *>(bnp) + g.tx), sh = __ldg(reinterpret_cast<const float4 *>(bnp + C) + g.tx);
        const float4 mi = __ldg(reinterpret_cast<const float4 *>(bnp + 2 * C) + g.tx), is = __ldg(reinterpret_cast<const float4 *>(bnp + 3 * C) + g.tx);
        const float4 m1 = reinterpret_cast<const float4 *>(s_m)[g.tx], m2 = reinterpret_cast<const float4 *>(s_m + kTMaxC)[g.tx];
        const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w}, miv[4] = {mi.x, mi.y, mi.z, mi.w};
        const float isv[4] = {is.x, is.y, is.z, is.w}, m1v[4] = {m1.x, m1.y, m1.z, m1.w}, m2v[4] = {m2.x, m2.y, m2.z, m2.w};
        float4 t[4], d[4];
        float px[4], py[4], pz[4];
        rows4((long long)blockIdx.x * g.rpi + g.ty, (long long)gridDim.x * g.rpi, P,
              [&](int u, long long row, bool ok) {
                  t[u] = ok ? __ldg(reinterpret_cast<const float4 *>(z + row * C) + g.tx) : make_float4(0.f, 0.f, 0.f, 0.f);
                  d[u] = ok ? __ldg(reinterpret_cast<const float4 *>(da + row * C) + g.tx) : make_float4(0.f, 0.f, 0.f, 0.f);
                  px[u] = ok ? __ldg(x + 3 * row) : 0.0f; py[u] = ok ? __ldg(x + 3 * row + 1) : 0.0f; pz[u] = ok ? __ldg(x + 3 * row + 2) : 0.0f;
              },
              [&](int u, long long row) {
                  const float tz[4] = {t[u].x, t[u].y, t[u].z, t[u].w}, dd[4] = {d[u].x, d[u].y, d[u].z, d[u].w};
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                      const float dy = fmaf(tz[k], scv[k], shv[k]) > 0.0f ? dd[k] : 0.0f;
                      const float dz = scv[k] * (dy - m1v[k] - fmaf(tz[k], isv[k], -miv[k]) * m2v[k]);
                      v[0][k] = fmaf(dz, px[u], v[0][k]);
                      v[1][k] = fmaf(dz, py[u], v[1][k]);
                      v[2][k] = fmaf(dz, pz[u], v[2][k]);
                  }
              });
    }
    double out[3];
    col_reduce_sum<3>(v, g, C, red, out);
    if ((int)threadIdx.x < C) {
#pragma unroll
        for (int k = 0; k < 3; ++k) atomicAdd(w0acc + 3 * threadIdx.x + k, out[k]);
    }
}
__global__ void __launch_bounds__(256) layer0_wgrad_finish_kernel(const double *__restrict__ w0acc, int n, float *__restrict__ dw) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) dw[i] = (float)w0acc[i];
}

// ---- weight gradient on the tensor cores: dW[co][ci] = sum_p dz[p][co] * a[p][ci] ------------------------------------------
// Both operands are point-major rows (the K dimension of this GEMM is the slow axis), i.e. MN-major tcgen05 operands: a TMA
// box of 64 channels x 64 points with the 128-byte swizzle IS the canonical MN-major SWIZZLE_128B tile (8-point groups 1024 B
// apart = SBO, the two 64-channel halves of a 128-wide operand one box apart = LBO).  A CTA owns one 128 x 128 tile of dW and
// a contiguous range of 64-point chunks; warp 0 feeds a three-stage TMA ring, warp 1 issues the MMAs (dealt over four TMEM
// accumulators, two of them carrying negated products), warps 2-5 drain the accumulators into registers every three
// chunks (192 points: at most four truncating additions per accumulator) and finally store the partial tile.
static constexpr int kWgThreads = 192;
static constexpr int kWgStages = 3;
static constexpr int kWgFlush = 3;                     // chunks per accumulator drain
static constexpr uint32_t kWgBox = 64 * 128;           // bytes of one box: 64 points x 64 channels x 2 B
static constexpr uint32_t kWgStage = 8 * kWgBox;       // A: 2 pieces x 2 boxes, B: 2 pieces x 2 boxes

struct WgradArgs {
    int chunks, tiles_n, n_tiles, n_split, C_out, C_in;
    float *partial;                                    // [n_split][n_tiles][128][128]
};

// MN-major SWIZZLE_128B shared-memory matrix descriptor: LBO = one box (between 64-wide MN halves), SBO = 1024 B (8 K rows)
__device__ __forceinline__ uint64_t umma_desc_mn(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(kWgBox >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}

__global__ void __launch_bounds__(kWgThreads, 1)
encoder_wgrad_kernel(const __grid_constant__ CUtensorMap tma0, const __grid_constant__ CUtensorMap tma1,
                     const __grid_constant__ CUtensorMap tmb0, const __grid_constant__ CUtensorMap tmb1, WgradArgs a) {
    extern __shared__ unsigned char wg_smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t pad = (1024u - (smem_u32(wg_smem_raw) & 1023u)) & 1023u;
    unsigned char *smem = wg_smem_raw + pad;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bars = sbase + kWgStages * kWgStage;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * kWgStages, bar_accfull = bar_empty + 8 * kWgStages, bar_accempty = bar_accfull + 8;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + kWgStages * kWgStage + 8 * (2 * kWgStages + 2));

    if (tid == 0) {
        for (int s = 0; s < kWgStages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_accfull, 1);
        mbar_init(bar_accempty, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_launch_dependents();
    pdl_wait();

    const int tile = (int)blockIdx.x % a.n_tiles, split = (int)blockIdx.x / a.n_tiles;
    const int tm = tile / a.tiles_n, tn = tile - tm * a.tiles_n;
    const int c_begin = (int)((long long)split * a.chunks / a.n_split), c_end = (int)((long long)(split + 1) * a.chunks / a.n_split);
    const int n_chunks = c_end - c_begin, n_groups = (n_chunks + kWgFlush - 1) / kWgFlush;

    if (warp == 0) {
        if (elect_one()) {
            for (int i = 0; i < n_chunks; ++i) {
                const uint32_t s = (uint32_t)i % kWgStages, use = (uint32_t)i / kWgStages;
                mbar_wait_wd(bar_empty + 8 * s, (use & 1u) ^ 1u);
                mbar_arrive_expect_tx(bar_full + 8 * s, kWgStage);
                const uint32_t dst = sbase + s * kWgStage;
                const int p0 = (c_begin + i) * 64;
                for (int h = 0; h < 2; ++h) {                       // the two 64-channel halves of the 128-wide tile
                    tma_load_2d(dst + (0 + h) * kWgBox, &tma0, tm * 128 + h * 64, p0, bar_full + 8 * s);
                    tma_load_2d(dst + (2 + h) * kWgBox, &tma1, tm * 128 + h * 64, p0, bar_full + 8 * s);
                    tma_load_2d(dst + (4 + h) * kWgBox, &tmb0, tn * 128 + h * 64, p0, bar_full + 8 * s);
                    tma_load_2d(dst + (6 + h) * kWgBox, &tmb1, tn * 128 + h * 64, p0, bar_full + 8 * s);
                }
            }
        }
    } else if (warp == 1) {
        const bool leader = elect_one();
        // D fp32, A/B fp16, both MN-major, M = N = 128
        const uint32_t idesc = (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc_neg = idesc | (1u << 13);                // a_negate
        int i = 0;
        for (int grp = 0; grp < n_groups; ++grp) {
            mbar_wait_wd(bar_accempty, ((uint32_t)grp & 1u) ^ 1u);
            tc_fence_after();
            uint32_t used = 0, ks = 0;
            const int i_end = min(n_chunks, i + kWgFlush);
            for (; i < i_end; ++i) {
                const uint32_t s = (uint32_t)i % kWgStages, use = (uint32_t)i / kWgStages;
                mbar_wait_wd(bar_full + 8 * s, use & 1u);
                tc_fence_after();
                if (leader) {
                    const uint32_t st = sbase + s * kWgStage;
                    const uint64_t ah = umma_desc_mn(st), al = umma_desc_mn(st + 2 * kWgBox);
                    const uint64_t bh = umma_desc_mn(st + 4 * kWgBox), bl = umma_desc_mn(st + 6 * kWgBox);
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4, ++ks) {          // 16 points per MMA = two 8-point groups = 2048 bytes
                        const uint64_t ko = (uint64_t)(k4 * (2048 >> 4));
                        // accumulators 1 and 3 take negated products and are subtracted when drained: the fp32 accumulation of
                        // the tensor core rounds toward minus infinity, and carrying half of the sum negated cancels that bias.
                        // Per step: hi.hi -> accumulator ks % 4, the two cross terms -> its neighbours of the other sign.
                        const uint32_t m = ks & 3u, m1 = (ks + 1u) & 3u, m3 = (ks + 3u) & 3u;
                        tc_mma_bf16(tmem + m * 128u, ah + ko, bh + ko, (m & 1u) ? idesc_neg : idesc, (used >> m) & 1u);     // hi . hi
                        used |= 1u << m;
                        tc_mma_bf16(tmem + m1 * 128u, al + ko, bh + ko, (m1 & 1u) ? idesc_neg : idesc, (used >> m1) & 1u);  // lo . hi
                        used |= 1u << m1;
                        tc_mma_bf16(tmem + m3 * 128u, ah + ko, bl + ko, (m3 & 1u) ? idesc_neg : idesc, (used >> m3) & 1u);  // hi . lo
                        used |= 1u << m3;
                    }
                    tc_commit(bar_empty + 8 * s);
                }
                __syncwarp();
            }
            if (leader) tc_commit(bar_accfull);
            __syncwarp();
        }
    } else {
        const int q = warp & 3;
        const int row = q * 32 + lane;                              // output channel of the tile = TMEM lane
        const uint32_t lane_base = ((uint32_t)q * 32u) << 16;
        float r[128];
#pragma unroll
        for (int e = 0; e < 128; ++e) r[e] = 0.0f;
        for (int grp = 0; grp < n_groups; ++grp) {
            mbar_wait_wd(bar_accfull, (uint32_t)grp & 1u);
            tc_fence_after();
            // a group of fewer than three 16-point steps cannot happen (a chunk is four steps): all four accumulators are live
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float v[32], vx[32];
                tc_ld32(tmem + lane_base + (uint32_t)(c * 32), v);
#pragma unroll 1
                for (int m = 1; m < 4; ++m) {
                    tc_ld32(tmem + lane_base + (uint32_t)m * 128u + (uint32_t)(c * 32), vx);
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = (m & 1) ? v[e] - vx[e] : v[e] + vx[e];
                }
#pragma unroll
                for (int e = 0; e < 32; ++e) r[c * 32 + e] += v[e];
            }
            tc_fence_before();
            mbar_arrive(bar_accempty);
        }
        // only the part of the tile inside dW is stored (and read back by the reduction)
        if (tm * 128 + row < a.C_out) {
            float4 *dst = reinterpret_cast<float4 *>(a.partial + (((size_t)split * a.n_tiles + tile) * 128 + row) * 128);
            const int n4 = min(32, (a.C_in - tn * 128) >> 2);
#pragma unroll
            for (int e4 = 0; e4 < 32; ++e4)
                if (e4 < n4) dst[e4] = make_float4(r[4 * e4], r[4 * e4 + 1], r[4 * e4 + 2], r[4 * e4 + 3]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

// dW = inv_scale * sum over the splits of the partial tiles, in a fixed order (deterministic): 64 outputs per CTA, four
// threads per output each summing every fourth split, combined in shared memory
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float *__restrict__ partial, int n_split, int n_tiles, int tiles_n, int C_out,
                                                          int C_in, const float *__restrict__ dscale, float *__restrict__ dw) {
    __shared__ float red[4][64];
    const int e = threadIdx.x & 63, j = threadIdx.x >> 6;
    const int i = blockIdx.x * 64 + e;
    float s = 0.0f;
    if (i < C_out * C_in) {
        const int co = i / C_in, ci = i - co * C_in;
        const int tile = (co >> 7) * tiles_n + (ci >> 7);
        const float *src = partial + ((size_t)tile * 128 + (co & 127)) * 128 + (ci & 127);
        for (int k = j; k < n_split; k += 4) s += __ldg(src + (size_t)k * n_tiles * 128 * 128);
    }
    red[j][e] = s;
    __syncthreads();
    if (j == 0 && i < C_out * C_in) dw[i] = ((red[0][e] + red[1][e]) + (red[2][e] + red[3][e])) * __ldg(dscale + 1);
}

// ---- host side --------------------------------------------------------------------------------------------------------
static int train_check(const char *fn, const rlg_bn_layer *layers, int L, unsigned flags) {
    if (flags & ~(RLG_ENC_BATCH_STATS | RLG_ENC_RESERVE_SMS(0xff))) return fail(RLG_ERR_UNSUPPORTED, "%s: unknown flag bits 0x%x", fn, flags);
    if (!layers || L < 2 || L > kTMaxLayers) return fail(RLG_ERR_UNSUPPORTED, "%s: need 2..%d layers, got %d", fn, kTMaxLayers, L);
    if (layers[0].c_in != 3) return fail(RLG_ERR_UNSUPPORTED, "%s: layer 0 must have c_in == 3", fn);
    for (int l = 0; l < L; ++l) {
        const rlg_bn_layer &y = layers[l];
        if (!y.w) return fail(RLG_ERR_NULL_POINTER, "%s: layer %d has no weights", fn, l);
        if (l > 0 && y.c_in != layers[l - 1].c_out) return fail(RLG_ERR_BAD_SHAPE, "%s: layer %d c_in %d != previous c_out %d", fn, l, y.c_in, layers[l - 1].c_out);
        if (y.c_out % 64 != 0 || y.c_out < 64 || y.c_out > kTMaxC)
            return fail(RLG_ERR_UNSUPPORTED, "%s: width %d is not a multiple of 64 in [64, %d]", fn, y.c_out, kTMaxC);
        if (!(flags & RLG_ENC_BATCH_STATS) && (!y.running_mean || !y.running_var))
            return fail(RLG_ERR_NULL_POINTER, "%s: layer %d needs running statistics in eval mode", fn, l);
        if (((uintptr_t)y.w | (uintptr_t)y.b | (uintptr_t)y.gamma | (uintptr_t)y.beta) & 15u)
            return fail(RLG_ERR_UNSUPPORTED, "%s: layer %d parameters must be 16-byte aligned", fn, l);
    }
    return 0;
}

static inline unsigned row_grid(long long P, int C, int sms) {
    const int rpi = 256 / (C / 4);
    long long want = (P + (long long)rpi * 8 - 1) / ((long long)rpi * 8);      // at least ~8 rows per thread
    const long long cap = (long long)sms * 4;
    if (want > cap) want = cap;
    if (want < 1) want = 1;
    return (unsigned)want;
}

static int launch_wgrad(const void *dzhi, const void *dzlo, const void *ahi, const void *alo, long long P, int C_out, int C_in,
                        const float *dscale, float *partial, float *dw, int sms, cudaStream_t st) {
    WgradArgs a;
    a.C_out = C_out; a.C_in = C_in;
    a.chunks = (int)((P + 63) / 64);
    const int tiles_m = (C_out + 127) / 128;
    a.tiles_n = (C_in + 127) / 128;
    a.n_tiles = tiles_m * a.tiles_n;
    // every CTA writes (and the reduction re-reads) a 64 KB partial tile: give a CTA at least four chunks (256 points)
    a.n_split = sms / a.n_tiles;
    if (a.n_split > a.chunks / 4) a.n_split = a.chunks / 4;
    if (a.n_split < 1) a.n_split = 1;
    a.partial = partial;
    CUtensorMap tm[4];
    const void *base[4] = {dzhi, dzlo, ahi, alo};
    for (int k = 0; k < 4; ++k) {
        const int C = k < 2 ? C_out : C_in;
        const cuuint64_t d[2] = {(cuuint64_t)C, (cuuint64_t)P};
        const cuuint64_t s[1] = {(cuuint64_t)C * 2};
        const cuuint32_t b[2] = {64, 64};
        int rc = make_map(&tm[k], 2, const_cast<void *>(base[k]), 2, d, s, b);
        if (rc) return rc;
    }
    const size_t smem_bytes = (size_t)kWgStages * kWgStage + 256 + 1024;
    cudaError_t ae = cudaFuncSetAttribute(encoder_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (ae != cudaSuccess) { cudaGetLastError(); return fail((int)ae, "encoder_wgrad_kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(ae)); }
    cudaError_t le = launch_pdl(encoder_wgrad_kernel, dim3((unsigned)(a.n_split * a.n_tiles)), dim3(kWgThreads), smem_bytes, st,
                                tm[0], tm[1], tm[2], tm[3], a);
    if (le != cudaSuccess) { cudaGetLastError(); return fail((int)le, "encoder_wgrad_kernel: %s", cudaGetErrorString(le)); }
    const int n = C_out * C_in;
    wgrad_reduce_kernel<<<(n + 63) / 64, 256, 0, st>>>(partial, a.n_split, a.n_tiles, a.tiles_n, C_out, C_in, dscale, dw);
    return 0;
}

static int launch_pack(const rlg_bn_layer *layers, int L, char *sv, const SavedLayout &sl, char *ws, const WsLayout &w, cudaStream_t st) {
    PackArgs pa = {};
    for (int l = 1; l < L; ++l) {
        pa.w[l] = layers[l].w;
        pa.c_out[l] = layers[l].c_out;
        pa.c_in[l] = layers[l].c_in;
        for (int k = 0; k < 4; ++k) pa.p[l][k] = reinterpret_cast<unsigned short *>(sv + sl.wp[l][k]);
    }
    pa.wamax = reinterpret_cast<unsigned *>(ws + w.wamax);
    pa.wscale = reinterpret_cast<float *>(sv + sl.wscale);
    train_wamax_kernel<<<dim3(kPackChunks, L - 1), 256, 0, st>>>(pa);
    train_pack_kernel<<<dim3(kPackChunks, L - 1), 256, 0, st>>>(pa);
    return 0;
}

}  // namespace rlg

using namespace rlg;

extern "C" {

size_t rlg_encoder_train_saved_bytes(int B, int N, const rlg_bn_layer *layers, int L) {
    if (B < 1 || N < 1 || train_check("rlg_encoder_train_saved_bytes", layers, L, RLG_ENC_BATCH_STATS) != 0) return 0;
    return saved_layout((long long)B * N, B, layers, L).total;
}

int rlg_encoder_train_saved_layout(int B, int N, const rlg_bn_layer *layers, int L, size_t *offsets, int n_offsets) {
    const char *fn = "rlg_encoder_train_saved_layout";
    if (B < 1 || N < 1) return fail(RLG_ERR_BAD_SHAPE, "%s: bad shape B=%d N=%d", fn, B, N);
    int rc = train_check(fn, layers, L, RLG_ENC_BATCH_STATS);
    if (rc) return rc;
    if (!offsets || n_offsets < 5 * L + 1) return fail(RLG_ERR_NULL_POINTER, "%s: need room for %d offsets", fn, 5 * L + 1);
    const SavedLayout sl = saved_layout((long long)B * N, B, layers, L);
    for (int l = 0; l < L; ++l) {
        offsets[5 * l] = sl.z[l];
        offsets[5 * l + 1] = sl.ahi[l];
        offsets[5 * l + 2] = sl.alo[l];
        offsets[5 * l + 3] = sl.bnp[l];
        offsets[5 * l + 4] = sl.zmax[l];
    }
    offsets[5 * L] = sl.keys;
    return 0;
}

size_t rlg_encoder_train_ws_bytes(int B, int N, const rlg_bn_layer *layers, int L) {
    if (B < 1 || N < 1 || train_check("rlg_encoder_train_ws_bytes", layers, L, RLG_ENC_BATCH_STATS) != 0) return 0;
    return ws_layout((long long)B * N, layers, L, sm_count()).total;
}

int rlg_encoder_train_fwd(const float *x, int B, int N, const rlg_bn_layer *layers, int L, unsigned flags, float *pooled, void *saved,
                          size_t saved_bytes, void *ws, size_t ws_bytes, void *stream) {
    const char *fn = "rlg_encoder_train_fwd";
    if (B < 1 || N < 1) return fail(RLG_ERR_BAD_SHAPE, "%s: bad shape B=%d N=%d", fn, B, N);
    int rc = train_check(fn, layers, L, flags);
    if (rc) return rc;
    const bool batch = (flags & RLG_ENC_BATCH_STATS) != 0;
    const long long P = (long long)B * N;
    if (batch && P < 2) return fail(RLG_ERR_BAD_SHAPE, "%s: batch statistics need more than one point per channel", fn);
    if (P > 0x7fffffffLL / 256) return fail(RLG_ERR_TOO_LARGE, "%s: B*N too large", fn);
    if (!x || !pooled) return fail(RLG_ERR_NULL_POINTER, "%s: null pointer", fn);
    const int sms_all = sm_count();
    if (sms_all <= 0) return fail((int)cudaErrorNoDevice, "%s: no CUDA device", fn);
    const int reserve = (int)((flags >> 16) & 0xffu);
    const int sms = sms_all - reserve > 1 ? sms_all - reserve : 1;          // grids of the persistent kernels
    const SavedLayout sl = saved_layout(P, B, layers, L);
    const WsLayout wl = ws_layout(P, layers, L, sms_all);
    if (!saved || saved_bytes < sl.total || ((uintptr_t)saved & 255u))
        return fail(RLG_ERR_WORKSPACE, "%s: saved buffer %p/%zu bytes, need %zu bytes 256-B aligned", fn, saved, saved_bytes, sl.total);
    if (!ws || ws_bytes < wl.total || ((uintptr_t)ws & 255u))
        return fail(RLG_ERR_WORKSPACE, "%s: workspace %p/%zu bytes, need %zu bytes 256-B aligned", fn, ws, ws_bytes, wl.total);
    cudaStream_t st = (cudaStream_t)stream;
    char *sv = (char *)saved, *w = (char *)ws;
    const int CL = layers[L - 1].c_out;

    cudaError_t e = cudaMemsetAsync(w, 0, wl.zero_bytes, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(sv + sl.keys, 0, (size_t)B * CL * 8, st);
    for (int l = 0; l < L && e == cudaSuccess; ++l) e = cudaMemsetAsync(sv + sl.zmax[l], 0, (size_t)layers[l].c_out * 4, st);
    if (e != cudaSuccess) { cudaGetLastError(); return fail((int)e, "%s: cudaMemsetAsync: %s", fn, cudaGetErrorString(e)); }
    rc = launch_pack(layers, L, sv, sl, w, wl, st);
    if (rc) return rc;

    for (int l = 0; l < L; ++l) {
        const rlg_bn_layer &y = layers[l];
        const int C = y.c_out;
        float *z = reinterpret_cast<float *>(sv + sl.z[l]);
        float *bnp = reinterpret_cast<float *>(sv + sl.bnp[l]);
        unsigned *zmax = reinterpret_cast<unsigned *>(sv + sl.zmax[l]);
        double *acc = reinterpret_cast<double *>(w + wl.stat) + (size_t)l * 2 * kTMaxC;
        if (l == 0) {
            const long long n = P * (C / 8);
            train_layer0_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, P, C, y.w, y.b, z);
        } else {
            GemmCall g;
            g.B = B; g.N = N; g.K = y.c_in; g.C_out = C; g.pieces = 2;
            g.x0 = sv + sl.ahi[l - 1]; g.x1 = sv + sl.alo[l - 1];
            g.w0 = sv + sl.wp[l][0]; g.w1 = sv + sl.wp[l][1];
            g.bias = y.b;
            g.out_scale = 1.0f;
            g.dscale0 = reinterpret_cast<const float *>(sv + sl.wscale) + 2 * l + 1;
            g.dscale1 = nullptr;
            g.epi = EPI_RAW_;
            g.y0 = z; g.y1 = nullptr; g.pooled = nullptr;
            rc = launch_layer_gemm(g, sms, st);
            if (rc) return rc;
        }
        if (batch) bn_stats_kernel<<<row_grid(P, C, sms), 256, 0, st>>>(z, P, C, acc);
        BnArgs bn;
        bn.acc = acc; bn.gamma = y.gamma; bn.beta = y.beta; bn.rmean = y.running_mean; bn.rvar = y.running_var;
        bn.eps = y.eps; bn.momentum = y.momentum; bn.batch_stats = batch ? 1 : 0; bn.bnp = bnp;
        if (l < L - 1) {
            bn_act_kernel<<<row_grid(P, C, sms), 256, 0, st>>>(z, P, C, bn, reinterpret_cast<unsigned short *>(sv + sl.ahi[l]),
                                                                reinterpret_cast<unsigned short *>(sv + sl.alo[l]), zmax);
        } else {
            const int rpi = 256 / (C / 4);
            int per = (N + 7) / 8;                                   // up to 8 CTAs per cloud
            per = (per + rpi - 1) / rpi * rpi;
            if (per < rpi) per = rpi;
            const unsigned gx = (unsigned)((N + per - 1) / per);
            u64 *keys = reinterpret_cast<u64 *>(sv + sl.keys);
            bn_pool_kernel<<<dim3(gx, (unsigned)B), 256, 0, st>>>(z, P, N, C, per, bn, keys, zmax);
            pool_decode_kernel<<<(B * C + 255) / 256, 256, 0, st>>>(keys, B * C, pooled);
        }
    }
    return check_launch(fn);
}

int rlg_encoder_train_bwd(const float *x, int B, int N, const rlg_bn_layer *layers, int L, unsigned flags, const float *g_pooled,
                          const void *saved, size_t saved_bytes, const rlg_bn_grads *grads, void *ws, size_t ws_bytes, void *stream) {
    const char *fn = "rlg_encoder_train_bwd";
    if (B < 1 || N < 1) return fail(RLG_ERR_BAD_SHAPE, "%s: bad shape B=%d N=%d", fn, B, N);
    int rc = train_check(fn, layers, L, flags);
    if (rc) return rc;
    const bool batch = (flags & RLG_ENC_BATCH_STATS) != 0;
    const long long P = (long long)B * N;
    if (P > 0x7fffffffLL / 256) return fail(RLG_ERR_TOO_LARGE, "%s: B*N too large", fn);
    if (!x || !g_pooled || !grads) return fail(RLG_ERR_NULL_POINTER, "%s: null pointer", fn);
    const int sms_all = sm_count();
    if (sms_all <= 0) return fail((int)cudaErrorNoDevice, "%s: no CUDA device", fn);
    const int reserve = (int)((flags >> 16) & 0xffu);
    const int sms = sms_all - reserve > 1 ? sms_all - reserve : 1;          // grids of the persistent kernels
    const SavedLayout sl = saved_layout(P, B, layers, L);
    const WsLayout wl = ws_layout(P, layers, L, sms_all);
    if (!saved || saved_bytes < sl.total || ((uintptr_t)saved & 255u))
        return fail(RLG_ERR_WORKSPACE, "%s: saved buffer %p/%zu bytes, need %zu bytes 256-B aligned", fn, saved, saved_bytes, sl.total);
    if (!ws || ws_bytes < wl.total || ((uintptr_t)ws & 255u))
        return fail(RLG_ERR_WORKSPACE, "%s: workspace %p/%zu bytes, need %zu bytes 256-B aligned", fn, ws, ws_bytes, wl.total);
    cudaStream_t st = (cudaStream_t)stream;
    const char *sv = (const char *)saved;
    char *w = (char *)ws;

    cudaError_t e = cudaMemsetAsync(w, 0, wl.zero_bytes, st);
    if (e != cudaSuccess) { cudaGetLastError(); return fail((int)e, "%s: cudaMemsetAsync: %s", fn, cudaGetErrorString(e)); }
    float *da = reinterpret_cast<float *>(w + wl.da);
    unsigned short *dzhi = reinterpret_cast<unsigned short *>(w + wl.dzhi), *dzlo = reinterpret_cast<unsigned short *>(w + wl.dzlo);
    const u64 *keys = reinterpret_cast<const u64 *>(sv + sl.keys);

    for (int l = L - 1; l >= 0; --l) {
        const rlg_bn_layer &y = layers[l];
        const int C = y.c_out;
        const float *z = reinterpret_cast<const float *>(sv + sl.z[l]);
        const float *bnp = reinterpret_cast<const float *>(sv + sl.bnp[l]);
        const unsigned *zmax = reinterpret_cast<const unsigned *>(sv + sl.zmax[l]);
        double *acc = reinterpret_cast<double *>(w + wl.stat) + (size_t)l * 2 * kTMaxC;
        unsigned *mxdy = reinterpret_cast<unsigned *>(w + wl.mxdy) + (size_t)l * kTMaxC;
        float *m12 = reinterpret_cast<float *>(w + wl.m12) + (size_t)l * 2 * kTMaxC;
        float *dscale = reinterpret_cast<float *>(w + wl.dscale) + 2 * l;
        const bool last = l == L - 1;
        if (last) pool_bwd_stats_kernel<<<(B * C + 255) / 256, 256, 0, st>>>(z, g_pooled, keys, B, N, C, bnp, acc, mxdy);
        else bn_bwd_stats_kernel<<<row_grid(P, C, sms), 256, 0, st>>>(z, da, P, C, bnp, acc, mxdy);
        BwdArgs bw;
        bw.acc = acc; bw.mxdy = mxdy; bw.zmax = zmax; bw.bnp = bnp; bw.batch_stats = batch ? 1 : 0;
        bw.m12 = m12; bw.dscale = dscale; bw.dgamma = grads[l].dgamma; bw.dbeta = grads[l].dbeta; bw.dbias = grads[l].db;
        if (l == 0) {
            double *w0acc = reinterpret_cast<double *>(w + wl.w0acc);
            layer0_wgrad_kernel<<<row_grid(P, C, sms), 256, 0, st>>>(z, da, x, P, C, bw, w0acc);
            if (grads[0].dw) layer0_wgrad_finish_kernel<<<(3 * C + 255) / 256, 256, 0, st>>>(w0acc, 3 * C, grads[0].dw);
            break;
        }
        if (last) {
            bn_bwd_dz_kernel<true><<<row_grid(P, C, sms), 256, 0, st>>>(z, nullptr, P, C, bw, dzhi, dzlo);
            pool_bwd_patch_kernel<<<(B * C + 255) / 256, 256, 0, st>>>(z, g_pooled, keys, B, N, C, bnp, m12, dscale, dzhi, dzlo);
        } else {
            bn_bwd_dz_kernel<false><<<row_grid(P, C, sms), 256, 0, st>>>(z, da, P, C, bw, dzhi, dzlo);
        }
        if (grads[l].dw) {
            rc = launch_wgrad(dzhi, dzlo, sv + sl.ahi[l - 1], sv + sl.alo[l - 1], P, C, y.c_in, dscale, reinterpret_cast<float *>(w + wl.partial),
                              grads[l].dw, sms, st);
            if (rc) return rc;
        }
        // da_{l-1} = dz_l . W_l   (the dead da_l buffer is overwritten)
        GemmCall g;
        g.B = B; g.N = N; g.K = C; g.C_out = y.c_in; g.pieces = 2;
        g.x0 = dzhi; g.x1 = dzlo;
        g.w0 = sv + sl.wp[l][2]; g.w1 = sv + sl.wp[l][3];           // the forward's transposed weight pieces
        g.bias = nullptr;
        g.out_scale = 1.0f;
        g.dscale0 = dscale + 1;
        g.dscale1 = reinterpret_cast<const float *>(sv + sl.wscale) + 2 * l + 1;
        g.epi = EPI_RAW_;
        g.y0 = da; g.y1 = nullptr; g.pooled = nullptr;
        rc = launch_layer_gemm(g, sms, st);
        if (rc) return rc;
    }
    return check_launch(fn);
}

}  // extern "C"
