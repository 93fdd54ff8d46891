// encoder_fp32.cu -- PointNet encoder trunk, fp32 CUDA-core path (exact path, any layer widths).
//
// Replaces models/autoencoder.py:65-71 of the reference in eval mode: transpose, L x [Conv1d(k=1) ->
// BatchNorm1d -> ReLU] and torch.max over the points.  The reference round-trips a (B,C,N) activation
// through memory three times per layer; here a CTA keeps a 64-point tile of ONE cloud on chip through all
// layers (activations ping-pong in shared memory, channel-major so point-parallel reads are conflict free),
// streams each layer's folded weights through shared memory in 64-channel chunks, and max-pools the last
// layer straight out of registers -- the (B,C_last,N) activation never exists.  The pool is merged across
// CTAs with a 64-bit atomicMax on (relu_value_bits << 32 | ~point_index): post-ReLU values are >= 0 so they
// order as unsigned ints, and the inverted index makes the LOWEST point index win ties.
#include "common.cuh"

namespace rlg {

static constexpr int kMaxLayers = 8;
static constexpr int kP = 64;          // points per CTA tile
static constexpr int kOC = 64;         // output channels per weight chunk
static constexpr int kWStride = 68;    // padded row stride of the transposed weight chunk (16-B aligned rows)
static constexpr int kEncThreads = 256;

struct EncLayers {
    const float *w[kMaxLayers];
    const float *b[kMaxLayers];
    int cin[kMaxLayers];
    int cout[kMaxLayers];
    int L;
    int act_rows;   // rows of one activation buffer (max stored channel count)
};

__global__ void __launch_bounds__(kEncThreads) encoder_fp32_kernel(const float *__restrict__ x, int N,
                                                                  EncLayers lay, u64 *__restrict__ keys) {
    extern __shared__ __align__(16) float smem[];
    float *act0 = smem;
    float *act1 = act0 + (size_t)lay.act_rows * kP;
    float *wt = act1 + (size_t)lay.act_rows * kP;

    const int b = blockIdx.y;
    const int n0 = blockIdx.x * kP;
    const int tid = threadIdx.x;
    const int pg = tid & 15;     // points 4*pg .. 4*pg+3 of the tile
    const int og = tid >> 4;     // channels 4*og .. 4*og+3 of the chunk
    const int c_last = lay.cout[lay.L - 1];

    // layer-0 input: x (B,N,3) -> act0[c][p]
    for (int e = tid; e < kP * 3; e += kEncThreads) {
        const int p = e / 3, c = e % 3;
        const int n = n0 + p;
        act0[c * kP + p] = (n < N) ? x[((size_t)b * N + n) * 3 + c] : 0.0f;
    }
    float *cur = act0, *nxt = act1;

    for (int l = 0; l < lay.L; ++l) {
        const int cin = lay.cin[l], cout = lay.cout[l];
        const float *__restrict__ W = lay.w[l];
        const float *__restrict__ bias = lay.b[l];
        const bool last = (l == lay.L - 1);
        for (int oc0 = 0; oc0 < cout; oc0 += kOC) {
            __syncthreads();   // previous chunk's readers are done with wt; layer input is complete
            const int och = min(kOC, cout - oc0);
            for (int e = tid; e < kOC * cin; e += kEncThreads) {
                const int o = e / cin, c = e - o * cin;
                wt[c * kWStride + o] = (o < och) ? __ldg(W + (size_t)(oc0 + o) * cin + c) : 0.0f;
            }
            __syncthreads();
            float acc[4][4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
#pragma unroll
                for (int p = 0; p < 4; ++p) acc[k][p] = 0.0f;
            // two-level summation: 32-channel partial sums added into the total, so the rounding error grows with
            // sqrt(32) + sqrt(cin/32) instead of sqrt(cin) (keeps the trunk within 1e-5 of the float64 stack)
            for (int c0 = 0; c0 < cin; c0 += 32) {
                float part[4][4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int p = 0; p < 4; ++p) part[k][p] = 0.0f;
                const int c1 = min(c0 + 32, cin);
#pragma unroll 4
                for (int c = c0; c < c1; ++c) {
                    const float4 a = *reinterpret_cast<const float4 *>(cur + c * kP + 4 * pg);
                    const float4 w = *reinterpret_cast<const float4 *>(wt + c * kWStride + 4 * og);
                    const float av[4] = {a.x, a.y, a.z, a.w};
                    const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k)
#pragma unroll
                        for (int p = 0; p < 4; ++p) part[k][p] = __fmaf_rn(wv[k], av[p], part[k][p]);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k)
#pragma unroll
                    for (int p = 0; p < 4; ++p) acc[k][p] += part[k][p];
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int o = oc0 + 4 * og + k;
                const float bo = (o < cout) ? __ldg(bias + o) : 0.0f;
                float v[4];
#pragma unroll
                for (int p = 0; p < 4; ++p) v[p] = fmaxf(acc[k][p] + bo, 0.0f);
                if (!last) {
                    if (o < cout)
                        *reinterpret_cast<float4 *>(nxt + (size_t)o * kP + 4 * pg) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
                    u64 key = 0;
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const int n = n0 + 4 * pg + p;
                        if (n < N) {
                            const u64 kk = ((u64)__float_as_uint(v[p]) << 32) | (u64)(0xffffffffu - (unsigned)n);
                            key = kk > key ? kk : key;
                        }
                    }
#pragma unroll
                    for (int s = 8; s > 0; s >>= 1) {
                        const u64 other = __shfl_xor_sync(0xffffffffu, key, s);
                        key = other > key ? other : key;
                    }
                    if (pg == 0 && o < cout) atomicMax(&keys[(size_t)b * c_last + o], key);
                }
            }
        }
        float *t = cur; cur = nxt; nxt = t;
    }
}

__global__ void __launch_bounds__(256) encoder_unpack_kernel(const u64 *__restrict__ keys, long long n,
                                                            float *__restrict__ pooled,
                                                            int32_t *__restrict__ argmax) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const u64 k = keys[i];
    pooled[i] = __uint_as_float((unsigned)(k >> 32));
    if (argmax) argmax[i] = (int32_t)(0xffffffffu - (unsigned)(k & 0xffffffffu));
}

static int plan(const rlg_layer *layers, int L, EncLayers &lay, size_t &smem_bytes) {
    if (!layers || L < 1 || L > kMaxLayers)
        return fail(RLG_ERR_UNSUPPORTED, "rlg_encoder: need 1..%d layers, got %d", kMaxLayers, L);
    if (layers[0].c_in != 3) return fail(RLG_ERR_UNSUPPORTED, "rlg_encoder: layer 0 must have c_in == 3");
    int act_rows = 3, cin_max = 3;
    for (int l = 0; l < L; ++l) {
        if (!layers[l].w || !layers[l].b) return fail(RLG_ERR_NULL_POINTER, "rlg_encoder: layer %d has null weights", l);
        if (layers[l].c_in < 1 || layers[l].c_out < 1) return fail(RLG_ERR_BAD_SHAPE, "rlg_encoder: layer %d bad widths", l);
        if (l > 0 && layers[l].c_in != layers[l - 1].c_out)
            return fail(RLG_ERR_BAD_SHAPE, "rlg_encoder: layer %d c_in %d != previous c_out %d", l, layers[l].c_in,
                        layers[l - 1].c_out);
        lay.w[l] = layers[l].w; lay.b[l] = layers[l].b;
        lay.cin[l] = layers[l].c_in; lay.cout[l] = layers[l].c_out;
        if (layers[l].c_in > cin_max) cin_max = layers[l].c_in;
        if (l < L - 1 && layers[l].c_out > act_rows) act_rows = layers[l].c_out;
    }
    lay.L = L;
    lay.act_rows = act_rows;
    smem_bytes = sizeof(float) * ((size_t)2 * act_rows * kP + (size_t)cin_max * kWStride);
    if (smem_bytes > 227 * 1024)
        return fail(RLG_ERR_UNSUPPORTED, "rlg_encoder: layer widths need %zu bytes of shared memory (> 227 KB)", smem_bytes);
    return 0;
}

}  // namespace rlg

using namespace rlg;

extern "C" {

size_t rlg_encoder_ws_bytes(int B, int N, const rlg_layer *layers, int L) {
    (void)N;
    if (B < 0 || !layers || L < 1) return 0;
    return align_up(sizeof(u64) * (size_t)B * (size_t)layers[L - 1].c_out, 256);
}

int rlg_encoder_fwd(const float *x, int B, int N, const rlg_layer *layers, int L, float *pooled, int32_t *argmax,
                    void *ws, size_t ws_bytes, void *stream) {
    if (B < 0 || N < 1) return fail(RLG_ERR_BAD_SHAPE, "rlg_encoder_fwd: bad shape B=%d N=%d", B, N);
    if (B == 0) return 0;
    if (!x || !pooled) return fail(RLG_ERR_NULL_POINTER, "rlg_encoder_fwd: null pointer");
    if (B > 65535) return fail(RLG_ERR_TOO_LARGE, "rlg_encoder_fwd: B=%d exceeds 65535", B);
    EncLayers lay;
    size_t smem_bytes = 0;
    int rc = plan(layers, L, lay, smem_bytes);
    if (rc) return rc;
    const size_t need = rlg_encoder_ws_bytes(B, N, layers, L);
    if (!ws || ws_bytes < need || ((uintptr_t)ws & 255u))
        return fail(RLG_ERR_WORKSPACE, "rlg_encoder_fwd: workspace %p/%zu bytes, need %zu bytes 256-B aligned", ws,
                    ws_bytes, need);
    cudaStream_t st = (cudaStream_t)stream;
    const long long n_keys = (long long)B * lay.cout[L - 1];
    cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(u64) * (size_t)n_keys, st);
    if (e != cudaSuccess) { cudaGetLastError(); return fail((int)e, "rlg_encoder_fwd: cudaMemsetAsync: %s", cudaGetErrorString(e)); }
    e = cudaFuncSetAttribute(encoder_fp32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return fail((int)e, "rlg_encoder_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); }
    dim3 grid((N + kP - 1) / kP, B);
    encoder_fp32_kernel<<<grid, kEncThreads, smem_bytes, st>>>(x, N, lay, (u64 *)ws);
    rc = check_launch("encoder_fp32_kernel");
    if (rc) return rc;
    encoder_unpack_kernel<<<(unsigned)((n_keys + 255) / 256), 256, 0, st>>>((const u64 *)ws, n_keys, pooled, argmax);
    return check_launch("encoder_unpack_kernel");
}

}  // extern "C"
