// batch_prepare.cu -- the numeric part of the reference's input pipeline on the device, one CTA per cloud.
//
// Replaces, for a batch of B items taken from a binary cache of complete clouds (items x N x 3 fp32, resident in HBM):
//   utils/dataset.py:252-276   _create_incomplete_pc   random subset, or removal of the points inside a sphere whose radius
//                                                      is np.percentile(distances, ratio * 100) (float64, numpy's lerp)
//   utils/dataset.py:278-297   _augment_point_cloud    rotation pc @ R^T, clipped jitter, scale (each optional)
//   utils/data_utils.py:15-60  normalize_point_cloud   centre on the centroid, divide by the largest norm
//   utils/dataset.py:393-421   shapenet_collate_fn     pad the incomplete clouds to the longest of the batch by repeating points
// The reference does this per sample in Python/numpy inside DataLoader workers, after parsing text files with np.loadtxt.
// Every random decision is an INPUT here (struct rlg_prepare_plan): the host draws them (a few integers and floats per
// cloud), the device does the per-point work.  HBM-bound byte work: 12 B read + 2 x 12 B written per point.
#include "common.cuh"
#include <math.h>

namespace rlg {

static constexpr int kBpThreads = 256;
static constexpr int kBpMaxN = 4096;

__device__ __forceinline__ float block_sum_f(float v, float *red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.0f;
    for (int k = 0; k < kBpThreads / 32; ++k) s += red[k];
    return s;
}
__device__ __forceinline__ double block_sum_d(double v, double *red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int k = 0; k < kBpThreads / 32; ++k) s += red[k];
    return s;
}
__device__ __forceinline__ float block_max_f(float v, float *red) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.0f;
    for (int k = 0; k < kBpThreads / 32; ++k) s = fmaxf(s, red[k]);
    return s;
}

// augmentation (rotation, jitter, scale) + normalisation of `len` points held in shared memory, written to `out`
__device__ void augment_normalize_store(float *xs, float *ys, float *zs, int len, const float *rot, const float *noise, float scale,
                                        float *__restrict__ out, float *redf, double *redd) {
    const int tid = threadIdx.x;
    float R[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
    if (rot != nullptr) {
#pragma unroll
        for (int k = 0; k < 9; ++k) R[k] = __ldg(rot + k);
    }
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int i = tid; i < len; i += kBpThreads) {
        const float x = xs[i], y = ys[i], z = zs[i];
        float ax = fmaf(z, R[2], fmaf(y, R[1], x * R[0]));            // pc @ R^T: row i times row k of R
        float ay = fmaf(z, R[5], fmaf(y, R[4], x * R[3]));
        float az = fmaf(z, R[8], fmaf(y, R[7], x * R[6]));
        if (noise != nullptr) { ax += __ldg(noise + 3 * i); ay += __ldg(noise + 3 * i + 1); az += __ldg(noise + 3 * i + 2); }
        ax *= scale; ay *= scale; az *= scale;
        xs[i] = ax; ys[i] = ay; zs[i] = az;
        sx += (double)ax; sy += (double)ay; sz += (double)az;
    }
    sx = block_sum_d(sx, redd); sy = block_sum_d(sy, redd); sz = block_sum_d(sz, redd);
    const float cx = (float)(sx / (double)len), cy = (float)(sy / (double)len), cz = (float)(sz / (double)len);
    float mx = 0.0f;
    for (int i = tid; i < len; i += kBpThreads) {
        const float x = xs[i] - cx, y = ys[i] - cy, z = zs[i] - cz;
        xs[i] = x; ys[i] = y; zs[i] = z;
        mx = fmaxf(mx, sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z))));
    }
    mx = block_max_f(mx, redf);
    for (int i = tid; i < len; i += kBpThreads) {
        float x = xs[i], y = ys[i], z = zs[i];
        if (mx > 0.0f) { x = __fdiv_rn(x, mx); y = __fdiv_rn(y, mx); z = __fdiv_rn(z, mx); }
        out[3 * i] = x; out[3 * i + 1] = y; out[3 * i + 2] = z;
    }
    __syncthreads();
}

struct PrepareArgs {
    const float *cache;
    int n_items, N, B, P2;                 // P2 = N rounded up to a power of two (the sort's size)
    rlg_prepare_plan plan;
    float *complete_out, *incomplete_out;
    int32_t *lengths, *max_len;
};

__global__ void __launch_bounds__(kBpThreads) batch_prepare_kernel(PrepareArgs a) {
    extern __shared__ __align__(16) unsigned char bp_smem[];
    __shared__ float redf[kBpThreads / 32];
    __shared__ double redd[kBpThreads / 32];
    __shared__ int s_scan[kBpThreads / 32];
    __shared__ int s_len;
    const int N = a.N, P2 = a.P2, b = blockIdx.x, tid = threadIdx.x;
    double *dist = reinterpret_cast<double *>(bp_smem);                        // [P2] sorted copy of the distances
    float *xs = reinterpret_cast<float *>(dist + P2), *ys = xs + N, *zs = ys + N;   // working copy (complete, then incomplete)
    float *rx = zs + N, *ry = rx + N, *rz = ry + N;                            // the raw cloud
    const int item = min(max(a.plan.item ? a.plan.item[b] : b, 0), a.n_items - 1);     // never outside the cache
    const float *src = a.cache + (size_t)item * N * 3;
    for (int i = tid; i < N; i += kBpThreads) {
        const float x = __ldg(src + 3 * i), y = __ldg(src + 3 * i + 1), z = __ldg(src + 3 * i + 2);
        rx[i] = x; ry[i] = y; rz[i] = z;
        xs[i] = x; ys[i] = y; zs[i] = z;
    }
    __syncthreads();
    const size_t B = (size_t)a.B;
    // ---- complete cloud: augment, normalise
    augment_normalize_store(xs, ys, zs, N, a.plan.rot ? a.plan.rot + (size_t)b * 9 : nullptr,
                            a.plan.jitter ? a.plan.jitter + (size_t)b * N * 3 : nullptr, a.plan.scale ? a.plan.scale[b] : 1.0f,
                            a.complete_out + (size_t)b * N * 3, redf, redd);
    // ---- incomplete cloud: select from the RAW cloud
    int len;
    if (a.plan.method[b] == 0) {
        len = min(max(a.plan.n_keep[b], 0), N);
        const int32_t *keep = a.plan.keep_idx + (size_t)b * N;
        for (int s = tid; s < len; s += kBpThreads) {
            const int i = min(max(__ldg(keep + s), 0), N - 1);
            xs[s] = rx[i]; ys[s] = ry[i]; zs[s] = rz[i];
        }
        __syncthreads();
    } else {
        const int c = min(max(a.plan.center[b], 0), N - 1);
        const double cx = (double)rx[c], cy = (double)ry[c], cz = (double)rz[c];
        for (int i = tid; i < P2; i += kBpThreads) {
            double d = INFINITY;
            if (i < N) {
                const double dx = (double)rx[i] - cx, dy = (double)ry[i] - cy, dz = (double)rz[i] - cz;
                d = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));   // np.linalg.norm(.., axis=1)
            }
            dist[i] = d;
        }
        __syncthreads();
        for (int k = 2; k <= P2; k <<= 1) {                                    // bitonic sort, ascending
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int i = tid; i < P2; i += kBpThreads) {
                    const int p = i ^ j;
                    if (p > i) {
                        const double u = dist[i], v = dist[p];
                        const bool up = (i & k) == 0;
                        if ((u > v) == up) { dist[i] = v; dist[p] = u; }
                    }
                }
                __syncthreads();
            }
        }
        const int k = min(max(a.plan.q_index[b], 0), N - 1);
        const double lo = dist[k], hi = dist[min(k + 1, N - 1)], t = a.plan.q_gamma[b];
        const double diff = __dsub_rn(hi, lo);
        const double radius = t >= 0.5 ? __dsub_rn(hi, __dmul_rn(diff, __dsub_rn(1.0, t))) : __dadd_rn(lo, __dmul_rn(diff, t));   // numpy _lerp
        __syncthreads();
        // stable compaction of the points with distance > radius: every thread owns a contiguous run of points
        const int per = (N + kBpThreads - 1) / kBpThreads, i0 = tid * per, i1 = min(N, i0 + per);
        int cnt = 0;
        for (int i = i0; i < i1; ++i) {
            const double dx = (double)rx[i] - cx, dy = (double)ry[i] - cy, dz = (double)rz[i] - cz;
            const double d = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
            cnt += d > radius;
        }
        int incl = cnt;
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if ((tid & 31) >= o) incl += v;
        }
        if ((tid & 31) == 31) s_scan[tid >> 5] = incl;
        __syncthreads();
        int base = 0;
        for (int w = 0; w < (tid >> 5); ++w) base += s_scan[w];
        int pos = base + incl - cnt;
        if (tid == kBpThreads - 1) s_len = base + incl;
        for (int i = i0; i < i1; ++i) {
            const double dx = (double)rx[i] - cx, dy = (double)ry[i] - cy, dz = (double)rz[i] - cz;
            const double d = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
            if (d > radius) { xs[pos] = rx[i]; ys[pos] = ry[i]; zs[pos] = rz[i]; ++pos; }
        }
        __syncthreads();
        len = s_len;
    }
    if (len > 0)
        augment_normalize_store(xs, ys, zs, len, a.plan.rot ? a.plan.rot + (B + b) * 9 : nullptr,
                                a.plan.jitter ? a.plan.jitter + (B + b) * N * 3 : nullptr, a.plan.scale ? a.plan.scale[B + b] : 1.0f,
                                a.incomplete_out + (size_t)b * N * 3, redf, redd);
    if (tid == 0) {
        a.lengths[b] = len;
        atomicMax(a.max_len, len);
    }
}

// pad slot s >= len_b of cloud b repeats kept point pad_idx[b][s - len_b] % len_b (zeros for an empty cloud)
__global__ void __launch_bounds__(256) batch_pad_kernel(float *__restrict__ incomplete, const int32_t *__restrict__ lengths,
                                                       const int32_t *__restrict__ max_len, const int32_t *__restrict__ pad_idx, int N) {
    const int b = blockIdx.y;
    const int len = lengths[b], m = *max_len;
    const int s = len + blockIdx.x * 256 + threadIdx.x;
    if (s >= m) return;
    float *row = incomplete + (size_t)b * N * 3;
    float x = 0.0f, y = 0.0f, z = 0.0f;
    if (len > 0) {
        const int r = __ldg(pad_idx + (size_t)b * N + (s - len));
        const int j = (int)((unsigned)(r < 0 ? -r : r) % (unsigned)len);
        x = row[3 * j]; y = row[3 * j + 1]; z = row[3 * j + 2];
    }
    row[3 * s] = x; row[3 * s + 1] = y; row[3 * s + 2] = z;
}

}  // namespace rlg

using namespace rlg;

extern "C" int rlg_batch_prepare(const float *cache, int n_items, int N, int B, const rlg_prepare_plan *plan, float *complete_out,
                                 float *incomplete_out, int32_t *lengths, int32_t *max_len, void *stream) {
    const char *fn = "rlg_batch_prepare";
    if (B < 0 || N < 2 || n_items < 1) return fail(RLG_ERR_BAD_SHAPE, "%s: bad shape B=%d N=%d items=%d", fn, B, N, n_items);
    if (N > kBpMaxN) return fail(RLG_ERR_TOO_LARGE, "%s: clouds of more than %d points are not covered", fn, kBpMaxN);
    if (B == 0) return 0;
    if (!cache || !plan || !complete_out || !incomplete_out || !lengths || !max_len) return fail(RLG_ERR_NULL_POINTER, "%s: null pointer", fn);
    if (!plan->method || !plan->n_keep || !plan->keep_idx || !plan->center || !plan->q_index || !plan->q_gamma || !plan->pad_idx)
        return fail(RLG_ERR_NULL_POINTER, "%s: the plan needs method, n_keep, keep_idx, center, q_index, q_gamma and pad_idx", fn);
    cudaStream_t st = (cudaStream_t)stream;
    PrepareArgs a;
    a.cache = cache; a.n_items = n_items; a.N = N; a.B = B;
    a.P2 = 1;
    while (a.P2 < N) a.P2 <<= 1;
    a.plan = *plan;
    a.complete_out = complete_out; a.incomplete_out = incomplete_out; a.lengths = lengths; a.max_len = max_len;
    const size_t smem = (size_t)a.P2 * 8 + (size_t)N * 6 * 4;
    cudaError_t e = cudaMemsetAsync(max_len, 0, sizeof(int32_t), st);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(batch_prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cudaGetLastError(); return fail((int)e, "%s: %s", fn, cudaGetErrorString(e)); }
    batch_prepare_kernel<<<B, kBpThreads, smem, st>>>(a);
    batch_pad_kernel<<<dim3((N + 255) / 256, B), 256, 0, st>>>(incomplete_out, lengths, max_len, plan->pad_idx, N);
    return check_launch(fn);
}
