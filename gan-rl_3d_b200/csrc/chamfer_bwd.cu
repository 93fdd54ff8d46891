// chamfer_bwd.cu -- backward of the batched Chamfer distance.
//
// Replaces the reference's autograd graph behind utils/losses.py:29-37 (MeanBackward, MinBackward0 x2,
// EuclideanDistBackward0), which materialises dense one-hot (B,N,M) gradient matrices and two
// (B,N,M)x(B,M,3) GEMMs.  With the argmin indices saved by the forward the gradient is a gather plus a
// scatter-add of 3 floats per point:
//
//   own term      gpc1[b,i]      =  g1[b]/N * (pc1[b,i] - pc2[b,i1[b,i]]) / d1[b,i]        (0 where d == 0)
//   partner term  gpc2[b,i1[..]] -= the same vector                                         (atomic)
//   and symmetrically for pc2 -> pc1 with g2[b]/M, i2, d2.
//
// One launch, no global atomics and no memset: each CTA accumulates a slice of the gradient rows of one cloud in
// shared memory (own terms with plain stores, partner terms with shared-memory atomics found by scanning the
// cloud's argmin indices) and writes it out coalesced.
// HBM-bound: 56 bytes per point; at the headline shape it is launch-latency bound.
#include "common.cuh"

namespace rlg {

struct BwdArgs {
    const float *pc1, *pc2, *d1, *d2;
    const int32_t *i1, *i2;
    const float *g1, *g2;   // upstream of mean1 / mean2: per pair (gstride 1) or one shared scalar (gstride 0)
    int gstride;
    float scale1, scale2;   // multiplies the upstream (loss weights 0.5/B etc. for the fused ChamferLoss)
    float *gpc1, *gpc2;
    int N, M;
};

// term of point `i` of cloud `b` in direction dir (0: pc1 -> pc2, 1: pc2 -> pc1): u = w/d * (own - partner);
// false (and u = 0) where the distance or the upstream weight is zero (EuclideanDistBackward0 gives 0 at d == 0)
__device__ __forceinline__ bool bwd_term(const BwdArgs &a, int dir, int b, int i, float &ux, float &uy, float &uz) {
    const int n = dir ? a.M : a.N, m = dir ? a.N : a.M;
    const size_t p = (size_t)b * n + i;
    const float *own = (dir ? a.pc2 : a.pc1) + 3 * p;
    const int32_t j = __ldg((dir ? a.i2 : a.i1) + p);
    const float d = __ldg((dir ? a.d2 : a.d1) + p);
    const float *g = dir ? a.g2 : a.g1;
    const float *oth = (dir ? a.pc1 : a.pc2) + 3 * ((size_t)b * m + j);
    const float w = g ? __ldg(g + (size_t)b * a.gstride) * (dir ? a.scale2 : a.scale1) / (float)n : 0.0f;
    if (d == 0.0f || w == 0.0f) { ux = uy = uz = 0.0f; return false; }
    const float s = w / d;
    ux = (__ldg(own) - __ldg(oth)) * s;
    uy = (__ldg(own + 1) - __ldg(oth + 1)) * s;
    uz = (__ldg(own + 2) - __ldg(oth + 2)) * s;
    return true;
}

// One launch.  CTA (c, b) owns rows [c*rn, (c+1)*rn) of gpc1[b] and rows [c*rm, (c+1)*rm) of gpc2[b] and
// accumulates them in shared memory (outputs fully overwritten: no memset, no global atomics):
//   WHOLE (one CTA per cloud, the common case: (N+M)*12 bytes <= 200 KB)
//     every point's term u is computed once, added to its own row and subtracted from its partner's row with
//     shared-memory atomics on a zeroed accumulator
//   chunked (larger clouds)
//     phase 1  own terms of the CTA's rows
//     phase 2  scan ALL argmin indices of the cloud (coalesced int32 reads); every point of the other direction
//              whose partner falls into this CTA's rows adds its term with a shared-memory atomic
//   phase 3  coalesced write-out
static constexpr int kBwdThreads = 1024;

template <bool WHOLE>
__global__ void __launch_bounds__(kBwdThreads) chamfer_bwd_fused_kernel(BwdArgs a, int rn, int rm) {
    extern __shared__ float s_grad[];
    float *s1 = s_grad, *s2 = s_grad + 3 * rn;
    const int b = blockIdx.y, c = blockIdx.x, tid = threadIdx.x;
    const int n0 = min(a.N, c * rn), n1 = min(a.N, n0 + rn);
    const int m0 = min(a.M, c * rm), m1 = min(a.M, m0 + rm);
    float ux, uy, uz;
    if (WHOLE) {
        // points 0..N-1 are pc1's, N..N+M-1 are pc2's (s1 and s2 are contiguous in the same numbering)
        const int total = a.N + a.M;
        for (int e = tid; e < 3 * total; e += kBwdThreads) s_grad[e] = 0.0f;
        __syncthreads();
        for (int p = tid; p < total; p += kBwdThreads) {
            const int dir = p >= a.N, i = dir ? p - a.N : p;
            if (bwd_term(a, dir, b, i, ux, uy, uz)) {
                const int partner = __ldg((dir ? a.i2 : a.i1) + (size_t)b * (dir ? a.M : a.N) + i) + (dir ? 0 : a.N);
                float *own = s_grad + 3 * p, *oth = s_grad + 3 * partner;
                atomicAdd(own, ux); atomicAdd(own + 1, uy); atomicAdd(own + 2, uz);
                atomicAdd(oth, -ux); atomicAdd(oth + 1, -uy); atomicAdd(oth + 2, -uz);
            }
        }
    } else {
        for (int i = n0 + tid; i < n1; i += kBwdThreads) {
            bwd_term(a, 0, b, i, ux, uy, uz);
            s1[3 * (i - n0)] = ux; s1[3 * (i - n0) + 1] = uy; s1[3 * (i - n0) + 2] = uz;
        }
        for (int j = m0 + tid; j < m1; j += kBwdThreads) {
            bwd_term(a, 1, b, j, ux, uy, uz);
            s2[3 * (j - m0)] = ux; s2[3 * (j - m0) + 1] = uy; s2[3 * (j - m0) + 2] = uz;
        }
        __syncthreads();
        if (n1 > n0) {                                   // pc2 points whose nearest pc1 point is one of my rows
            const int32_t *idx = a.i2 + (size_t)b * a.M;
            for (int j = tid; j < a.M; j += kBwdThreads) {
                const int t = __ldg(idx + j);
                if (t >= n0 && t < n1 && bwd_term(a, 1, b, j, ux, uy, uz)) {
                    atomicAdd(s1 + 3 * (t - n0), -ux); atomicAdd(s1 + 3 * (t - n0) + 1, -uy); atomicAdd(s1 + 3 * (t - n0) + 2, -uz);
                }
            }
        }
        if (m1 > m0) {
            const int32_t *idx = a.i1 + (size_t)b * a.N;
            for (int i = tid; i < a.N; i += kBwdThreads) {
                const int t = __ldg(idx + i);
                if (t >= m0 && t < m1 && bwd_term(a, 0, b, i, ux, uy, uz)) {
                    atomicAdd(s2 + 3 * (t - m0), -ux); atomicAdd(s2 + 3 * (t - m0) + 1, -uy); atomicAdd(s2 + 3 * (t - m0) + 2, -uz);
                }
            }
        }
    }
    __syncthreads();
    float *o1 = a.gpc1 + ((size_t)b * a.N + n0) * 3, *o2 = a.gpc2 + ((size_t)b * a.M + m0) * 3;
    for (int e = tid; e < 3 * (n1 - n0); e += kBwdThreads) o1[e] = s1[e];
    for (int e = tid; e < 3 * (m1 - m0); e += kBwdThreads) o2[e] = s2[e];
}

}  // namespace rlg

using namespace rlg;

static int bwd_launch(const float *pc1, const float *pc2, const float *d1, const float *d2, const int32_t *i1,
                      const int32_t *i2, const float *g1, const float *g2, int gstride, float scale1, float scale2,
                      int B, int N, int M, float *gpc1, float *gpc2, void *stream) {
    if (B < 0 || N < 1 || M < 1)
        return fail(RLG_ERR_BAD_SHAPE, "rlg_chamfer_bwd: bad shape B=%d N=%d M=%d", B, N, M);
    if (B == 0) return 0;
    if (!pc1 || !pc2 || !d1 || !d2 || !i1 || !i2 || !gpc1 || !gpc2)
        return fail(RLG_ERR_NULL_POINTER, "rlg_chamfer_bwd: null pointer");
    if (B > 65535) return fail(RLG_ERR_TOO_LARGE, "rlg_chamfer_bwd: B=%d exceeds 65535 (grid.y)", B);
    BwdArgs a{pc1, pc2, d1, d2, i1, i2, g1, g2, gstride, scale1, scale2, gpc1, gpc2, N, M};
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)N + M;
    cudaError_t e;
    if (total * 12 <= 200 * 1024) {
        const size_t smem = sizeof(float) * 3 * (size_t)total;
        if (smem > 40u * 1024u) {
            e = cudaFuncSetAttribute(chamfer_bwd_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
            if (e != cudaSuccess) { cudaGetLastError(); return fail((int)e, "rlg_chamfer_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); }
        }
        chamfer_bwd_fused_kernel<true><<<dim3(1, B), kBwdThreads, smem, st>>>(a, N, M);
        return check_launch("chamfer_bwd_fused_kernel");
    }
    // larger clouds: 4096 rows of each gradient per CTA (96 KB of shared memory)
    const int longest = N > M ? N : M;
    const int chunks = (longest + 4095) / 4096;
    const int rn = (N + chunks - 1) / chunks, rm = (M + chunks - 1) / chunks;
    const size_t smem = sizeof(float) * 3 * ((size_t)rn + rm);
    e = cudaFuncSetAttribute(chamfer_bwd_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) { cudaGetLastError(); return fail((int)e, "rlg_chamfer_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); }
    chamfer_bwd_fused_kernel<false><<<dim3(chunks, B), kBwdThreads, smem, st>>>(a, rn, rm);
    return check_launch("chamfer_bwd_fused_kernel");
}

extern "C" int rlg_chamfer_bwd(const float *pc1, const float *pc2, const float *d1, const float *d2,
                               const int32_t *i1, const int32_t *i2, const float *g1, const float *g2, int B,
                               int N, int M, float *gpc1, float *gpc2, void *stream) {
    return bwd_launch(pc1, pc2, d1, d2, i1, i2, g1, g2, 1, 1.0f, 1.0f, B, N, M, gpc1, gpc2, stream);
}

extern "C" int rlg_chamfer_loss_bwd(const float *pc1, const float *pc2, const float *d1, const float *d2,
                                    const int32_t *i1, const int32_t *i2, const float *gloss, float w1, float w2,
                                    int B, int N, int M, float *gpc1, float *gpc2, void *stream) {
    if (!gloss) return fail(RLG_ERR_NULL_POINTER, "rlg_chamfer_loss_bwd: null upstream gradient");
    return bwd_launch(pc1, pc2, d1, d2, i1, i2, gloss, w2 != 0.0f ? gloss : nullptr, 0, w1, w2, B, N, M, gpc1, gpc2,
                      stream);
}
