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
// One launch of fire-and-forget float atomics onto zero-filled outputs.  The zero-fill costs nothing when the
// forward did it on the way (rlg_chamfer_loss_fwd's gz1/gz2); otherwise it is a memset node in front.
// (Shared-memory float atomics compile to CAS loops on sm_100a -- ATOMS.CAST.SPIN -- so accumulating in shared
// memory is slower than RED.ADD.F32 at L2.)
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

// One thread per point: u = w/d * (own - partner) is added to the point's own gradient row and subtracted from its
// partner's row with fire-and-forget float atomics (RED.ADD.F32 at L2) on zero-filled (or to-be-accumulated) outputs.
static constexpr int kBwdThreads = 256;

__global__ void __launch_bounds__(kBwdThreads) chamfer_bwd_kernel(BwdArgs a, int B) {
    pdl_launch_dependents();
    pdl_wait();                       // distances, indices and the upstream gradient come from the kernels before
    const int per_cloud = a.N + a.M;
    const long long t = (long long)blockIdx.x * kBwdThreads + threadIdx.x;
    const int b = (int)(t / per_cloud);
    if (b >= B) return;
    const int p = (int)(t - (long long)b * per_cloud);
    const int dir = p >= a.N, i = dir ? p - a.N : p;
    float ux, uy, uz;
    if (!bwd_term(a, dir, b, i, ux, uy, uz)) return;
    const int j = __ldg((dir ? a.i2 : a.i1) + (size_t)b * (dir ? a.M : a.N) + i);
    float *own = (dir ? a.gpc2 : a.gpc1) + ((size_t)b * (dir ? a.M : a.N) + i) * 3;
    float *oth = (dir ? a.gpc1 : a.gpc2) + ((size_t)b * (dir ? a.N : a.M) + j) * 3;
    atomicAdd(own, ux); atomicAdd(own + 1, uy); atomicAdd(own + 2, uz);
    atomicAdd(oth, -ux); atomicAdd(oth + 1, -uy); atomicAdd(oth + 2, -uz);
}

}  // namespace rlg

using namespace rlg;

static int bwd_launch(const float *pc1, const float *pc2, const float *d1, const float *d2, const int32_t *i1,
                      const int32_t *i2, const float *g1, const float *g2, int gstride, float scale1, float scale2,
                      int B, int N, int M, float *gpc1, float *gpc2, unsigned flags, void *stream) {
    if (B < 0 || N < 1 || M < 1)
        return fail(RLG_ERR_BAD_SHAPE, "rlg_chamfer_bwd: bad shape B=%d N=%d M=%d", B, N, M);
    if (B == 0) return 0;
    if (!pc1 || !pc2 || !d1 || !d2 || !i1 || !i2 || !gpc1 || !gpc2)
        return fail(RLG_ERR_NULL_POINTER, "rlg_chamfer_bwd: null pointer");
    BwdArgs a{pc1, pc2, d1, d2, i1, i2, g1, g2, gstride, scale1, scale2, gpc1, gpc2, N, M};
    cudaStream_t st = (cudaStream_t)stream;
    if (!(flags & RLG_CHAMFER_BWD_ACCUMULATE)) {
        cudaError_t e = cudaMemsetAsync(gpc1, 0, sizeof(float) * 3 * (size_t)B * N, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(gpc2, 0, sizeof(float) * 3 * (size_t)B * M, st);
        if (e != cudaSuccess) { cudaGetLastError(); return fail((int)e, "rlg_chamfer_bwd: cudaMemsetAsync: %s", cudaGetErrorString(e)); }
    }
    const long long total = (long long)B * ((long long)N + M);
    const long long blocks = (total + kBwdThreads - 1) / kBwdThreads;
    if (blocks > 0x7fffffffLL) return fail(RLG_ERR_TOO_LARGE, "rlg_chamfer_bwd: too many points");
    cudaError_t le = launch_pdl(chamfer_bwd_kernel, dim3((unsigned)blocks), dim3(kBwdThreads), 0, st, a, B);
    if (le != cudaSuccess) { cudaGetLastError(); return fail((int)le, "chamfer_bwd_kernel: %s", cudaGetErrorString(le)); }
    return check_launch("chamfer_bwd_kernel");
}

extern "C" int rlg_chamfer_bwd(const float *pc1, const float *pc2, const float *d1, const float *d2,
                               const int32_t *i1, const int32_t *i2, const float *g1, const float *g2, int B,
                               int N, int M, float *gpc1, float *gpc2, unsigned flags, void *stream) {
    return bwd_launch(pc1, pc2, d1, d2, i1, i2, g1, g2, 1, 1.0f, 1.0f, B, N, M, gpc1, gpc2, flags, stream);
}

extern "C" int rlg_chamfer_loss_bwd(const float *pc1, const float *pc2, const float *d1, const float *d2,
                                    const int32_t *i1, const int32_t *i2, const float *gloss, float w1, float w2,
                                    int B, int N, int M, float *gpc1, float *gpc2, unsigned flags, void *stream) {
    if (!gloss) return fail(RLG_ERR_NULL_POINTER, "rlg_chamfer_loss_bwd: null upstream gradient");
    return bwd_launch(pc1, pc2, d1, d2, i1, i2, gloss, w2 != 0.0f ? gloss : nullptr, 0, w1, w2, B, N, M, gpc1, gpc2,
                      flags, stream);
}
