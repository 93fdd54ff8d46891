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
// Two launches: `own` writes every gradient row once with coalesced plain stores (so the outputs need no
// memset), `scatter` adds the partner terms with fire-and-forget float atomics (RED.ADD.F32).
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
    long long n1, n2;   // B*N, B*M
};

// term for point `p` (flat index into direction dir): returns u = w/d * (own - partner)
__device__ __forceinline__ bool bwd_term(const BwdArgs &a, int dir, long long p, float &ux, float &uy, float &uz,
                                         long long &partner_flat) {
    const int n = dir ? a.M : a.N, m = dir ? a.N : a.M;
    const long long b = p / n;
    const float *own = (dir ? a.pc2 : a.pc1) + 3 * p;
    const int32_t j = (dir ? a.i2 : a.i1)[p];
    const float d = (dir ? a.d2 : a.d1)[p];
    const float *g = dir ? a.g2 : a.g1;
    partner_flat = b * m + j;
    const float *oth = (dir ? a.pc1 : a.pc2) + 3 * partner_flat;
    const float w = g ? g[b * a.gstride] * (dir ? a.scale2 : a.scale1) / (float)n : 0.0f;
    if (d == 0.0f || w == 0.0f) { ux = uy = uz = 0.0f; return false; }
    const float s = w / d;
    ux = (own[0] - oth[0]) * s;
    uy = (own[1] - oth[1]) * s;
    uz = (own[2] - oth[2]) * s;
    return true;
}

__global__ void __launch_bounds__(256) chamfer_bwd_own_kernel(BwdArgs a) {
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long total = a.n1 + a.n2;
    if (t >= total) return;
    const int dir = t >= a.n1;
    const long long p = dir ? t - a.n1 : t;
    float ux, uy, uz;
    long long partner;
    bwd_term(a, dir, p, ux, uy, uz, partner);
    float *out = (dir ? a.gpc2 : a.gpc1) + 3 * p;
    out[0] = ux; out[1] = uy; out[2] = uz;
}

__global__ void __launch_bounds__(256) chamfer_bwd_scatter_kernel(BwdArgs a) {
    const long long t = (long long)blockIdx.x * 256 + threadIdx.x;
    const long long total = a.n1 + a.n2;
    if (t >= total) return;
    const int dir = t >= a.n1;
    const long long p = dir ? t - a.n1 : t;
    float ux, uy, uz;
    long long partner;
    if (!bwd_term(a, dir, p, ux, uy, uz, partner)) return;
    float *out = (dir ? a.gpc1 : a.gpc2) + 3 * partner;
    atomicAdd(out + 0, -ux);
    atomicAdd(out + 1, -uy);
    atomicAdd(out + 2, -uz);
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
    BwdArgs a{pc1, pc2, d1, d2, i1, i2, g1, g2, gstride, scale1, scale2, gpc1, gpc2, N, M, (long long)B * N,
              (long long)B * M};
    const long long total = a.n1 + a.n2;
    const long long blocks = (total + 255) / 256;
    if (blocks > 0x7fffffffLL) return fail(RLG_ERR_TOO_LARGE, "rlg_chamfer_bwd: too many points");
    cudaStream_t st = (cudaStream_t)stream;
    chamfer_bwd_own_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
    chamfer_bwd_scatter_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
    return check_launch("chamfer_bwd kernels");
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
