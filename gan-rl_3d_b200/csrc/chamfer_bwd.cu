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
// One launch of fire-and-forget float atomics (vector reductions: about 2.3 per point instead of 6 scalars) onto
// zero-filled outputs.  The zero-fill costs nothing when the forward did it on the way (rlg_chamfer_loss_fwd's gz1/gz2);
// otherwise it is a memset node in front.
// (Shared-memory float atomics compile to CAS loops on sm_100a -- ATOMS.CAST.SPIN -- so accumulating in shared
// memory is slower than RED.ADD.F32 at L2.)
// HBM-bound: 56 bytes per point; at the headline shape it is launch-latency bound.
//
// rlg_chamfer_bwd_det / rlg_chamfer_loss_bwd_det: the same gradient, reproducible run to run.  Float atomics add in the
// order of arrival; here the partner terms are summed as 64-bit fixed-point integers (RED.ADD.64, commutative AND
// associative) in a workspace and a second kernel adds each point's own term and converts.  Every term of direction
// `dir` in pair b has norm |w| = |g[b]|/n exactly, so the quantum 2^(ilogb|w| - s), s = min(40, 61 - ceil(log2 n)), keeps
// 16 more bits per term than fp32 does and the sum of n terms stays below 2^62.
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

// upstream weight of every term of direction dir in pair b
__device__ __forceinline__ float bwd_weight(const BwdArgs &a, int dir, int b) {
    const float *g = dir ? a.g2 : a.g1;
    return g ? __ldg(g + (size_t)b * a.gstride) * (dir ? a.scale2 : a.scale1) / (float)(dir ? a.M : a.N) : 0.0f;
}

// term of point `i` of cloud `b` in direction dir (0: pc1 -> pc2, 1: pc2 -> pc1): u = w/d * (own - partner);
// false (and u = 0) where the distance or the upstream weight is zero (EuclideanDistBackward0 gives 0 at d == 0)
__device__ __forceinline__ bool bwd_term(const BwdArgs &a, int dir, int b, int i, float &ux, float &uy, float &uz, int &j) {
    const int n = dir ? a.M : a.N, m = dir ? a.N : a.M;
    const size_t p = (size_t)b * n + i;
    const float *own = (dir ? a.pc2 : a.pc1) + 3 * p;
    j = __ldg((dir ? a.i2 : a.i1) + p);
    const float d = __ldg((dir ? a.d2 : a.d1) + p);
    const float *oth = (dir ? a.pc1 : a.pc2) + 3 * ((size_t)b * m + j);
    const float w = bwd_weight(a, dir, b);
    if (d == 0.0f || w == 0.0f) { ux = uy = uz = 0.0f; return false; }
    const float s = w / d;
    ux = (__ldg(own) - __ldg(oth)) * s;
    uy = (__ldg(own + 1) - __ldg(oth + 1)) * s;
    uz = (__ldg(own + 2) - __ldg(oth + 2)) * s;
    return true;
}

// Row `row` of a (rows, 3) fp32 array whose base is 16-byte aligned (VEC 2) sits inside ONE aligned 16-byte window when its
// float offset 3*row is 0 or 1 mod 4 (the window's fourth float belongs to a neighbouring row), else across two: one 4-float
// reduction (else a 2-float one + a scalar) instead of three scalars.  Measured at cfg5 (64 x 16384^2, memsets included):
// 52.9 us with scalars, 42.4 with 2-float reductions, 40.3 with 4-float ones, 37.7 at 32 registers (8 CTAs per SM).  The same
// windows as LOADS (4-float own rows through shared memory, 4-float partner rows) were measured and dropped: 41-43 us, the
// longer instruction stream costs more than the saved L1 requests.  The window of the very last row would end outside the
// buffer: that row takes the two-access form.
// Adds (x, y, z) onto gradient row `row` with L2 reductions (fire and forget).  VEC 1: base 8-byte aligned.
template <int VEC>
__device__ __forceinline__ void red_add_row(float *base, size_t row, size_t rows, float x, float y, float z) {
    const size_t f = row * 3;
    float *p = base + f;
#ifndef RLG_BWD_SCALAR_RED
    if (VEC == 2) {
        const unsigned a = (unsigned)f & 3u;
        if (a == 1u || (a == 0u && row + 1 < rows)) {          // the neighbour's float receives +0
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p - a), "f"(a ? 0.0f : x), "f"(a ? x : y),
                         "f"(a ? y : z), "f"(a ? z : 0.0f) : "memory");
            return;
        }
    }
    if (VEC >= 1) {
        const bool odd = f & 1;       // 8-byte aligned at the row's start (even offset) or one float in (odd offset)
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p + (odd ? 1 : 0)), "f"(odd ? y : x), "f"(odd ? z : y) : "memory");
        atomicAdd(p + (odd ? 0 : 2), odd ? x : z);
        return;
    }
#endif
    atomicAdd(p, x); atomicAdd(p + 1, y); atomicAdd(p + 2, z);
}

// One thread per point: u = w/d * (own - partner) is added to the point's own gradient row and subtracted from its
// partner's row with float reductions at L2 on zero-filled (or to-be-accumulated) outputs.  A warp's 32 own rows are 384
// contiguous bytes: when they start on a 16-byte boundary (VEC 2, cloud sizes that are multiples of 4) they leave as 24
// four-float reductions, transposed through shared memory.  Latency bound (two dependent memory round trips per point:
// index, then partner): 32 registers for 8 CTAs per SM.
static constexpr int kBwdThreads = 256;

// Grid of every backward kernel: x = blocks over the N + M points of one pair, y = pair (B <= 65535 like the forward) -- no
// integer division per thread (a flat 64-bit index costs two emulated 64-bit divisions, ~45 % of the kernel's instructions;
// the kernel is latency bound, so this only shortens the instruction stream: 306 -> 228 per warp).
template <int VEC>
__global__ void __launch_bounds__(kBwdThreads, 8) chamfer_bwd_kernel(BwdArgs a, int B) {
    __shared__ __align__(16) float stage[kBwdThreads / 32][96];
    pdl_launch_dependents();
    pdl_wait();                       // distances, indices and the upstream gradient come from the kernels before
    const int per_cloud = a.N + a.M;
    const int lane = threadIdx.x & 31;
    const int p = blockIdx.x * kBwdThreads + threadIdx.x, b = blockIdx.y;
    const bool live = p < per_cloud;
    const int dir = p >= a.N, i = dir ? p - a.N : p;
    const int n = dir ? a.M : a.N, m = dir ? a.N : a.M;
    float ux = 0.0f, uy = 0.0f, uz = 0.0f;
    int j = 0;
    const bool has = live && bwd_term(a, dir, b, i, ux, uy, uz, j);
    const size_t row = (size_t)b * n + i;
    float *own = dir ? a.gpc2 : a.gpc1;
    bool staged = false;
#ifndef RLG_BWD_SCALAR_RED
    if (VEC == 2) {
        // the whole warp inside one cloud of the pair, first row on a 16-byte boundary (the same answer in every lane)
        const int p0 = p - lane;
        const size_t row0 = row - lane;
        staged = (p0 < a.N ? p0 + 31 < a.N : p0 + 31 < per_cloud) && (row0 & 3) == 0;
        if (staged && __any_sync(0xffffffffu, has)) {
            float *sw = stage[threadIdx.x >> 5];
            sw[3 * lane] = ux; sw[3 * lane + 1] = uy; sw[3 * lane + 2] = uz;
            __syncwarp();
            if (lane < 24) {
                const float4 v = reinterpret_cast<const float4 *>(sw)[lane];
                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(own + row0 * 3 + 4 * lane), "f"(v.x), "f"(v.y),
                             "f"(v.z), "f"(v.w) : "memory");
            }
        }
    }
#endif
    if (!has) return;
    if (!staged) red_add_row<VEC>(own, row, (size_t)B * n, ux, uy, uz);
    red_add_row<VEC>(dir ? a.gpc1 : a.gpc2, (size_t)b * m + j, (size_t)B * m, -ux, -uy, -uz);
}

// ---- reproducible variant: fixed-point scatter, then own term + conversion ---------------------------------------
struct DetArgs {
    long long *acc1, *acc2;   // (B,N,3), (B,M,3): partner sums of gpc1 / gpc2 rows, in quanta of the OTHER direction's weight
    int s1, s2;               // fraction bits of direction 0 (N queries) and 1 (M queries)
    int accumulate;
};

__device__ __forceinline__ bool det_usable(float w) { return w != 0.0f && isfinite(w); }
// 2^(s - ilogb|w|) (to quanta, sign +1) or its inverse (sign -1); w finite and non-zero
__device__ __forceinline__ double det_scale(float w, int s, int sign) {
    return scalbn(1.0, sign * (s - ilogb((double)fabsf(w))));
}

__global__ void __launch_bounds__(kBwdThreads) chamfer_bwd_scatter_kernel(BwdArgs a, DetArgs q, int B) {
    pdl_launch_dependents();
    pdl_wait();
    const int p = blockIdx.x * kBwdThreads + threadIdx.x, b = blockIdx.y;
    if (p >= a.N + a.M) return;
    const int dir = p >= a.N, i = dir ? p - a.N : p;
    float ux, uy, uz;
    int j;
    if (!bwd_term(a, dir, b, i, ux, uy, uz, j)) return;
    const float w = bwd_weight(a, dir, b);
    if (!det_usable(w)) return;       // non-finite weight: the gather kernel writes NaN rows
    const double to_q = det_scale(w, dir ? q.s2 : q.s1, +1);
    unsigned long long *oth = reinterpret_cast<unsigned long long *>(dir ? q.acc1 : q.acc2) + ((size_t)b * (dir ? a.N : a.M) + j) * 3;
    atomicAdd(oth, (unsigned long long)__double2ll_rn(-(double)ux * to_q));
    atomicAdd(oth + 1, (unsigned long long)__double2ll_rn(-(double)uy * to_q));
    atomicAdd(oth + 2, (unsigned long long)__double2ll_rn(-(double)uz * to_q));
}

__global__ void __launch_bounds__(kBwdThreads) chamfer_bwd_gather_kernel(BwdArgs a, DetArgs q, int B) {
    pdl_launch_dependents();
    pdl_wait();                       // the scatter kernel has finished (and flushed) when this returns
    const int p = blockIdx.x * kBwdThreads + threadIdx.x, b = blockIdx.y;
    if (p >= a.N + a.M) return;
    const int dir = p >= a.N, i = dir ? p - a.N : p;
    float ux, uy, uz;
    int j;
    bwd_term(a, dir, b, i, ux, uy, uz, j);                   // zeros where the term vanishes
    const size_t row = ((size_t)b * (dir ? a.M : a.N) + i) * 3;
    const float wo = bwd_weight(a, dir ^ 1, b);              // the partner terms of this row come from the other direction
    double px = 0.0, py = 0.0, pz = 0.0;
    if (det_usable(wo)) {
        const long long *acc = (dir ? q.acc2 : q.acc1) + row;
        const double from_q = det_scale(wo, dir ? q.s1 : q.s2, -1);
        px = (double)acc[0] * from_q; py = (double)acc[1] * from_q; pz = (double)acc[2] * from_q;
    } else if (wo != 0.0f) {
        px = py = pz = (double)NAN;
    }
    float *out = (dir ? a.gpc2 : a.gpc1) + row;
    const float rx = (float)((double)ux + px), ry = (float)((double)uy + py), rz = (float)((double)uz + pz);
    if (q.accumulate) { out[0] += rx; out[1] += ry; out[2] += rz; }
    else { out[0] = rx; out[1] = ry; out[2] = rz; }
}

}  // namespace rlg

using namespace rlg;

static int det_fraction_bits(int n) {       // min(40, 61 - ceil(log2 n)): n terms of < 2^(s+1) quanta each stay below 2^62
    int lg = 0;
    while ((1LL << lg) < (long long)n) ++lg;
    return 61 - lg < 40 ? 61 - lg : 40;
}

static int bwd_launch(const float *pc1, const float *pc2, const float *d1, const float *d2, const int32_t *i1,
                      const int32_t *i2, const float *g1, const float *g2, int gstride, float scale1, float scale2,
                      int B, int N, int M, float *gpc1, float *gpc2, unsigned flags, void *stream,
                      bool deterministic = false, void *ws = nullptr, size_t ws_bytes = 0) {
    if (B < 0 || N < 1 || M < 1)
        return fail(RLG_ERR_BAD_SHAPE, "rlg_chamfer_bwd: bad shape B=%d N=%d M=%d", B, N, M);
    if (B > 65535) return fail(RLG_ERR_TOO_LARGE, "rlg_chamfer_bwd: B=%d exceeds 65535 (grid.y)", B);
    if (flags & ~RLG_CHAMFER_BWD_ACCUMULATE)
        return fail(RLG_ERR_UNSUPPORTED, "rlg_chamfer_bwd: unknown flag bits 0x%x", flags & ~RLG_CHAMFER_BWD_ACCUMULATE);
    if (B == 0) return 0;
    if (!pc1 || !pc2 || !d1 || !d2 || !i1 || !i2 || !gpc1 || !gpc2)
        return fail(RLG_ERR_NULL_POINTER, "rlg_chamfer_bwd: null pointer");
    BwdArgs a{pc1, pc2, d1, d2, i1, i2, g1, g2, gstride, scale1, scale2, gpc1, gpc2, N, M};
    cudaStream_t st = (cudaStream_t)stream;
    if ((long long)N + M > 0x7fffffffLL - kBwdThreads) return fail(RLG_ERR_TOO_LARGE, "rlg_chamfer_bwd: too many points");
    const dim3 grid((unsigned)(((long long)N + M + kBwdThreads - 1) / kBwdThreads), (unsigned)B);
    if (deterministic) {
        const size_t need = rlg_chamfer_bwd_ws_bytes(B, N, M);
        if (!ws) return fail(RLG_ERR_NULL_POINTER, "rlg_chamfer_bwd_det: null workspace");
        if (ws_bytes < need) return fail(RLG_ERR_WORKSPACE, "rlg_chamfer_bwd_det: workspace %zu < %zu bytes", ws_bytes, need);
        if ((uintptr_t)ws & 15) return fail(RLG_ERR_WORKSPACE, "rlg_chamfer_bwd_det: workspace not 16-byte aligned");
        cudaError_t e = cudaMemsetAsync(ws, 0, need, st);
        if (e != cudaSuccess) { cudaGetLastError(); return fail((int)e, "rlg_chamfer_bwd_det: cudaMemsetAsync: %s", cudaGetErrorString(e)); }
        long long *acc1 = static_cast<long long *>(ws);
        DetArgs q{acc1, acc1 + 3 * (size_t)B * N, det_fraction_bits(N), det_fraction_bits(M),
                  (flags & RLG_CHAMFER_BWD_ACCUMULATE) ? 1 : 0};
        chamfer_bwd_scatter_kernel<<<grid, dim3(kBwdThreads), 0, st>>>(a, q, B);       // (plain launch: see below)
        e = cudaGetLastError();
        if (e != cudaSuccess) { cudaGetLastError(); return fail((int)e, "chamfer_bwd_scatter_kernel: %s", cudaGetErrorString(e)); }
        e = launch_pdl(chamfer_bwd_gather_kernel, grid, dim3(kBwdThreads), 0, st, a, q, B);
        if (e != cudaSuccess) { cudaGetLastError(); return fail((int)e, "chamfer_bwd_gather_kernel: %s", cudaGetErrorString(e)); }
        return check_launch("chamfer_bwd_gather_kernel");
    }
    if (!(flags & RLG_CHAMFER_BWD_ACCUMULATE)) {
        cudaError_t e = cudaMemsetAsync(gpc1, 0, sizeof(float) * 3 * (size_t)B * N, st);
        if (e == cudaSuccess) e = cudaMemsetAsync(gpc2, 0, sizeof(float) * 3 * (size_t)B * M, st);
        if (e != cudaSuccess) { cudaGetLastError(); return fail((int)e, "rlg_chamfer_bwd: cudaMemsetAsync: %s", cudaGetErrorString(e)); }
    }
    const uintptr_t align = (uintptr_t)gpc1 | (uintptr_t)gpc2;
    auto kernel = (align & 15) == 0 ? chamfer_bwd_kernel<2> : (align & 7) == 0 ? chamfer_bwd_kernel<1> : chamfer_bwd_kernel<0>;
    // A plain launch, not a programmatic dependent one: at 28 registers one CTA of this kernel fits beside the forward's
    // persistent CTA on every SM, and such early-resident CTAs cost far more than the launch gap they hide (measured:
    // cfg5 step 3.00 -> 2.82 ms, cfg2 step 50.5 -> 47.5 us without the attribute).
    kernel<<<grid, dim3(kBwdThreads), 0, st>>>(a, B);
    cudaError_t le = cudaGetLastError();
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

extern "C" size_t rlg_chamfer_bwd_ws_bytes(int B, int N, int M) {
    if (B <= 0 || N < 1 || M < 1) return 0;
    return 3 * sizeof(long long) * (size_t)B * ((size_t)N + (size_t)M);
}

extern "C" int rlg_chamfer_bwd_det(const float *pc1, const float *pc2, const float *d1, const float *d2,
                                   const int32_t *i1, const int32_t *i2, const float *g1, const float *g2, int B,
                                   int N, int M, float *gpc1, float *gpc2, void *ws, size_t ws_bytes, unsigned flags,
                                   void *stream) {
    return bwd_launch(pc1, pc2, d1, d2, i1, i2, g1, g2, 1, 1.0f, 1.0f, B, N, M, gpc1, gpc2, flags, stream, true, ws, ws_bytes);
}

extern "C" int rlg_chamfer_loss_bwd_det(const float *pc1, const float *pc2, const float *d1, const float *d2,
                                        const int32_t *i1, const int32_t *i2, const float *gloss, float w1, float w2,
                                        int B, int N, int M, float *gpc1, float *gpc2, void *ws, size_t ws_bytes,
                                        unsigned flags, void *stream) {
    if (!gloss) return fail(RLG_ERR_NULL_POINTER, "rlg_chamfer_loss_bwd_det: null upstream gradient");
    return bwd_launch(pc1, pc2, d1, d2, i1, i2, gloss, w2 != 0.0f ? gloss : nullptr, 0, w1, w2, B, N, M, gpc1, gpc2,
                      flags, stream, true, ws, ws_bytes);
}
