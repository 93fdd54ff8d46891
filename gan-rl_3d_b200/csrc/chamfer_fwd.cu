// chamfer_fwd.cu -- batched Chamfer distance forward (replaces utils/losses.py:29-37 of the reference).
//
// The reference materialises the (B,N,M) matrix of cdist and reduces it twice.  Here every squared
// distance t(i,j) is evaluated ONCE, in registers, and feeds both directions:
//
//   tile kernel   one warp owns 32*R rows of pc1 (R per lane, in registers) and sweeps 32-column
//                 groups of pc2 staged through its private shared-memory slice.  Distances use the
//                 packed fp32x2 pipe (FADD2/FMUL2/FFMA2: two IEEE operations per issue slot) in the
//                 reference's direct-mode operation order, so they are bit-identical to ATen's
//                 direct cdist.  Row minima stay in registers (FMNMX3 over column pairs); column
//                 minima are reduced over the lane's R rows (FMNMX3), then over the warp with one
//                 REDUX.MIN + ballot per column.  Inside the sweep only minimum VALUES are tracked;
//                 the argmin is kept at group granularity (which 32-column group / which lane's R-row
//                 group) and merged across warps with 64-bit atomicMin on (t_bits << 32 | group),
//                 whose ordering gives the lowest-index tie-break for free (t >= 0 orders as uint).
//   finalize      per point, re-evaluates the <=32 candidates of the winning group to recover the
//                 exact index, writes sqrtf(min t) and the index, restores the workspace to the
//                 all-ones pattern, and reduces the per-pair means in a fixed order (deterministic).
//
// Work is cut into (cloud, row block, 32-column group) units and the linear unit range is split evenly
// over all warps of a persistent grid (148 SMs x resident warps), so the tiny headline shape
// (B=32, N=M=2048: 16384 units) still balances to within one unit per warp.
#include "common.cuh"
#include <math.h>

namespace rlg {

static constexpr float kPad = 3.0e18f;     // sentinel coordinate for out-of-range rows/columns

// ------------------------------------------------------------------------------------------------
// simple cross-check path: one thread per query point, candidates tiled through shared memory
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) chamfer_simple_kernel(const float *__restrict__ q,
                                                            const float *__restrict__ c, int nq, int nc,
                                                            float *__restrict__ dist,
                                                            int32_t *__restrict__ idx) {
    __shared__ float tile[128 * 3];
    const int b = blockIdx.y;
    const int i = blockIdx.x * 128 + threadIdx.x;
    const float *qb = q + (size_t)b * nq * 3;
    const float *cb = c + (size_t)b * nc * 3;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (i < nq) { px = qb[3 * i]; py = qb[3 * i + 1]; pz = qb[3 * i + 2]; }
    float best = INFINITY;
    int bj = 0;
    for (int j0 = 0; j0 < nc; j0 += 128) {
        const int cnt = min(128, nc - j0);
        __syncthreads();
        for (int e = threadIdx.x; e < cnt * 3; e += 128) tile[e] = cb[(size_t)j0 * 3 + e];
        __syncthreads();
        for (int j = 0; j < cnt; ++j) {
            float t = sqdist(px, py, pz, tile[3 * j], tile[3 * j + 1], tile[3 * j + 2]);
            if (t < best) { best = t; bj = j0 + j; }
        }
    }
    if (i < nq) {
        // all-NaN row: nothing compared less than +inf; report the candidate-0 distance like torch.min
        if (best == INFINITY) best = sqdist(px, py, pz, cb[0], cb[1], cb[2]);
        dist[(size_t)b * nq + i] = sqrtf(best);
        idx[(size_t)b * nq + i] = bj;
    }
}

// deterministic per-pair mean: one CTA per (cloud, direction)
__global__ void __launch_bounds__(256) chamfer_mean_kernel(const float *__restrict__ d1,
                                                          const float *__restrict__ d2, int N, int M,
                                                          float *__restrict__ mean1,
                                                          float *__restrict__ mean2) {
    __shared__ double red[256];
    const int b = blockIdx.x, dir = blockIdx.y;
    const int n = dir ? M : N;
    const float *d = (dir ? d2 : d1) + (size_t)b * n;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) acc += (double)d[i];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) (dir ? mean2 : mean1)[b] = (float)(red[0] / (double)n);
}

// ------------------------------------------------------------------------------------------------
// tile kernel
// ------------------------------------------------------------------------------------------------
template <int R>
struct TileCfg {
    static constexpr int kWarps = 4;                 // warps per CTA (one per SM sub-partition)
    static constexpr int kRowsPerWarp = 32 * R;
    static constexpr int kStride = 44;               // floats between the x/y/z arrays of a staged group
    static constexpr int kBufFloats = 3 * kStride;   // one staged group
};

template <bool MIN3>
__device__ __forceinline__ float minacc(float acc, float a, float b) {
    // FMNMX3 saves an issue slot but measures ~1.7 cycles next to packed FMA traffic vs ~0.6 per FMNMX
    if (MIN3) return min3(acc, a, b);
    return min2(min2(acc, a), b);
}

template <int R, int OCC, bool MIN3>
__global__ void __launch_bounds__(TileCfg<R>::kWarps * 32, OCC)
chamfer_tile_kernel(const float *__restrict__ pc1, const float *__restrict__ pc2, int N, int M,
                    int n_rb, int n_cg, long long total_units, u64 *__restrict__ rowkey,
                    u64 *__restrict__ colkey) {
    using Cfg = TileCfg<R>;
    __shared__ __align__(16) float s_stage[Cfg::kWarps][2][Cfg::kBufFloats];

    const int lane = threadIdx.x & 31;
    const int warp_in_cta = threadIdx.x >> 5;
    const long long warp = (long long)blockIdx.x * Cfg::kWarps + warp_in_cta;
    const long long n_warps = (long long)gridDim.x * Cfg::kWarps;
    const long long u0 = total_units * warp / n_warps;
    const long long u1 = total_units * (warp + 1) / n_warps;
    if (u0 >= u1) return;

    float x0[R], x1[R], x2[R];     // this lane's R rows (contiguous rows lane*R .. lane*R+R-1 of the block)
    float best[R];                 // running min t per row over all groups seen for the current row block
    int bgrp[R];                   // earliest column group attaining it
    long long cur_brb = -1;
    int b = 0, rb = 0;

    // staging element map: lane handles flat floats lane, lane+32, lane+64 of the 96-float group
    int st_off[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const int e = lane + 32 * k;
        st_off[k] = (e % 3) * Cfg::kStride + e / 3;
    }
    float pre[3];
    auto prefetch = [&](long long u) {
        const int cg = (int)(u % n_cg);
        const long long brb = u / n_cg;
        const int bb = (int)(brb / n_rb);
        const float *src = pc2 + (size_t)bb * M * 3;
        const int base = cg * (kGroup * 3);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int f = base + lane + 32 * k;
            pre[k] = (f < M * 3) ? __ldg(src + f) : kPad;
        }
    };
    auto flush_rows = [&]() {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = rb * Cfg::kRowsPerWarp + lane * R + r;
            if (i < N) {
                const u64 key = ((u64)__float_as_uint(best[r]) << 32) | (unsigned)bgrp[r];
                atomicMin(&rowkey[(size_t)b * N + i], key);
            }
        }
    };

    prefetch(u0);
    int buf = 0;
    for (long long u = u0; u < u1; ++u) {
        const int cg = (int)(u % n_cg);
        const long long brb = u / n_cg;
        if (brb != cur_brb) {
            if (cur_brb >= 0) flush_rows();
            cur_brb = brb;
            b = (int)(brb / n_rb);
            rb = (int)(brb % n_rb);
            const float *src = pc1 + (size_t)b * N * 3;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = rb * Cfg::kRowsPerWarp + lane * R + r;
                if (i < N) {
                    x0[r] = __ldg(src + 3 * i);
                    x1[r] = __ldg(src + 3 * i + 1);
                    x2[r] = __ldg(src + 3 * i + 2);
                } else {
                    x0[r] = x1[r] = x2[r] = -kPad;
                }
                best[r] = INFINITY;
                bgrp[r] = 0;
            }
        }
        // stage this group (fetched during the previous iteration), start fetching the next one
        float *stage = s_stage[warp_in_cta][buf];
#pragma unroll
        for (int k = 0; k < 3; ++k) stage[st_off[k]] = pre[k];
        __syncwarp();
        if (u + 1 < u1) prefetch(u + 1);

        float rowmin[R];
#pragma unroll
        for (int r = 0; r < R; ++r) rowmin[r] = INFINITY;
        unsigned keep_m = 0x7f800000u, keep_b = 1u;

#pragma unroll 1
        for (int step = 0; step < kGroup / 4; ++step) {
            const float4 X = *reinterpret_cast<const float4 *>(stage + 4 * step);
            const float4 Y = *reinterpret_cast<const float4 *>(stage + Cfg::kStride + 4 * step);
            const float4 Z = *reinterpret_cast<const float4 *>(stage + 2 * Cfg::kStride + 4 * step);
            const u64 Xa = pack2(X.x, X.y), Xb = pack2(X.z, X.w);
            const u64 Ya = pack2(Y.x, Y.y), Yb = pack2(Y.z, Y.w);
            const u64 Za = pack2(Z.x, Z.y), Zb = pack2(Z.z, Z.w);
            float c0 = INFINITY, c1 = INFINITY, c2 = INFINITY, c3 = INFINITY;
#pragma unroll
            for (int r = 0; r < R; r += 2) {
                float t[2][4];
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const u64 px = pack2(x0[r + rr], x0[r + rr]);
                    const u64 py = pack2(x1[r + rr], x1[r + rr]);
                    const u64 pz = pack2(x2[r + rr], x2[r + rr]);
                    u64 d0 = sub2(px, Xa), d1 = sub2(py, Ya), d2 = sub2(pz, Za);
                    u64 ta = mul2(d0, d0);
                    ta = fma2(d1, d1, ta);
                    ta = fma2(d2, d2, ta);
                    d0 = sub2(px, Xb); d1 = sub2(py, Yb); d2 = sub2(pz, Zb);
                    u64 tb = mul2(d0, d0);
                    tb = fma2(d1, d1, tb);
                    tb = fma2(d2, d2, tb);
                    unpack2(ta, t[rr][0], t[rr][1]);
                    unpack2(tb, t[rr][2], t[rr][3]);
                    rowmin[r + rr] = minacc<MIN3>(rowmin[r + rr], t[rr][0], t[rr][1]);
                    rowmin[r + rr] = minacc<MIN3>(rowmin[r + rr], t[rr][2], t[rr][3]);
                }
                c0 = minacc<MIN3>(c0, t[0][0], t[1][0]);
                c1 = minacc<MIN3>(c1, t[0][1], t[1][1]);
                c2 = minacc<MIN3>(c2, t[0][2], t[1][2]);
                c3 = minacc<MIN3>(c3, t[0][3], t[1][3]);
            }
            // warp-wide column minima; lane (4*step+q) keeps the result of column q of this step
            const float cq[4] = {c0, c1, c2, c3};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const unsigned mine = __float_as_uint(cq[q]);
                const unsigned m = __reduce_min_sync(0xffffffffu, mine);
                const unsigned bal = __ballot_sync(0xffffffffu, mine == m);
                if (lane == 4 * step + q) { keep_m = m; keep_b = bal; }
            }
        }
        // group epilogue: rows (strict < keeps the earliest group on ties)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (rowmin[r] < best[r]) { best[r] = rowmin[r]; bgrp[r] = cg; }
        }
        // columns: lane owns column cg*32+lane; the lowest lane attaining the min = lowest row group
        {
            const int j = cg * kGroup + lane;
            if (j < M) {
                const unsigned src_lane = (unsigned)(__ffs((int)keep_b) - 1);
                const unsigned rowgroup = (unsigned)rb * 32u + src_lane;
                const u64 key = ((u64)keep_m << 32) | rowgroup;
                atomicMin(&colkey[(size_t)b * M + j], key);
            }
        }
        buf ^= 1;
    }
    flush_rows();
}

template <int R, int OCC, bool MIN3>
static int launch_tile(const float *pc1, const float *pc2, int B, int N, int M, u64 *rowkey, u64 *colkey,
                       cudaStream_t st) {
    using Cfg = TileCfg<R>;
    const int n_rb = (N + Cfg::kRowsPerWarp - 1) / Cfg::kRowsPerWarp;
    const int n_cg = (M + kGroup - 1) / kGroup;
    const long long total = (long long)B * n_rb * n_cg;
    const int sms = sm_count();
    if (sms <= 0) return fail((int)cudaErrorNoDevice, "rlg_chamfer_fwd: no CUDA device");
    int ctas_per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, chamfer_tile_kernel<R, OCC, MIN3>,
                                                                  Cfg::kWarps * 32, 0);
    if (e != cudaSuccess || ctas_per_sm < 1) ctas_per_sm = 1;
    long long grid = (long long)sms * ctas_per_sm;
    const long long max_useful = (total + Cfg::kWarps - 1) / Cfg::kWarps;   // at least one unit per warp
    if (grid > max_useful) grid = max_useful;
    if (grid < 1) grid = 1;
    chamfer_tile_kernel<R, OCC, MIN3><<<(unsigned)grid, Cfg::kWarps * 32, 0, st>>>(pc1, pc2, N, M, n_rb, n_cg, total,
                                                                       rowkey, colkey);
    return check_launch("chamfer_tile_kernel");
}

}  // namespace rlg

using namespace rlg;

extern "C" {

static size_t keys_bytes(int B, int N, int M) { return align_up(sizeof(u64) * ((size_t)B * N + (size_t)B * M), 256); }
static size_t secs_bytes(int B, int N, int M) { return align_up(sizeof(unsigned) * ((size_t)B * N + (size_t)B * M), 256); }
static size_t nrm_bytes(int B) { return align_up(sizeof(unsigned) * 2 * (size_t)B, 256); }

size_t rlg_chamfer_ws_bytes(int B, int N, int M) {
    if (B < 0 || N < 1 || M < 1) return 0;
    // packed (value, group) keys + second-best values + cloud norms + finalize counters/partials (the direct
    // cross-check path uses a prefix of the same layout)
    const size_t fin = finalize2_ws_bytes(B, N, M) > finalize_ws_bytes(B, N, M) ? finalize2_ws_bytes(B, N, M)
                                                                                : finalize_ws_bytes(B, N, M);
    // + second group / third value of the tensor-core sweep (two more arrays shaped like the second-best values)
    return keys_bytes(B, N, M) + secs_bytes(B, N, M) + nrm_bytes(B) + fin + 2 * secs_bytes(B, N, M);
}

int rlg_chamfer_fwd(const float *pc1, const float *pc2, int B, int N, int M, float *d1, float *d2,
                    int32_t *i1, int32_t *i2, float *mean1, float *mean2, void *ws, size_t ws_bytes,
                    unsigned flags, void *stream) {
    return rlg_chamfer_loss_fwd(pc1, pc2, B, N, M, d1, d2, i1, i2, mean1, mean2, nullptr, 0.0f, 0.0f, nullptr, nullptr,
                                ws, ws_bytes, flags, stream);
}

int rlg_chamfer_loss_fwd(const float *pc1, const float *pc2, int B, int N, int M, float *d1, float *d2,
                         int32_t *i1, int32_t *i2, float *mean1, float *mean2, float *loss, float w1, float w2,
                         float *gz1, float *gz2, void *ws, size_t ws_bytes, unsigned flags, void *stream) {
    if (B < 0 || N < 1 || M < 1)
        return fail(RLG_ERR_BAD_SHAPE, "rlg_chamfer_fwd: bad shape B=%d N=%d M=%d (need B>=0, N>=1, M>=1)", B, N, M);
    if (B == 0) return 0;
    if (!pc1 || !pc2 || !d1 || !d2 || !i1 || !i2) return fail(RLG_ERR_NULL_POINTER, "rlg_chamfer_fwd: null pointer");
    if ((mean1 == nullptr) != (mean2 == nullptr))
        return fail(RLG_ERR_NULL_POINTER, "rlg_chamfer_fwd: mean1/mean2 must both be given or both be null");
    if (loss != nullptr && mean1 == nullptr)
        return fail(RLG_ERR_NULL_POINTER, "rlg_chamfer_loss_fwd: the batch loss needs the mean1/mean2 buffers");
    if ((long long)B > 65535) return fail(RLG_ERR_TOO_LARGE, "rlg_chamfer_fwd: B=%d exceeds 65535 (grid.y of the finalize)", B);
    if ((long long)N * 3 > 0x7fffffffLL || (long long)M * 3 > 0x7fffffffLL)
        return fail(RLG_ERR_TOO_LARGE, "rlg_chamfer_fwd: N or M too large for 32-bit indexing");
    cudaStream_t st = (cudaStream_t)stream;
    const bool fused_zero = !(flags & (RLG_CHAMFER_ALGO_SIMPLE | RLG_CHAMFER_ALGO_DIRECT | RLG_CHAMFER_TILE_ONLY));
    if (!fused_zero) {              // the cross-check paths have no fused zero-fill: plain memsets
        if (gz1) cudaMemsetAsync(gz1, 0, sizeof(float) * 3 * (size_t)B * N, st);
        if (gz2) cudaMemsetAsync(gz2, 0, sizeof(float) * 3 * (size_t)B * M, st);
    }

    if (flags & RLG_CHAMFER_ALGO_SIMPLE) {
        dim3 g1((N + 127) / 128, B), g2((M + 127) / 128, B);
        chamfer_simple_kernel<<<g1, 128, 0, st>>>(pc1, pc2, N, M, d1, i1);
        chamfer_simple_kernel<<<g2, 128, 0, st>>>(pc2, pc1, M, N, d2, i2);
        if (mean1) chamfer_mean_kernel<<<dim3(B, 2), 256, 0, st>>>(d1, d2, N, M, mean1, mean2);
        if (loss) return fail(RLG_ERR_UNSUPPORTED, "rlg_chamfer_loss_fwd: the simple cross-check path has no fused loss");
        return check_launch("chamfer_simple_kernel");
    }

    const size_t need = rlg_chamfer_ws_bytes(B, N, M);
    if (!ws || ws_bytes < need || ((uintptr_t)ws & 255u))
        return fail(RLG_ERR_WORKSPACE, "rlg_chamfer_fwd: workspace %p/%zu bytes, need %zu bytes 256-B aligned", ws,
                    ws_bytes, need);
    if (!(flags & RLG_CHAMFER_WS_CLEAN)) {
        cudaError_t e = cudaMemsetAsync(ws, 0xff, need, st);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail((int)e, "rlg_chamfer_fwd: cudaMemsetAsync: %s", cudaGetErrorString(e));
        }
    }
    u64 *rowkey = (u64 *)ws;
    u64 *colkey = rowkey + (size_t)B * N;
    const unsigned variant = (flags >> 8) & 15u;      // experimental kernel variants (tools/sweep_tile.py); 0 = production
    if (flags & RLG_CHAMFER_ALGO_DIRECT) {
        // direct-form tile kernel: every pair evaluated exactly (6 FP32 operations); kept as a cross-check
        int R = 8, rc = 0;
        switch (variant) {
            default:
            case 0: rc = launch_tile<8, 3, false>(pc1, pc2, B, N, M, rowkey, colkey, st); break;
            case 1: rc = launch_tile<8, 3, true>(pc1, pc2, B, N, M, rowkey, colkey, st); break;
            case 2: rc = launch_tile<8, 4, false>(pc1, pc2, B, N, M, rowkey, colkey, st); break;
            case 4: R = 16; rc = launch_tile<16, 2, false>(pc1, pc2, B, N, M, rowkey, colkey, st); break;
        }
        if (rc) return rc;
        if (flags & RLG_CHAMFER_TILE_ONLY) return 0;
        char *fin = (char *)ws + keys_bytes(B, N, M) + secs_bytes(B, N, M) + nrm_bytes(B);
        return launch_finalize(pc1, pc2, B, N, M, R, rowkey, colkey, fin, d1, d2, i1, i2, mean1, mean2, loss, w1, w2, st);
    }
    FwdWs w;
    w.rowkey = rowkey;
    w.colkey = colkey;
    w.rowsec = (unsigned *)((char *)ws + keys_bytes(B, N, M));
    w.colsec = w.rowsec + (size_t)B * N;
    w.nrm = (unsigned *)((char *)ws + keys_bytes(B, N, M) + secs_bytes(B, N, M));
    w.rowsg = (unsigned *)((char *)ws + need - 2 * secs_bytes(B, N, M));
    w.colsg = w.rowsg + (size_t)B * N;
    w.rowth = (unsigned *)((char *)ws + need - secs_bytes(B, N, M));
    w.colth = w.rowth + (size_t)B * N;
    int R = 0;
    int rc = (flags & RLG_CHAMFER_ALGO_TENSOR) ? launch_tcfilter(pc1, pc2, B, N, M, w, &R, st)
                                                : launch_filter(pc1, pc2, B, N, M, (int)variant, w, &R, st);
    if (rc) return rc;
    if (flags & RLG_CHAMFER_TILE_ONLY) return 0;
    char *fin = (char *)ws + keys_bytes(B, N, M) + secs_bytes(B, N, M) + nrm_bytes(B);
    return launch_finalize2(pc1, pc2, B, N, M, R, w, fin, d1, d2, i1, i2, mean1, mean2, loss, w1, w2, gz1, gz2, st);
}

}  // extern "C"
