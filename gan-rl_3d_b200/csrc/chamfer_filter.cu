// chamfer_filter.cu -- batched Chamfer distance forward, filter-and-refine (replaces utils/losses.py:29-37).
//
// The reference materialises the (B,N,M) matrix of torch.cdist and reduces it twice.  Here every pair (i,j) is
// visited ONCE by a cheap FP32 filter that serves both directions, and the exact (direct-difference, bit-identical
// to ATen's direct cdist) arithmetic is spent only on the few candidates that can still win:
//
//   filter kernel   g(i,j) = |y_j|^2 - 2 x_i.y_j   as a 3-FFMA chain seeded with |y_j|^2   (row direction: min_j g)
//                   t(i,j) = g(i,j) + |x_i|^2      one FADD                                (column direction: min_i t)
//                   = 4 FP32-pipe operations per pair instead of the 6 of the direct form, issued as packed
//                   fp32x2 instructions (FFMA2/FADD2: two rows per issue slot).  A warp owns 32*R rows of pc1
//                   (R per lane, in registers, as row PAIRS so no operand needs duplicating) and sweeps
//                   32-column groups of pc2 that it stages itself in shared memory as pre-scaled, pre-duplicated
//                   operands (-2y0,-2y0,-2y1,-2y1 | -2y2,-2y2,|y|^2,|y|^2: two broadcast LDS.128 per column).
//                   Per row it keeps the best and SECOND-best group minimum and the best group; per column it
//                   reduces over the lane's R rows and parks (value bits & ~31 | lane) in shared memory -- the sweep
//                   itself has no cross-lane instruction -- and after the group each lane scans one column's 32
//                   entries for the best and second best.  Results are merged across warps with 64-bit atomicMin on
//                   (value bits << 32 | group); whatever loses a merge is pushed into a second-best array.
//   finalize        per query point: if the second-best group is farther than best + margin (a rigorous bound on
//                   filter-vs-direct rounding, below), the exact winner must lie in the best group: evaluate its
//                   <= 32 candidates in the direct form.  Otherwise (0.5 % of the points) scan all candidates
//                   exactly.  Either way the outputs are sqrtf(min t) and the lowest index attaining it -- the
//                   same bits the direct-form kernel produces -- independent of how the filter rounded.
//
// Error bound (u = 2^-24, a = |x_i|, b = |y_j|, exact t = |x_i - y_j|^2, fp32 inputs taken as exact):
//   filter   |g~ - (t - a^2)| <= 3u b^2 + 3u (a+b)^2          (|y|^2: 3 roundings; chain: 3 roundings of partial
//                                                              sums that never exceed (a+b)^2)
//   t~ = fl(g~ + nx):  adds 3u a^2 (nx) + u (a+b)^2            => |t~ - t| <= 10u (a+b)^2
//   direct   |t^ - t| <= 5u t <= 5u (a+b)^2                   (d_k = fl(x_k - y_k), square, two FMAs)
//   => |filter - direct| <= 15u (a+b)^2 <= 30u (a^2 + b^2).  Two candidates can swap order only if their filter
//   values differ by <= 60u (a^2 + b^2); kMarginC = 96u adds slack for the fmaxf(.,0)/rounding of the merged keys
//   and for using computed norms.  Column keys drop 5 mantissa bits for the lane id: +2^-17 relative (kMarginQ).
//
// Work is cut into (cloud, row block, 32-column group) units; the linear unit range is split evenly over all warps
// of a persistent grid, so the headline shape (B=32, N=M=2048) balances to within one unit in seven per warp.
#include "common.cuh"
#include <math.h>

namespace rlg {

static constexpr float kBig = 1.0e30f;                 // |x|^2 or |y|^2 of a padding row / column: never wins
static constexpr int kFWarps = 4;                      // warps per CTA; each warp works alone (no __syncthreads)
static constexpr float kMarginC = 96.0f / 16777216.0f; // 96 u
static constexpr float kMarginQ = 1.0f / 65536.0f;     // column keys: 2 x 32 ulp of quantisation, with slack
static constexpr float kMarginT = 128.0f / 16777216.0f; // 128 u: the tensor-core filter (chamfer_tcfilter.cu)

template <bool MIN3>
__device__ __forceinline__ float facc(float acc, float a, float b) {
    if (MIN3) return min3(acc, a, b);
    return min2(min2(acc, a), b);
}

// |p|^2 in a fixed operation order (used by the filter and by the finalize's margin; not part of the outputs)
__device__ __forceinline__ float norm2(float x, float y, float z) { return fmaf(z, z, fmaf(y, y, x * x)); }

// KO: timing experiments only (tools/sweep_tile.py variants 10..15; results are WRONG with bits 0,1,3 set):
//   1 skip the column scan, 2 skip the row second-best bookkeeping, 4 integer min3 on t for both directions,
//   8 stage the columns only for a warp's first unit
template <int R, int OCC, bool PP, int KO = 0>   // rows per lane, CTAs per SM the register budget is cut for, operand ping-pong
__global__ void __launch_bounds__(kFWarps * 32, OCC)
chamfer_filter_kernel(const float *__restrict__ pc1, const float *__restrict__ pc2, int B, int N, int M, int n_rb,
                      int n_cg, long long total_units, FwdWs w) {
    constexpr int P = R / 2;                      // row pairs per lane
    constexpr int kRows = 32 * R;                 // rows per warp row block; lane owns rows r*32 + lane (coalesced)
    __shared__ __align__(16) ulonglong2 s_col[kFWarps][32][2];
    __shared__ int s_cmin[kFWarps][32][32];       // [column][lane]: per-lane column minima of the current group
    __shared__ float s_second[kFWarps][R][32];    // [row][lane]: second-smallest group minimum of g per row
    __shared__ int s_bgrp[kFWarps][R][32];        // [row][lane]: earliest group attaining the smallest

    const int lane = threadIdx.x & 31;
    const int wic = threadIdx.x >> 5;
    // even split of the unit range over all warps: the first `rem` warps take one unit more
    const long long warp = (long long)blockIdx.x * kFWarps + wic;
    const long long n_warps = (long long)gridDim.x * kFWarps;
    const long long per = total_units / n_warps, rem = total_units - per * n_warps;
    const long long u0 = warp * per + (warp < rem ? warp : rem);
    const long long u1 = u0 + per + (warp < rem ? 1 : 0);
    pdl_launch_dependents();
    pdl_wait();                       // the clouds may come from the kernel right before this one; the workspace does
    if (u0 >= u1) return;
    const int n_units = (int)(u1 - u0);

    u64 xp0[P], xp1[P], xp2[P], nxp[P];           // rows (2p, 2p+1)*32+lane of the block, packed; |x|^2 likewise
    float best[R];                                // smallest group minimum of g per row (second-best and the group id
                                                  // are touched once per group: they live in shared memory)
    // unit = (cloud b, row block rb, column group cg), decoded once and then advanced incrementally
    int cg = (int)(u0 % n_cg);
    int rb = (int)((u0 / n_cg) % n_rb);
    int b = (int)(u0 / n_cg / n_rb);
    bool new_block = true;
    int ymax_b = -1;
    unsigned ymax_bits = 0;                       // largest |y|^2 this warp has published for cloud ymax_b

    // column results of the previous unit: merged into the workspace at the top of the next unit, the returned
    // previous key consumed at its bottom (one sweep later), so the atomic's latency is never waited for
    bool prev = false;
    u64 prev_key = 0;
    unsigned prev_sec = 0;
    size_t prev_idx = 0;

    float pre[3];
    auto prefetch = [&](int bb, int cgg) {
        const int j = min(cgg * kGroup + lane, M - 1);
        const float *src = pc2 + ((size_t)bb * M + j) * 3;
        pre[0] = __ldg(src); pre[1] = __ldg(src + 1); pre[2] = __ldg(src + 2);
    };
    auto merge_second = [&](unsigned *secs, size_t idx, u64 old, u64 key, unsigned sec) {
        const u64 loser = old > key ? old : key;          // the key that did not (or no longer does) hold the slot
        unsigned s = (unsigned)(loser >> 32);
        s = s < sec ? s : sec;
        if (s != 0xffffffffu) atomicMin(secs + idx, s);
    };
    auto flush_rows = [&](int fb, int frb) {
        u64 key[R], old[R];
        unsigned sec[R];
#pragma unroll
        for (int p = 0; p < P; ++p) {
            // t = g + |x|^2 for both rows of the pair at once (nxp is only ever used as a packed operand)
            const u64 rowoff = (KO & 4) ? pack2(0.0f, 0.0f) : nxp[p];       // KO&4 keeps the rows in t already
            const u64 tb2 = add2(pack2(best[2 * p], best[2 * p + 1]), rowoff);
            const u64 ts2 = add2(pack2(s_second[wic][2 * p][lane], s_second[wic][2 * p + 1][lane]), rowoff);
            float tb[2], ts[2];
            unpack2(tb2, tb[0], tb[1]);
            unpack2(ts2, ts[0], ts[1]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = 2 * p + h;
                const int i = frb * kRows + r * 32 + lane;
                key[r] = ((u64)__float_as_uint(fmaxf(tb[h], 0.0f)) << 32) | (unsigned)s_bgrp[wic][r][lane];
                sec[r] = __float_as_uint(fmaxf(ts[h], 0.0f));
                old[r] = kKeyInit;
                if (i < N) old[r] = atomicMin(&w.rowkey[(size_t)fb * N + i], key[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = frb * kRows + r * 32 + lane;
            if (i < N) merge_second(w.rowsec, (size_t)fb * N + i, old[r], key[r], sec[r]);
        }
    };

    prefetch(b, cg);
#pragma unroll 1
    for (int it = 0; it < n_units; ++it) {
        if (new_block) {
            new_block = false;
            const float *src = pc1 + (size_t)b * N * 3;
            float xmax = 0.0f;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                float c[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
                float pad[2] = {kBig, kBig};
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i = rb * kRows + (2 * p + h) * 32 + lane;
                    if (i < N) {
                        c[h][0] = __ldg(src + 3 * (size_t)i);
                        c[h][1] = __ldg(src + 3 * (size_t)i + 1);
                        c[h][2] = __ldg(src + 3 * (size_t)i + 2);
                        pad[h] = 0.0f;
                    }
                }
                xp0[p] = pack2(c[0][0], c[1][0]);
                xp1[p] = pack2(c[0][1], c[1][1]);
                xp2[p] = pack2(c[0][2], c[1][2]);
                // |x|^2 of both rows with packed operations, so the pair is born (and stays) in one 64-bit register
                const u64 nn = fma2(xp2[p], xp2[p], fma2(xp1[p], xp1[p], mul2(xp0[p], xp0[p])));
                float n0, n1;
                unpack2(nn, n0, n1);
                xmax = fmaxf(xmax, fmaxf(n0, n1));                   // padding rows are all-zero: they add nothing
                nxp[p] = add2(nn, pack2(pad[0], pad[1]));
            }
#pragma unroll
            for (int r = 0; r < R; ++r) { best[r] = INFINITY; s_second[wic][r][lane] = INFINITY; s_bgrp[wic][r][lane] = 0; }
            const unsigned wx = __reduce_max_sync(0xffffffffu, __float_as_uint(xmax));
            if (lane == 0) atomicMin(&w.nrm[b], ~wx);
        }

        // ---- stage this group's 32 columns as FFMA2 operands; start fetching the next group
        if (!(KO & 8) || it == 0) {
            const bool cvalid = cg * kGroup + lane < M;
            const float y0 = pre[0], y1 = pre[1], y2 = pre[2];
            const float ny = cvalid ? norm2(y0, y1, y2) : kBig;
            const float s = cvalid ? -2.0f : 0.0f;
            const float Y0 = s * y0, Y1 = s * y1, Y2 = s * y2;
            s_col[wic][lane][0] = make_ulonglong2(pack2(Y0, Y0), pack2(Y1, Y1));
            s_col[wic][lane][1] = make_ulonglong2(pack2(Y2, Y2), pack2(ny, ny));
            const unsigned wy = __reduce_max_sync(0xffffffffu, cvalid ? __float_as_uint(ny) : 0u);
            if (b != ymax_b) { ymax_b = b; ymax_bits = 0; }
            if (wy > ymax_bits) {
                ymax_bits = wy;
                if (lane == 0) atomicMin(&w.nrm[B + b], ~wy);
            }
        }
        __syncwarp();
        // next unit's coordinates (incremental decode) and its column prefetch
        int ncg = cg + 1, nrb = rb, nb = b;
        if (ncg == n_cg) { ncg = 0; if (++nrb == n_rb) { nrb = 0; ++nb; } }
        if (it + 1 < n_units) prefetch(nb, ncg);
        // previous unit's column merge goes out now; its answer is needed only after this unit's sweep
        u64 prev_old = kKeyInit;
        if (prev) prev_old = atomicMin(&w.colkey[prev_idx], prev_key);

        float rowmin[R];
#pragma unroll
        for (int r = 0; r < R; ++r) rowmin[r] = INFINITY;

        // Column pairs, two per iteration with ping-pong operand registers: the operands of the next pair are
        // fetched while the current pair is computed.  Per-lane column minima go to shared memory as
        // (value bits & ~31 | lane) -- no cross-lane instruction inside the loop.
        auto sweep_pair = [&](const ulonglong2 &a0, const ulonglong2 &a1, const ulonglong2 &b0, const ulonglong2 &b1,
                              int cp) {
            float cA = INFINITY, cB = INFINITY;
#pragma unroll
            for (int p = 0; p < P; ++p) {
                const u64 gA = fma2(xp0[p], a0.x, fma2(xp1[p], a0.y, fma2(xp2[p], a1.x, a1.y)));
                const u64 gB = fma2(xp0[p], b0.x, fma2(xp1[p], b0.y, fma2(xp2[p], b1.x, b1.y)));
                const u64 tA = add2(gA, nxp[p]);
                const u64 tB = add2(gB, nxp[p]);
                float gAl, gAh, gBl, gBh, tAl, tAh, tBl, tBh;
                unpack2(gA, gAl, gAh); unpack2(gB, gBl, gBh);
                unpack2(tA, tAl, tAh); unpack2(tB, tBl, tBh);
                if (KO & 4) {
                    rowmin[2 * p] = __int_as_float(__vimin3_s32(__float_as_int(rowmin[2 * p]), __float_as_int(tAl), __float_as_int(tBl)));
                    rowmin[2 * p + 1] = __int_as_float(__vimin3_s32(__float_as_int(rowmin[2 * p + 1]), __float_as_int(tAh), __float_as_int(tBh)));
                    cA = __int_as_float(__vimin3_s32(__float_as_int(cA), __float_as_int(tAl), __float_as_int(tAh)));
                    cB = __int_as_float(__vimin3_s32(__float_as_int(cB), __float_as_int(tBl), __float_as_int(tBh)));
                } else {
                rowmin[2 * p] = facc<true>(rowmin[2 * p], gAl, gBl);
                rowmin[2 * p + 1] = facc<true>(rowmin[2 * p + 1], gAh, gBh);
                cA = facc<true>(cA, tAl, tAh);
                cB = facc<true>(cB, tBl, tBh);
                }
            }
            s_cmin[wic][2 * cp][lane] = (__float_as_int(cA) & ~31) | lane;
            s_cmin[wic][2 * cp + 1][lane] = (__float_as_int(cB) & ~31) | lane;
        };
        if (PP) {
            ulonglong2 pa0 = s_col[wic][0][0], pa1 = s_col[wic][0][1], pb0 = s_col[wic][1][0], pb1 = s_col[wic][1][1];
#pragma unroll 1
            for (int cp = 0; cp < kGroup / 2; cp += 2) {
                const ulonglong2 qa0 = s_col[wic][2 * cp + 2][0], qa1 = s_col[wic][2 * cp + 2][1];
                const ulonglong2 qb0 = s_col[wic][2 * cp + 3][0], qb1 = s_col[wic][2 * cp + 3][1];
                sweep_pair(pa0, pa1, pb0, pb1, cp);
                const int nx = cp + 2 < kGroup / 2 ? 2 * cp + 4 : 0;      // the last prefetch re-reads pair 0 (unused)
                pa0 = s_col[wic][nx][0]; pa1 = s_col[wic][nx][1];
                pb0 = s_col[wic][nx + 1][0]; pb1 = s_col[wic][nx + 1][1];
                sweep_pair(qa0, qa1, qb0, qb1, cp + 1);
            }
        } else {
#pragma unroll 1
            for (int cp = 0; cp < kGroup / 2; ++cp)
                sweep_pair(s_col[wic][2 * cp][0], s_col[wic][2 * cp][1], s_col[wic][2 * cp + 1][0], s_col[wic][2 * cp + 1][1], cp);
        }
        __syncwarp();
        // lane j reduces column j over the 32 lanes' entries (rotated start: conflict-free): smallest key (its low
        // 5 bits name the lane, i.e. which rows r*32+lane) and the runner-up.  Signed order: a rounding residue below
        // zero sorts first, which is what we want.
        // four independent (best, runner-up) chains of 8 entries each, then a 2-level merge: short dependency chains
        int bst[4], snd[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) { bst[c] = 0x7fffffff; snd[c] = 0x7fffffff; }
        if (KO & 1) { bst[0] = s_cmin[wic][lane][lane]; snd[0] = s_cmin[wic][lane][(lane + 1) & 31]; }
#pragma unroll
        for (int l = 0; l < ((KO & 1) ? 0 : 8); ++l) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int k = s_cmin[wic][lane][(8 * c + l + lane) & 31];
                snd[c] = min(snd[c], max(bst[c], k));
                bst[c] = min(bst[c], k);
            }
        }
        auto merge2 = [](int &b0, int &s0, int b1, int s1) {
            s0 = min(max(b0, b1), min(s0, s1));
            b0 = min(b0, b1);
        };
        merge2(bst[0], snd[0], bst[1], snd[1]);
        merge2(bst[2], snd[2], bst[3], snd[3]);
        merge2(bst[0], snd[0], bst[2], snd[2]);
        const int2 mm = make_int2(bst[0], snd[0]);
        __syncwarp();            // s_cmin and s_col may be rewritten by the next unit from here on

        // ---- rows: best / second-best group minimum (strict < keeps the earliest group on ties)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float v = rowmin[r];
            if (KO & 2) { best[r] = fminf(best[r], v); continue; }
            s_second[wic][r][lane] = fminf(s_second[wic][r][lane], fmaxf(best[r], v));
            if (v < best[r]) { best[r] = v; s_bgrp[wic][r][lane] = cg; }
        }
        // ---- columns: finish the previous unit's merge, remember this unit's result for the next iteration
        if (prev) merge_second(w.colsec, prev_idx, prev_old, prev_key, prev_sec);
        {
            const int j = cg * kGroup + lane;
            prev = j < M;
            int mv = mm.x & ~31, sv = mm.y & ~31;
            if (mv < 0) mv = 0;                 // a rounding residue below zero
            if (sv < 0) sv = 0;
            prev_key = ((u64)(unsigned)mv << 32) | ((unsigned)rb * 32u + (unsigned)(mm.x & 31));
            prev_sec = (unsigned)sv;
            prev_idx = (size_t)b * M + j;
        }
        // ---- advance; a new row block (or the end of the range) flushes this block's rows
        if (nrb != rb || nb != b || it + 1 == n_units) {
            flush_rows(b, rb);
            new_block = true;
        }
        cg = ncg; rb = nrb; b = nb;
    }
    if (prev) {
        const u64 old = atomicMin(&w.colkey[prev_idx], prev_key);
        merge_second(w.colsec, prev_idx, old, prev_key, prev_sec);
    }
}

template <int R, int OCC, bool PP, int KO = 0>
static int launch_filter_t(const float *pc1, const float *pc2, int B, int N, int M, const FwdWs &w, cudaStream_t st) {
    const int rows = 32 * R;
    const int n_rb = (N + rows - 1) / rows;
    const int n_cg = (M + kGroup - 1) / kGroup;
    const long long total = (long long)B * n_rb * n_cg;
    const int sms = sm_count();
    if (sms <= 0) return fail((int)cudaErrorNoDevice, "rlg_chamfer_fwd: no CUDA device");
    int ctas_per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, chamfer_filter_kernel<R, OCC, PP, KO>,
                                                                  kFWarps * 32, 0);
    if (e != cudaSuccess || ctas_per_sm < 1) { cudaGetLastError(); ctas_per_sm = 1; }
    long long grid = (long long)sms * ctas_per_sm;
    const long long max_useful = (total + kFWarps - 1) / kFWarps;      // at least one unit per warp
    if (grid > max_useful) grid = max_useful;
    if (grid < 1) grid = 1;
    cudaError_t le = launch_pdl(chamfer_filter_kernel<R, OCC, PP, KO>, dim3((unsigned)grid), dim3(kFWarps * 32), 0, st, pc1, pc2, B, N, M,
                                n_rb, n_cg, total, w);
    if (le != cudaSuccess) { cudaGetLastError(); return fail((int)le, "chamfer_filter_kernel: %s", cudaGetErrorString(le)); }
    return check_launch("chamfer_filter_kernel");
}

// rows per lane: the widest tile that does not pad the row count by more than ~12 %
int filter_pick_rows(int N) {
    const int cand[3] = {16, 8, 4};
    for (int k = 0; k < 3; ++k) {
        const int rows = 32 * cand[k];
        const long long padded = (long long)((N + rows - 1) / rows) * rows;
        if (padded * 8 <= (long long)N * 9) return cand[k];
    }
    return 4;
}

int launch_filter(const float *pc1, const float *pc2, int B, int N, int M, int variant, const FwdWs &w, int *rows_per_lane,
                  cudaStream_t st) {
#ifdef RLG_EXPERIMENTS
    // experimental tile shapes and knock-out timing variants (tools/sweep_tile.py; variants >= 10 return WRONG results).
    // They exist only in the experiments build (build.py --experiments -> librlg_b200_exp.so).
    if (variant != 0) {
        const int Rv = (variant >= 7 && variant <= 8) ? 8 : (variant == 9 ? 4 : 16);
        *rows_per_lane = Rv;
        switch (variant) {
            case 1: return launch_filter_t<16, 2, true>(pc1, pc2, B, N, M, w, st);
            case 2: return launch_filter_t<16, 2, false>(pc1, pc2, B, N, M, w, st);
            case 3: return launch_filter_t<16, 3, true>(pc1, pc2, B, N, M, w, st);
            case 4: return launch_filter_t<16, 3, false>(pc1, pc2, B, N, M, w, st);
            case 7: return launch_filter_t<8, 3, true>(pc1, pc2, B, N, M, w, st);
            case 8: return launch_filter_t<8, 4, false>(pc1, pc2, B, N, M, w, st);
            case 9: return launch_filter_t<4, 4, true>(pc1, pc2, B, N, M, w, st);
            case 10: return launch_filter_t<16, 2, false, 1>(pc1, pc2, B, N, M, w, st);
            case 11: return launch_filter_t<16, 2, false, 2>(pc1, pc2, B, N, M, w, st);
            case 12: return launch_filter_t<16, 2, false, 3>(pc1, pc2, B, N, M, w, st);
            case 13: return launch_filter_t<16, 2, false, 4>(pc1, pc2, B, N, M, w, st);
            case 14: return launch_filter_t<16, 2, false, 11>(pc1, pc2, B, N, M, w, st);
            case 15: return launch_filter_t<16, 2, false, 15>(pc1, pc2, B, N, M, w, st);
            default: return fail(RLG_ERR_UNSUPPORTED, "rlg_chamfer_fwd: unknown experimental variant %d", variant);
        }
    }
#else
    if (variant != 0) return fail(RLG_ERR_UNSUPPORTED, "rlg_chamfer_fwd: kernel variants exist only in the experiments build");
#endif
    const int R = filter_pick_rows(N);
    *rows_per_lane = R;
    if (R == 16) return launch_filter_t<16, 2, true>(pc1, pc2, B, N, M, w, st);
    if (R == 8) return launch_filter_t<8, 3, true>(pc1, pc2, B, N, M, w, st);
    return launch_filter_t<4, 4, true>(pc1, pc2, B, N, M, w, st);
}

// ------------------------------------------------------------------------------------------------------------
// finalize: exact refinement, outputs, deterministic means / loss, workspace back to all-ones
// ------------------------------------------------------------------------------------------------------------
static constexpr int kFin2Threads = 256;                            // one query point per thread
static constexpr int kFin2Warps = kFin2Threads / 32;

struct Fin2Ws {
    unsigned *global_counter;   // 1
    unsigned *cloud_counter;    // B
    double *partial;            // 2*B*chunks_max
    int chunks_max;
};

static size_t fin2_counter_bytes(int B) { return align_up(sizeof(unsigned) * (1 + (size_t)B), 256); }
static int fin2_chunks(int n) { return (n + kFin2Threads - 1) / kFin2Threads; }

size_t finalize2_ws_bytes(int B, int N, int M) {
    const int cm = fin2_chunks(N > M ? N : M);
    return fin2_counter_bytes(B) + align_up(sizeof(double) * 2 * (size_t)B * cm, 256);
}

// exact direct-form scans of candidates [first, first+step, ...) of c for query (px,py,pz): this thread's smallest t
// (four independent candidates in flight), and the lowest index with t <= h
__device__ __forceinline__ float scan_min_strided(const float *__restrict__ c, int nc, int first, int step, float px, float py,
                                                  float pz) {
    float lt = INFINITY;
    int j = first;
    for (; j + 3 * step < nc; j += 4 * step) {
        const int j1 = j + step, j2 = j + 2 * step, j3 = j + 3 * step;
        const float t0 = sqdist(px, py, pz, c[3 * j], c[3 * j + 1], c[3 * j + 2]);
        const float t1 = sqdist(px, py, pz, c[3 * j1], c[3 * j1 + 1], c[3 * j1 + 2]);
        const float t2 = sqdist(px, py, pz, c[3 * j2], c[3 * j2 + 1], c[3 * j2 + 2]);
        const float t3 = sqdist(px, py, pz, c[3 * j3], c[3 * j3 + 1], c[3 * j3 + 2]);
        lt = fminf(fminf(lt, t0), fminf(t1, fminf(t2, t3)));
    }
    for (; j < nc; j += step) lt = fminf(lt, sqdist(px, py, pz, c[3 * j], c[3 * j + 1], c[3 * j + 2]));
    return lt;
}
__device__ __forceinline__ int scan_first_le_strided(const float *__restrict__ c, int nc, int first, int step, float px,
                                                     float py, float pz, float h) {
    for (int j = first; j < nc; j += step)
        if (sqdist(px, py, pz, c[3 * j], c[3 * j + 1], c[3 * j + 2]) <= h) return j;
    return 0x7fffffff;
}

// Tie rule (all paths): the reference's torch.min runs on the sqrt-ed distances, so candidates whose squared distances
// share one sqrtf tie and the lowest index wins (sqrt_window_top, common.cuh; oracle ORC_TIE_FAITHFUL).
// TWO (experiments build only): the sweep also reported the runner-up group and the third-smallest group minimum
// (first-generation tensor-core sweep, rows_per_lane == 0): ambiguous points are refined on two groups
template <bool SMEM, bool TWO>
__global__ void __launch_bounds__(kFin2Threads) chamfer_finalize2_kernel(
    const float *__restrict__ pc1, const float *__restrict__ pc2, int N, int M, int rows_per_lane, FwdWs w, Fin2Ws fw,
    float *__restrict__ d1, float *__restrict__ d2, int32_t *__restrict__ i1, int32_t *__restrict__ i2,
    float *__restrict__ mean1, float *__restrict__ mean2, float *__restrict__ loss, float loss_w1, float loss_w2,
    float *__restrict__ zero1, float *__restrict__ zero2) {
    extern __shared__ __align__(16) float s_cand[];
    __shared__ double red[kFin2Warps];
    __shared__ float s_q[3][kFin2Threads];            // query coordinates of the CTA's points
    __shared__ short s_list[kFin2Threads];            // ambiguous points of the CTA, in thread order
    __shared__ int s_wcount[kFin2Warps];
    __shared__ unsigned s_wmin[kFin2Warps];           // per-warp smallest t bits / lowest index of the point being scanned
    __shared__ int s_wj[kFin2Warps];
    const int chunk = blockIdx.x, b = blockIdx.y, dir = blockIdx.z;
    const int B = gridDim.y;
    // dir 0: queries = pc1 rows; a candidate group = 32 consecutive columns of pc2
    // dir 1: queries = pc2 columns; a candidate group = the rows {rb*32R + k*32 + lane : k < R} one filter lane owned
    const int nq = dir ? M : N, nc = dir ? N : M;
    const int n_chunks = (nq + kFin2Threads - 1) / kFin2Threads;
    if (chunk >= n_chunks) return;                      // grid.x is sized for the longer direction
    const int R = rows_per_lane;                        // 0: tensor-core filter, groups of 32 consecutive candidates both ways
    const bool strided = dir && R > 0;
    const int gsz = strided ? R : kGroup;               // a power of two <= 32
    const float *q = (dir ? pc2 : pc1) + (size_t)b * nq * 3;
    const float *cglob = (dir ? pc1 : pc2) + (size_t)b * nc * 3;
    u64 *keys = (dir ? w.colkey : w.rowkey) + (size_t)b * nq;
    unsigned *secs = (dir ? w.colsec : w.rowsec) + (size_t)b * nq;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

    // Everything up to pdl_wait() reads only the clouds, which were complete before the pair sweep started: under a
    // programmatic dependent launch it overlaps the tail of the pair sweep.
    pdl_launch_dependents();
    const int i = chunk * kFin2Threads + tid;
    const bool live = i < nq;
    u64 key = kKeyInit;
    unsigned sec = 0xffffffffu, sgrp = 0, third = 0xffffffffu;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (live) { qx = __ldg(q + 3 * (size_t)i); qy = __ldg(q + 3 * (size_t)i + 1); qz = __ldg(q + 3 * (size_t)i + 2); }

    const float *__restrict__ c = cglob;
    if (SMEM) {
        const int nf = nc * 3;
        if ((reinterpret_cast<uintptr_t>(cglob) & 15u) == 0) {
            const int n4 = nf >> 2;
            const float4 *src4 = reinterpret_cast<const float4 *>(cglob);
            float4 *dst4 = reinterpret_cast<float4 *>(s_cand);
            for (int e = tid; e < n4; e += kFin2Threads) dst4[e] = __ldg(src4 + e);
            for (int e = (n4 << 2) + tid; e < nf; e += kFin2Threads) s_cand[e] = __ldg(cglob + e);
        } else {
            for (int e = tid; e < nf; e += kFin2Threads) s_cand[e] = __ldg(cglob + e);
        }
        c = s_cand;
    }
    pdl_wait();                                         // from here on: results of the pair sweep
    if (live) {
        key = keys[i];
        sec = secs[i];
        if (TWO) {
            sgrp = (dir ? w.colsg : w.rowsg)[(size_t)b * nq + i];
            third = (dir ? w.colth : w.rowth)[(size_t)b * nq + i];
        }
    }
    // largest squared norm of the OTHER cloud (the filter kernel published its bitwise complement)
    const float onrm = __uint_as_float(~w.nrm[dir ? b : B + b]);
    if (live) {
        keys[i] = kKeyInit;
        secs[i] = 0xffffffffu;
        float *z = dir ? zero2 : zero1;                 // optional: zero-fill the gradient rows of this point
        if (z != nullptr) {
            z += ((size_t)b * nq + i) * 3;
            z[0] = 0.0f; z[1] = 0.0f; z[2] = 0.0f;
        }
    }
    s_q[0][tid] = qx; s_q[1][tid] = qy; s_q[2][tid] = qz;

    const float val = __uint_as_float((unsigned)(key >> 32));
    const float sv = __uint_as_float(sec);
    const float margin = (R > 0 ? kMarginC : kMarginT) * (norm2(qx, qy, qz) + onrm) + (strided ? kMarginQ * val : 0.0f);
    // any NaN (untouched key, non-finite input) makes the comparison false -> treated as ambiguous
    const bool amb2 = live && !(sv > val + margin);     // a second group is within the margin of the best one
    const bool two = TWO && amb2 && (__uint_as_float(third) > val + margin);
    const bool amb = amb2 && !two;
    // deterministic list of the CTA's ambiguous points
    const unsigned am = __ballot_sync(0xffffffffu, amb);
    if (lane == 0) s_wcount[wid] = __popc(am);
    __syncthreads();                                    // candidates staged, s_q and s_wcount visible
    int n_amb = 0, my_off = 0;
#pragma unroll
    for (int k = 0; k < kFin2Warps; ++k) {
        if (k == wid) my_off = n_amb;
        n_amb += s_wcount[k];
    }
    if (amb) s_list[my_off + __popc(am & ((1u << lane) - 1u))] = (short)tid;

    // ---- fast path: the exact winner lies in the best group (or, TWO, in one of two groups).  Pass 1 finds the smallest
    // squared distance, pass 2 the lowest index sharing its square root.  Indices past the end are clamped to the last
    // point: a real candidate with the highest index, so it never wins a tie it should not.
    float bt = INFINITY;
    int bj = 0x7fffffff;
    if (live && !amb) {
        const unsigned grp = (unsigned)(key & 0xffffffffu);
        const int base = strided ? (int)(grp >> 5) * (32 * R) + (int)(grp & 31u) : (int)grp * kGroup;
        const int stride = strided ? 32 : 1;
        const int base2 = (int)sgrp * kGroup;
#pragma unroll 4
        for (int k = 0; k < gsz; ++k) {
            const int j = min(base + ((k + lane) & (gsz - 1)) * stride, nc - 1);   // rotated per lane: conflict-free
            bt = fminf(bt, sqdist(qx, qy, qz, c[3 * j], c[3 * j + 1], c[3 * j + 2]));
        }
        if (TWO && two) {
            for (int k = 0; k < kGroup; ++k) {
                const int j = min(base2 + ((k + lane) & (kGroup - 1)), nc - 1);
                bt = fminf(bt, sqdist(qx, qy, qz, c[3 * j], c[3 * j + 1], c[3 * j + 2]));
            }
        }
        if (bt < INFINITY) {
            const float h = sqrt_window_top(bt, __fsqrt_rn(bt));
#pragma unroll 4
            for (int k = 0; k < gsz; ++k) {
                const int j = min(base + ((k + lane) & (gsz - 1)) * stride, nc - 1);
                if (sqdist(qx, qy, qz, c[3 * j], c[3 * j + 1], c[3 * j + 2]) <= h) bj = min(bj, j);
            }
            if (TWO && two) {
                for (int k = 0; k < kGroup; ++k) {
                    const int j = min(base2 + ((k + lane) & (kGroup - 1)), nc - 1);
                    if (sqdist(qx, qy, qz, c[3 * j], c[3 * j + 1], c[3 * j + 2]) <= h) bj = min(bj, j);
                }
            }
        }
    }
    __syncthreads();                                    // s_list complete
    // ---- ambiguous points, one after the other: all 256 threads scan the candidates exactly (two passes, as above)
    for (int a = 0; a < n_amb; ++a) {
        const int t_a = s_list[a];
        const float ax = s_q[0][t_a], ay = s_q[1][t_a], az = s_q[2][t_a];
        const float lt = scan_min_strided(c, nc, tid, kFin2Threads, ax, ay, az);
        const unsigned mt = __reduce_min_sync(0xffffffffu, __float_as_uint(lt));     // lt >= 0 or +inf: orders as unsigned
        if (lane == 0) s_wmin[wid] = mt;
        __syncthreads();
        unsigned mb = s_wmin[0];
#pragma unroll
        for (int k = 1; k < kFin2Warps; ++k) mb = min(mb, s_wmin[k]);
        const float m = __uint_as_float(mb);
        const float h = m < INFINITY ? sqrt_window_top(m, __fsqrt_rn(m)) : m;
        const int lj = scan_first_le_strided(c, nc, tid, kFin2Threads, ax, ay, az, h);
        const int mj = __reduce_min_sync(0xffffffffu, lj);
        if (lane == 0) s_wj[wid] = mj;
        __syncthreads();
        if (tid == t_a) {
            int jj = s_wj[0];
#pragma unroll
            for (int k = 1; k < kFin2Warps; ++k) jj = min(jj, s_wj[k]);
            bt = m;
            bj = jj;
        }
        __syncthreads();
    }
    double dist_d = 0.0;
    if (live) {
        if (!(bt < INFINITY) || bj == 0x7fffffff) {
            // only reachable with non-finite input (outside the contract): candidate 0
            bj = 0;
            bt = sqdist(qx, qy, qz, cglob[0], cglob[1], cglob[2]);
        }
        const float dist = __fsqrt_rn(bt);
        (dir ? d2 : d1)[(size_t)b * nq + i] = dist;
        (dir ? i2 : i1)[(size_t)b * nq + i] = bj;
        dist_d = (double)dist;
    }

    // CTA partial sum in a fixed order: lane tree, then warps 0..7 sequentially
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) dist_d += __shfl_down_sync(0xffffffffu, dist_d, s);
    if (lane == 0) red[wid] = dist_d;
    __syncthreads();
    if (tid != 0) return;
    double part = 0.0;
#pragma unroll
    for (int k = 0; k < kFin2Warps; ++k) part += red[k];
    volatile double *parts = fw.partial + (size_t)(b * 2) * fw.chunks_max;
    parts[(size_t)dir * fw.chunks_max + chunk] = part;
    __threadfence();
    const int nch1 = (N + kFin2Threads - 1) / kFin2Threads, nch2 = (M + kFin2Threads - 1) / kFin2Threads;
    // counters start at 0xffffffff (the workspace's all-ones state): the k-th arrival reads k-2
    if (atomicAdd(fw.cloud_counter + b, 1u) != (unsigned)(nch1 + nch2 - 2)) return;
    // last CTA of this cloud: every reader of its norms is done
    __threadfence();
    fw.cloud_counter[b] = 0xffffffffu;
    w.nrm[b] = 0xffffffffu;
    w.nrm[B + b] = 0xffffffffu;
    if (mean1 == nullptr) return;
    double t1 = 0.0, t2 = 0.0;
    for (int k = 0; k < nch1; ++k) t1 += parts[k];
    for (int k = 0; k < nch2; ++k) t2 += parts[(size_t)fw.chunks_max + k];
    mean1[b] = (float)(t1 / (double)N);
    mean2[b] = (float)(t2 / (double)M);
    if (loss == nullptr) return;
    __threadfence();
    if (atomicAdd(fw.global_counter, 1u) != (unsigned)(B - 2)) return;
    __threadfence();
    *fw.global_counter = 0xffffffffu;
    const volatile float *m1 = mean1, *m2 = mean2;
    double acc = 0.0;
    for (int k = 0; k < B; ++k) acc += (double)loss_w1 * (double)m1[k] + (double)loss_w2 * (double)m2[k];
    *loss = (float)acc;
}

int launch_finalize2(const float *pc1, const float *pc2, int B, int N, int M, int rows_per_lane, const FwdWs &w,
                     void *fin_ws, float *d1, float *d2, int32_t *i1, int32_t *i2, float *mean1, float *mean2,
                     float *loss, float w1, float w2, float *zero1, float *zero2, cudaStream_t st) {
    Fin2Ws fw;
    fw.global_counter = (unsigned *)fin_ws;
    fw.cloud_counter = fw.global_counter + 1;
    fw.partial = (double *)((char *)fin_ws + fin2_counter_bytes(B));
    fw.chunks_max = fin2_chunks(N > M ? N : M);
    // the candidate cloud of a (cloud, direction) is staged in shared memory when it fits (<= 200 KB)
    const size_t cand_bytes = (size_t)(N > M ? N : M) * 3 * sizeof(float);
    const bool in_smem = cand_bytes <= 200u * 1024u;
    dim3 grid(fw.chunks_max, B, 2);
    const bool two = rows_per_lane == 0;
    auto launch = [&](auto kernel, size_t dyn) -> int {
        if (dyn > 40u * 1024u) {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200u * 1024u));
            if (e != cudaSuccess) { cudaGetLastError(); return fail((int)e, "rlg_chamfer_fwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); }
        }
        cudaError_t le = launch_pdl(kernel, grid, dim3(kFin2Threads), dyn, st, pc1, pc2, N, M, rows_per_lane, w, fw, d1, d2, i1, i2,
                                    mean1, mean2, loss, w1, w2, zero1, zero2);
        if (le != cudaSuccess) { cudaGetLastError(); return fail((int)le, "chamfer_finalize2_kernel: %s", cudaGetErrorString(le)); }
        return 0;
    };
    const size_t dyn = in_smem ? align_up(cand_bytes, 16) : 0;
    int rc;
#ifdef RLG_EXPERIMENTS
    if (two) rc = in_smem ? launch(chamfer_finalize2_kernel<true, true>, dyn) : launch(chamfer_finalize2_kernel<false, true>, dyn);
    else
#else
    if (two) return fail(RLG_ERR_UNSUPPORTED, "chamfer_finalize2_kernel: two-group refinement exists only in the experiments build");
#endif
    rc = in_smem ? launch(chamfer_finalize2_kernel<true, false>, dyn) : launch(chamfer_finalize2_kernel<false, false>, dyn);
    if (rc) return rc;
    return check_launch("chamfer_finalize2_kernel");
}

}  // namespace rlg
