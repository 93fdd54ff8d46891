// chamfer_tcfilter.cu -- the Chamfer pair sweep with the contraction on the tensor cores (tcgen05, kind::tf32).
//
// Same filter-and-refine scheme as chamfer_filter.cu (replaces utils/losses.py:29-33): a cheap filter value
// t~(i,j) ~ |x_i - y_j|^2 for EVERY pair, group minima / second-best group per query, and the exact direct-form
// arithmetic only on the candidates that can still win (chamfer_finalize2_kernel).  The outputs therefore do not
// depend on how the filter rounds -- only the margin below has to bound its error.
//
// Why a second filter kernel: on the CUDA cores a pair costs 4 FP32-pipe slots (3 FFMA + 1 FADD) AND 2 min slots, all
// through one dispatch port per SM sub-partition (tools/ubench2.cu) -- the FFMA pipe cannot get past ~60 % busy.  Here
// the 3-term contraction moves to the otherwise idle tensor pipe and the CUDA cores only take minima:
//
//   t(i,j) = |x_i|^2 + |y_j|^2 - 2 x_i.y_j  =  A[i,:] . B[j,:]   with K = 16 tf32 columns (error-compensated split):
//       A row (query x)      x0h x0h x0l | x1h x1h x1l | x2h x2h x2l | nxh nxm nxl | 1   1   1   | 0
//       B row (candidate y)  Y0h Y0l Y0h | Y1h Y1l Y1h | Y2h Y2l Y2h | 1   1   1   | nyh nym nyl | 0      (Y = -2y)
//   where vh = tf32(v), vl = tf32(v - vh) (so v = vh + vl to 2^-22 relative) and the squared norms are split three ways
//   (exactly).  Every tf32 x tf32 product is exact in fp32; only the vl x vl terms are dropped.
//
//   D[128 queries x 128 candidates] (fp32, TMEM) = A[128 x 16] . B[128 x 16]^T : two tcgen05.mma (M=128, N=128, K=8)
//   per half of a 256-candidate tile.  Each query sits on one TMEM lane, so its minimum over the candidates is a
//   per-thread reduction of tcgen05.ld registers: no shuffles, no shared-memory traffic, no atomics inside the sweep.  Both directions run as the same problem
//   with the roles of the clouds swapped (the tensor pipe has the headroom), so a "group" is 32 consecutive candidates
//   in either direction.
//
// Error bound of the filter (u = 2^-24, a = |x_i|, b = |y_j|, S = a^2 + b^2):
//   dropped / residual split terms       3 * 2^-22 * 2 a b                                        <= 12 u S
//   accumulation inside the tensor core  not specified bit for bit; modelled as truncation to an fp32 significand
//                                        (1 ulp = 2 u of the running sum, which never exceeds S) at every one of the
//                                        16 product additions and the 2 accumulator updates                <= 36 u S
//   computed norms                                                                                <=  3 u S
//   => |t~ - t| <= 51 u S; with the direct form's <= 10 u S two candidates can swap only within 2 * 61 u S = 122 u S.
//   kMarginT = 128 u (chamfer_filter.cu, finalize).  tests/test_chamfer_gpu.py measures the filter against float64: the
//   largest error seen is ~4.5 u S, and the whole parity suite runs on this path as well as on the FP32 one.
//
// Besides the best group and the runner-up VALUE (what the FP32 sweep reports), this sweep can also report (TOP3, chosen
// by launch_tcfilter for candidate clouds beyond 4096 points) the runner-up's group and the third-smallest group minimum: a point whose runner-up is within the margin but whose third is not is
// refined exactly on two groups (64 candidates) instead of the whole candidate cloud -- at N = 16384 that is the
// difference between a finalize of ~1.6 ms and one that is a fraction of the sweep.
//
// Roles (one CTA per SM, persistent over (direction, cloud, 128-query block) tasks):
//   A CTA owns a contiguous range of the tasks and walks it in segments of up to 8 query blocks of one (direction,
//   cloud): each converted candidate tile is multiplied against all of them (a "visit" = one query block x one tile).
//   warps 0-15  four epilogue groups of four warps; a visit is issued as two 128-column halves and half-visit h goes to
//               TMEM accumulator h % 4 == group h % 4                  thread = query = TMEM lane
//   warps 16-17 producers: convert 256 candidates per tile (four per thread; and the segment's queries) to the split-tf32
//               rows above, written straight into the K-major SWIZZLE_128B operand tiles
//   warp 18     MMA issuer (one elected lane), TMEM owner
//   At the end of a segment the four groups' running results are merged per query in shared memory and published with
//   plain stores: every candidate of these queries was seen by this CTA, so the workspace needs no atomics here.
//
// Environment (measurement only): RLG_TF_DEBUG=1|2 prints per-role cycle counters and per-CTA time spans after a
// synchronize; RLG_TF_TOP3=0|1 and RLG_TF_SPLIT=0|1|2 override the runner-up report and the work split; RLG_TF_KO=2
// skips the minima (timing experiment, wrong results).
#ifdef RLG_EXPERIMENTS      // first-generation kernel, kept for A/B timing against chamfer_tcsweep.cu
#include "common.cuh"
#include "tcgen05.cuh"
#include <math.h>
#include <stdlib.h>
#include <stdio.h>

namespace rlg {

static constexpr int kTQ = 128;                       // queries per block (UMMA M)
static constexpr int kTC = 256;                       // candidates per tile (UMMA N)
static constexpr int kQmax = 8;                       // query blocks that share one pass over the candidate tiles
static constexpr int kTfEpGroups = 4;                 // epilogue warp groups == 128-column TMEM accumulators
static constexpr int kTfEpWarps = 4 * kTfEpGroups, kTfPrWarps = 2;
static constexpr int kTfMmaWarp = kTfEpWarps + kTfPrWarps;
static constexpr int kTfThreads = (kTfEpWarps + kTfPrWarps + 1) * 32;
static constexpr int kTfPrThreads = kTfPrWarps * 32;
static constexpr int kARows = kTQ / kTfPrThreads, kBRows = kTC / kTfPrThreads;   // rows per producer thread
static constexpr uint32_t kTfABytes = kTQ * 128, kTfBBytes = kTC * 128;
// A rows carry K = 16 tf32 = 64 bytes, half of a SWIZZLE_128B row: two query blocks share one 16 KB tile (block q sits in
// 16-byte chunks 4*(q&1) .. 4*(q&1)+3 of tile q>>1; the descriptor's start address selects the half)
static constexpr uint32_t kTfOffB = (kQmax / 2) * kTfABytes;                   // two candidate tiles follow the A tiles
static constexpr int kStW = 5;                                                 // state words: best, second, third, best grp, second grp
static constexpr uint32_t kTfOffState = kTfOffB + 2 * kTfBBytes;               // [groups][kQmax][kStW][128]
static constexpr uint32_t kTfStateBytes = (uint32_t)kTfEpGroups * kQmax * kTQ * kStW * 4u;
static constexpr uint32_t kTfOffBar = kTfOffState + kTfStateBytes;
static constexpr uint32_t kTfSmem = kTfOffBar + 256 + 1024;                    // + barriers + alignment slack
static constexpr float kTfBig = 1.0e30f;

__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void tf32_split2(float v, float &h, float &l) {
    h = tf32_rn(v);
    l = tf32_rn(v - h);                                // v - h is exact
}
__device__ __forceinline__ void tf32_split3(float v, float &h, float &m, float &l) {
    h = tf32_rn(v);
    const float r = v - h;                             // exact
    m = tf32_rn(r);
    l = r - m;                                         // exact, at most 3 significant bits
}
// row `row` of a K-major SW128 tile: 16 floats as four 16-byte chunks
__device__ __forceinline__ void tf_store_row(unsigned char *tile, int row, const float (&e)[16], uint32_t chunk0 = 0) {
    const uint32_t base = (uint32_t)row * 128u, x = (uint32_t)row & 7u;
#pragma unroll
    for (uint32_t c = 0; c < 4; ++c)
        *reinterpret_cast<float4 *>(tile + base + (((chunk0 + c) ^ x) << 4)) = make_float4(e[4 * c], e[4 * c + 1], e[4 * c + 2], e[4 * c + 3]);
}
__device__ __forceinline__ float norm2_tf(float x, float y, float z) { return fmaf(z, z, fmaf(y, y, x * x)); }

// minimum of 32 values, as a tree of 3-input minima
__device__ __forceinline__ float min32(const float *v) {
    float m[11];
#pragma unroll
    for (int i = 0; i < 10; ++i) m[i] = min3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
    m[10] = fminf(v[30], v[31]);
    const float n0 = min3(m[0], m[1], m[2]), n1 = min3(m[3], m[4], m[5]), n2 = min3(m[6], m[7], m[8]);
    return min3(min3(n0, n1, n2), m[9], m[10]);
}

struct TfSeg { int dir, b, qb0, Q, nq, nc, n_ct, next; };

// RLG_TF_DEBUG=1: cycle counters of CTA 0 (one lane per role), printed by launch_tcfilter after a synchronize
__device__ unsigned long long g_tf_dbg[16];
__device__ unsigned long long g_tf_span[160][6];      // per CTA: globaltimer at entry, after setup, MMA loop end, exit
__device__ __forceinline__ unsigned long long tf_gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define TF_T0(v) const long long v = dbg ? clock64() : 0
#define TF_ACC(slot, v) do { if (dbg) dbg_acc[(slot) & 3] += (unsigned long long)(clock64() - (v)); } while (0)
#define TF_FLUSH(base) do { if (dbg) { for (int z = 0; z < 4; ++z) g_tf_dbg[(base) + z] = dbg_acc[z]; } } while (0)

// TOP3: also track the runner-up's group and the third-smallest group minimum (5 more issue slots per 32 candidates;
// pays off when ambiguous points would otherwise rescan a large candidate cloud -- see launch_tcfilter)
template <bool TOP3>
__global__ void __launch_bounds__(kTfThreads, 1)
chamfer_tcfilter_kernel(const float *__restrict__ pc1, const float *__restrict__ pc2, int B, int N, int M, int n_tasks,
                        int qb1, int qb2, int split, FwdWs w, int ko) {
    extern __shared__ unsigned char tf_smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if ((ko & 16) && tid == 0 && blockIdx.x < 160) g_tf_span[blockIdx.x][0] = tf_gtime();
    const uint32_t pad = (1024u - (smem_u32(tf_smem_raw) & 1023u)) & 1023u;
    unsigned char *smem = tf_smem_raw + pad;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bars = sbase + kTfOffBar;
    const uint32_t bar_afull = bars, bar_aempty = bars + 8 * kQmax;                     // [kQmax], [1]
    const uint32_t bar_bfull = bar_aempty + 8, bar_bempty = bar_bfull + 16;             // [2] each
    const uint32_t bar_accfull = bar_bempty + 16, bar_accempty = bar_accfull + 32;      // [4] each: (accumulator, column half)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + kTfOffBar + 8 * (kQmax + 13));

    if (tid == 0) {
        for (int q = 0; q < kQmax; ++q) mbar_init(bar_afull + 8 * q, kTfPrWarps * 32);
        mbar_init(bar_aempty, 1);
        for (int k = 0; k < 2; ++k) {
            mbar_init(bar_bfull + 8 * k, kTfPrWarps * 32); mbar_init(bar_bempty + 8 * k, 1);
        }
        for (int k = 0; k < 4; ++k) { mbar_init(bar_accfull + 8 * k, 1); mbar_init(bar_accempty + 8 * k, 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kTfMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    unsigned long long dbg_acc[4] = {0, 0, 0, 0};
    const bool dbg = (ko & 16) && blockIdx.x == 0 && lane == 0 && (warp == kTfMmaWarp || warp == kTfEpWarps || warp == 0);
    pdl_launch_dependents();
    pdl_wait();                                   // clouds and workspace may come from the kernels right before this one
    if ((ko & 16) && tid == 0 && blockIdx.x < 160) g_tf_span[blockIdx.x][1] = tf_gtime();

    // Tasks (direction, cloud, 128-query block) in direction-major order; a CTA owns a contiguous range and walks it in
    // SEGMENTS of up to kQmax query blocks of one (direction, cloud): they share every converted candidate tile.
    //   split 0: even contiguous split of the task list (a CTA may straddle two clouds: two segments)
    //   split 1: at most as many (direction, cloud) units as CTAs: every unit is cut into equal chunks, one per CTA, so
    //            no CTA pays the per-segment costs (a pass of candidate conversions, a pipeline refill) twice
    //   split 2: many more units than CTAs: whole units per CTA
    const int G = (int)gridDim.x, cta = (int)blockIdx.x, U = 2 * B;
    auto unit_start = [&](int u) { return u < B ? u * qb1 : B * qb1 + (u - B) * qb2; };
    int t_begin, t_end;
    if (split == 1) {
        const int c_lo = G / U, extra = G - c_lo * U;                // the first `extra` units get one chunk more
        int u, ci, c;
        if (cta < extra * (c_lo + 1)) { u = cta / (c_lo + 1); ci = cta - u * (c_lo + 1); c = c_lo + 1; }
        else { const int r = cta - extra * (c_lo + 1); u = extra + r / c_lo; ci = r - (u - extra) * c_lo; c = c_lo; }
        const int T = u < B ? qb1 : qb2, s0 = unit_start(u);
        t_begin = s0 + (int)((long long)ci * T / c);
        t_end = s0 + (int)((long long)(ci + 1) * T / c);
    } else if (split == 2) {
        t_begin = unit_start((int)((long long)cta * U / G));
        t_end = unit_start((int)((long long)(cta + 1) * U / G));
    } else {
        const int per = n_tasks / G, rem = n_tasks - per * G;
        t_begin = cta * per + min(cta, rem);
        t_end = t_begin + per + (cta < rem ? 1 : 0);
    }
    if ((ko & 16) && tid == 0 && blockIdx.x < 160) { g_tf_span[blockIdx.x][4] = (unsigned long long)(t_end - t_begin); g_tf_span[blockIdx.x][5] = (unsigned long long)t_begin; }
    auto seg_at = [&](int task) {
        TfSeg sg;
        const int n0 = B * qb1;
        int qbn;
        if (task < n0) { sg.dir = 0; sg.b = task / qb1; sg.qb0 = task - sg.b * qb1; qbn = qb1; }
        else { const int r = task - n0; sg.dir = 1; sg.b = r / qb2; sg.qb0 = r - sg.b * qb2; qbn = qb2; }
        sg.Q = min(kQmax, min(t_end - task, qbn - sg.qb0));
        sg.nq = sg.dir ? M : N;
        sg.nc = sg.dir ? N : M;
        sg.n_ct = (sg.nc + kTC - 1) / kTC;
        sg.next = task + sg.Q;
        return sg;
    };

    if (warp == kTfMmaWarp) {
        // =========================== MMA issuer ===========================
        const bool leader = elect_one();
        TF_T0(t_all);
        const uint32_t idesc = umma_idesc_tf32(kTQ, kTC / 2);
        const uint64_t ad0 = umma_desc(sbase), bd0 = umma_desc(sbase + kTfOffB);
        const uint64_t a_inc = (uint64_t)(kTfABytes >> 4), b_inc = (uint64_t)(kTfBBytes >> 4);
        uint32_t gt = 0, bt = 0, sn = 0;
        for (int task = t_begin; task < t_end;) {
            const TfSeg sg = seg_at(task);
            task = sg.next;
            for (int k = 0; k < sg.n_ct; ++k, ++bt) {
                const uint32_t sb = bt & 1u;
                { TF_T0(t0); mbar_wait_spin(bar_bfull + 8 * sb, (bt >> 1) & 1u); TF_ACC(2, t0); }
                const uint64_t bd = bd0 + (uint64_t)sb * b_inc;
                for (int q = 0; q < sg.Q; ++q) {
                    if (k == 0) { TF_T0(t0); mbar_wait_spin(bar_afull + 8 * q, sn & 1u); TF_ACC(1, t0); }
                    // a visit (query block x candidate tile) is issued as two 128-column halves; half-visit ht goes to
                    // accumulator ht % 4, which belongs to epilogue group ht % 4
#pragma unroll
                    for (uint32_t hf = 0; hf < 2; ++hf, ++gt) {
                        const uint32_t ab = gt & 3u;
                        { TF_T0(t0); mbar_wait_spin(bar_accempty + 8 * ab, ((gt >> 2) & 1u) ^ 1u); TF_ACC(3, t0); }
                        tc_fence_after();
                        if (leader) {
                            const uint64_t ad = ad0 + (uint64_t)(q >> 1) * a_inc + (uint64_t)((q & 1) * 4);
                            const uint64_t bh = bd + (uint64_t)hf * (b_inc >> 1);
                            const uint32_t d = tmem + ab * (uint32_t)(kTC / 2);
                            tc_mma_tf32(d, ad, bh, idesc, 0u);                  // K columns 0..7  (32 bytes)
                            tc_mma_tf32(d, ad + 2, bh + 2, idesc, 1u);          // K columns 8..15
                            tc_commit(bar_accfull + 8 * ab);
                        }
                        __syncwarp();
                    }
                }
                if (leader) tc_commit(bar_bempty + 8 * sb);
                __syncwarp();
            }
            if (leader) tc_commit(bar_aempty);
            __syncwarp();
            ++sn;
        }
        TF_ACC(0, t_all);
        TF_FLUSH(0);
        if ((ko & 16) && lane == 0 && blockIdx.x < 160) g_tf_span[blockIdx.x][2] = tf_gtime();
    } else if (warp >= kTfEpWarps) {
        // =========================== producers ===========================
        const int ptid = tid - kTfEpWarps * 32;                         // 0..kTfPrThreads-1
        unsigned char *const tileA0 = smem, *const tileB0 = smem + kTfOffB;
        auto load3 = [&](const float *src, float &o0, float &o1, float &o2) {
            // volatile: the loads stay where they are written (ahead of the barrier wait that hides their latency)
            asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(o0) : "l"(src));
            asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(o1) : "l"(src + 1));
            asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(o2) : "l"(src + 2));
        };
        uint32_t bt = 0, sn = 0;
        TF_T0(t_allp);
        for (int task = t_begin; task < t_end;) {
            const TfSeg sg = seg_at(task);
            task = sg.next;
            const float *qc = (sg.dir ? pc2 : pc1) + (size_t)sg.b * sg.nq * 3;
            const float *cc = (sg.dir ? pc1 : pc2) + (size_t)sg.b * sg.nc * 3;
            // 128 queries of block qb0+q -> A tile q (rows past the end are all-zero); split in two so the loads of the
            // segment's first block can be in flight while the first candidate tile is produced
            auto load_a = [&](int q, float (&x)[kARows][3]) {
#pragma unroll
                for (int h = 0; h < kARows; ++h) {
                    const int i = (sg.qb0 + q) * kTQ + h * kTfPrThreads + ptid;
                    x[h][0] = 0.f; x[h][1] = 0.f; x[h][2] = 0.f;
                    if (i < sg.nq) load3(qc + (size_t)i * 3, x[h][0], x[h][1], x[h][2]);
                }
            };
            auto store_a = [&](int q, const float (&x)[kARows][3]) {
                float nmax = 0.0f;
#pragma unroll
                for (int h = 0; h < kARows; ++h) {
                    const int r = h * kTfPrThreads + ptid;
                    const int i = (sg.qb0 + q) * kTQ + r;
                    float e[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) e[k] = 0.0f;
                    if (i < sg.nq) {
                        const float nx = norm2_tf(x[h][0], x[h][1], x[h][2]);
                        nmax = fmaxf(nmax, nx);
                        float hh, l;
                        tf32_split2(x[h][0], hh, l); e[0] = hh; e[1] = hh; e[2] = l;
                        tf32_split2(x[h][1], hh, l); e[3] = hh; e[4] = hh; e[5] = l;
                        tf32_split2(x[h][2], hh, l); e[6] = hh; e[7] = hh; e[8] = l;
                        tf32_split3(nx, e[9], e[10], e[11]);
                        e[12] = 1.0f; e[13] = 1.0f; e[14] = 1.0f;
                    }
                    tf_store_row(tileA0 + (uint32_t)(q >> 1) * kTfABytes, r, e, 4u * (uint32_t)(q & 1));
                }
                // largest |q|^2 of the query cloud, for the finalize's margin (bitwise complement, atomicMin)
                const unsigned wx = __reduce_max_sync(0xffffffffu, __float_as_uint(nmax));
                if (lane == 0) atomicMin(&w.nrm[(sg.dir ? B : 0) + sg.b], ~wx);
                fence_async_proxy();
                mbar_arrive(bar_afull + 8 * q);
            };
            // 256 candidates of tile k -> B tile (two rows per thread); the loads are issued BEFORE the wait for the buffer
            auto produce_b = [&](int k) {
                const uint32_t sb = bt & 1u;
                float y[kBRows][3];
#pragma unroll
                for (int h = 0; h < kBRows; ++h) {
                    const int j = min(k * kTC + h * kTfPrThreads + ptid, sg.nc - 1);
                    load3(cc + (size_t)j * 3, y[h][0], y[h][1], y[h][2]);
                }
                { TF_T0(t0); mbar_wait_spin(bar_bempty + 8 * sb, ((bt >> 1) & 1u) ^ 1u); TF_ACC(5, t0); }
                TF_T0(t_cv);
#pragma unroll
                for (int h = 0; h < kBRows; ++h) {
                    const int r = h * kTfPrThreads + ptid;
                    const bool valid = k * kTC + r < sg.nc;
                    const float y0 = y[h][0], y1 = y[h][1], y2 = y[h][2];
                    float e[16];
                    float hh, l;
                    tf32_split2(valid ? -2.0f * y0 : 0.0f, hh, l); e[0] = hh; e[1] = l; e[2] = hh;
                    tf32_split2(valid ? -2.0f * y1 : 0.0f, hh, l); e[3] = hh; e[4] = l; e[5] = hh;
                    tf32_split2(valid ? -2.0f * y2 : 0.0f, hh, l); e[6] = hh; e[7] = l; e[8] = hh;
                    e[9] = 1.0f; e[10] = 1.0f; e[11] = 1.0f;
                    tf32_split3(valid ? norm2_tf(y0, y1, y2) : kTfBig, e[12], e[13], e[14]);
                    e[15] = 0.0f;
                    tf_store_row(tileB0 + sb * kTfBBytes, r, e);
                }
                fence_async_proxy();
                mbar_arrive(bar_bfull + 8 * sb);
                TF_ACC(7, t_cv);
                ++bt;
            };
            // query loads run one block ahead of their conversion; the first two are in flight under the first tile
            float ax[2][kARows][3];
            load_a(0, ax[0]);
            if (sg.Q > 1) load_a(1, ax[1]);
            produce_b(0);
            // the previous segment's MMAs must be done with the A tiles
            { TF_T0(t0); mbar_wait_spin(bar_aempty, (sn & 1u) ^ 1u); TF_ACC(6, t0); }
#pragma unroll
            for (int q = 0; q < kQmax; ++q) {                           // every barrier advances once per segment
                if (q < sg.Q) {
                    store_a(q, ax[q & 1]);
                    if (q + 2 < sg.Q) load_a(q + 2, ax[q & 1]);
                } else {
                    mbar_arrive(bar_afull + 8 * q);
                }
            }
            for (int k = 1; k < sg.n_ct; ++k) produce_b(k);
            ++sn;
        }
        TF_ACC(4, t_allp);
        TF_FLUSH(4);
    } else {
        // =========================== epilogue ===========================
        const int grp_id = warp >> 2;                                   // half-visits with running index % 4 == grp_id
        const int row = (warp & 3) * 32 + lane;                         // query of the block == TMEM lane
        const uint32_t lane_base = ((uint32_t)(warp & 3) * 32u) << 16;
        const uint32_t taddr = tmem + lane_base + (uint32_t)grp_id * (uint32_t)(kTC / 2);
        // this thread's running (best, second-best group minimum, best group) per query block of the segment
        float *st = reinterpret_cast<float *>(smem + kTfOffState) + (grp_id * kQmax) * kTQ * kStW + row;
        constexpr int kStQ = kTQ * kStW;                                // words per query block of one group
        uint32_t ht0 = 0;                                               // running half-visit index at the segment start
        TF_T0(t_alle);
        for (int task = t_begin; task < t_end;) {
            const TfSeg sg = seg_at(task);
            task = sg.next;
            for (int q = 0; q < sg.Q; ++q) {
                float *sq = st + q * kStQ;
                sq[0] = INFINITY; sq[kTQ] = INFINITY; sq[2 * kTQ] = INFINITY;
                reinterpret_cast<int *>(sq)[3 * kTQ] = 0; reinterpret_cast<int *>(sq)[4 * kTQ] = 0;
            }
            // half-visits of the segment in issue order: (tile k, query block q, half hf); this group takes every fourth one
            const uint32_t n_hv = (uint32_t)sg.n_ct * (uint32_t)sg.Q * 2u;
            uint32_t hv = ((uint32_t)grp_id - ht0) & 3u;                // first local index with (ht0 + hv) % 4 == grp_id
            int k = 0, q = (int)(hv >> 1);
            const int hf = (int)(hv & 1u);                              // a stride of 4 half-visits never changes the half
            while (q >= sg.Q) { q -= sg.Q; ++k; }
            for (; hv < n_hv; hv += 4u) {
                const uint32_t use = (ht0 + hv) >> 2;
                float *sq = st + q * kStQ;
                float best = sq[0], second = sq[kTQ], third = TOP3 ? sq[2 * kTQ] : INFINITY;
                int bgrp = reinterpret_cast<int *>(sq)[3 * kTQ], sgrp = TOP3 ? reinterpret_cast<int *>(sq)[4 * kTQ] : 0;
                // running three smallest group minima (strict <: the earliest group keeps a tie) and the groups of the first two
                auto group_done = [&](const float *v, int G) {
                    if (ko & 2) return;                                 // timing experiment; the branch also keeps ptxas from
                                                                        // interleaving the four reductions (which spills)
                    const float m = min32(v);
                    if (!TOP3) {
                        second = fminf(second, fmaxf(best, m));
                        if (m < best) { best = m; bgrp = G; }
                        return;
                    }
                    // nothing changes unless the group beats the running third; late in a large cloud that is rare for
                    // all 32 queries of the warp at once
                    if (!__any_sync(0xffffffffu, m < third)) return;
                    const bool p1 = m < best;
                    const float c1 = fmaxf(best, m);                    // what drops out of first place
                    const bool p2 = c1 < second;
                    third = fminf(third, fmaxf(second, c1));
                    second = fminf(second, c1);
                    sgrp = p2 ? (p1 ? bgrp : G) : sgrp;
                    best = fminf(best, m);
                    bgrp = p1 ? G : bgrp;
                };
                { TF_T0(t0); mbar_wait_spin(bar_accfull + 8 * grp_id, use & 1u); TF_ACC(9, t0); }
                tc_fence_after();
                TF_T0(t_ep);
                // 4 groups of 32 columns, two register sets: the load of group g+1 is in flight while g is reduced; the
                // accumulator goes back to the MMA warp as soon as its last load has landed
                float va[32], vb[32];
                int G0 = k * (kTC / 32) + hf * 4;
                asm volatile("mov.s32 %0, %0;" : "+r"(G0));             // pinned: otherwise recomputed under every predicate
                tc_ld32_nowait(taddr, va);
                tc_wait_ld(va);
                tc_ld32_nowait(taddr + 32u, vb);
                group_done(va, G0);
                tc_wait_ld(vb);
                tc_ld32_nowait(taddr + 64u, va);
                group_done(vb, G0 + 1);
                tc_wait_ld(va);
                tc_ld32_nowait(taddr + 96u, vb);
                group_done(va, G0 + 2);
                tc_wait_ld(vb);
                tc_fence_before();
                mbar_arrive(bar_accempty + 8 * grp_id);
                group_done(vb, G0 + 3);
                sq[0] = best; sq[kTQ] = second;
                reinterpret_cast<int *>(sq)[3 * kTQ] = bgrp;
                if (TOP3) { sq[2 * kTQ] = third; reinterpret_cast<int *>(sq)[4 * kTQ] = sgrp; }
                TF_ACC(10, t_ep);
                q += 2;
                while (q >= sg.Q) { q -= sg.Q; ++k; }
            }
            ht0 += n_hv;
            TF_T0(t_fl);
            // ---- merge the four groups' results per query and publish them (plain stores: every candidate of these
            // queries was seen by this CTA in this segment).  Group g merges query blocks g, g+4.
            asm volatile("bar.sync 1, %0;" ::"n"(kTfEpWarps * 32) : "memory");
            u64 *keys = (sg.dir ? w.colkey : w.rowkey) + (size_t)sg.b * sg.nq;
            unsigned *secs = (sg.dir ? w.colsec : w.rowsec) + (size_t)sg.b * sg.nq;
            unsigned *sgs = (sg.dir ? w.colsg : w.rowsg) + (size_t)sg.b * sg.nq;
            unsigned *ths = (sg.dir ? w.colth : w.rowth) + (size_t)sg.b * sg.nq;
            const float *all_st = reinterpret_cast<const float *>(smem + kTfOffState) + row;
            for (int qq = grp_id; qq < sg.Q; qq += kTfEpGroups) {
                const int i = (sg.qb0 + qq) * kTQ + row;
                // three smallest (value, group) keys over the four groups' top-two lists, third value also over their thirds
                u64 k1 = kKeyInit, k2 = kKeyInit;
                float t3 = INFINITY;
                auto insert = [&](u64 key) {
                    const u64 c1 = key > k1 ? key : k1;
                    k1 = key < k1 ? key : k1;
                    const u64 c2 = c1 > k2 ? c1 : k2;
                    k2 = c1 < k2 ? c1 : k2;
                    t3 = fminf(t3, __uint_as_float((unsigned)(c2 >> 32) & 0x7fffffffu));    // all-ones init -> NaN: ignored
                };
#pragma unroll
                for (int g = 0; g < kTfEpGroups; ++g) {
                    const float *sp = all_st + (g * kQmax + qq) * kStQ;
                    const int *ip = reinterpret_cast<const int *>(sp);
                    insert(((u64)__float_as_uint(fmaxf(sp[0], 0.0f)) << 32) | (unsigned)ip[3 * kTQ]);       // +inf if nothing seen
                    insert(((u64)__float_as_uint(fmaxf(sp[kTQ], 0.0f)) << 32) | (unsigned)ip[4 * kTQ]);
                    t3 = fminf(t3, fmaxf(sp[2 * kTQ], 0.0f));
                }
                if (i < sg.nq) {
                    keys[i] = k1;
                    const bool has2 = k2 != kKeyInit;                   // a single candidate group: no runner-up
                    secs[i] = has2 ? (unsigned)(k2 >> 32) : 0x7f800000u;
                    if (TOP3) {
                        sgs[i] = has2 ? (unsigned)(k2 & 0xffffffffu) : 0u;
                        ths[i] = __float_as_uint(t3);
                    }
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kTfEpWarps * 32) : "memory");
            TF_ACC(11, t_fl);
        }
        TF_ACC(8, t_alle);
        TF_FLUSH(8);
    }
    tc_fence_before();
    __syncthreads();
    if ((ko & 16) && tid == 0 && blockIdx.x < 160) g_tf_span[blockIdx.x][3] = tf_gtime();
    if (warp == kTfMmaWarp) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

// rows_per_lane tells the finalize what the sweep reported: 0 = consecutive groups + runner-up group + third value,
// -1 = consecutive groups only
int launch_tcfilter(const float *pc1, const float *pc2, int B, int N, int M, const FwdWs &w, int *rows_per_lane, cudaStream_t st) {
    const int qb1 = (N + kTQ - 1) / kTQ, qb2 = (M + kTQ - 1) / kTQ;
    const long long n_tasks = (long long)B * (qb1 + qb2);
    if (n_tasks > 0x7fffffffLL) return fail(RLG_ERR_TOO_LARGE, "rlg_chamfer_fwd: too many query blocks");
    const int sms = sm_count();
    if (sms <= 0) return fail((int)cudaErrorNoDevice, "rlg_chamfer_fwd: no CUDA device");
    // Ambiguous points (runner-up group within the margin) cost a rescan of the whole candidate cloud unless the sweep
    // tracks three groups; their share grows with the point density.  Break-even is around 4096 candidates.
    const bool top3 = getenv("RLG_TF_TOP3") ? atoi(getenv("RLG_TF_TOP3")) != 0 : (N > 4096 || M > 4096);
    *rows_per_lane = top3 ? 0 : -1;
    // per launch (a host-side attribute write, no device work): the setting is per device and this library keeps no
    // per-device state of its own
    {
        cudaError_t e = top3 ? cudaFuncSetAttribute(chamfer_tcfilter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTfSmem)
                             : cudaFuncSetAttribute(chamfer_tcfilter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTfSmem);
        if (e != cudaSuccess) { cudaGetLastError(); return fail((int)e, "chamfer_tcfilter_kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); }
    }
    const int grid = (int)(n_tasks < sms ? n_tasks : sms);
    const int units = 2 * B;
    // measured (B=32, N=M=2048): unit-aligned chunks (split 1: 8-task chunks, one segment) tie with the even split (7 tasks,
    // two segments for a third of the CTAs); at 128 units over 148 CTAs they lose badly.  The even split stays the default.
    int split = 0;
    if (getenv("RLG_TF_SPLIT")) split = atoi(getenv("RLG_TF_SPLIT"));
    if ((split == 1 && units > grid) || split < 0 || split > 2) split = 0;
    const int ko = (getenv("RLG_TF_KO") ? atoi(getenv("RLG_TF_KO")) : 0) | (getenv("RLG_TF_DEBUG") ? 16 : 0);
    cudaError_t le = top3 ? launch_pdl(chamfer_tcfilter_kernel<true>, dim3((unsigned)grid), dim3(kTfThreads), (size_t)kTfSmem, st, pc1,
                                       pc2, B, N, M, (int)n_tasks, qb1, qb2, split, w, ko)
                          : launch_pdl(chamfer_tcfilter_kernel<false>, dim3((unsigned)grid), dim3(kTfThreads), (size_t)kTfSmem, st, pc1,
                                       pc2, B, N, M, (int)n_tasks, qb1, qb2, split, w, ko);
    if (le != cudaSuccess) { cudaGetLastError(); return fail((int)le, "chamfer_tcfilter_kernel: %s", cudaGetErrorString(le)); }
    if (getenv("RLG_TF_DEBUG")) {
        cudaStreamSynchronize(st);
        unsigned long long c[16];
        cudaMemcpyFromSymbol(c, g_tf_dbg, sizeof(c));
        {
            static unsigned long long sp[160][6];
            cudaMemcpyFromSymbol(sp, g_tf_span, sizeof(sp));
            unsigned long long t0 = ~0ull, t3 = 0, setup_max = 0, loop_min = ~0ull, loop_max = 0, tail_max = 0, last_start = 0;
            for (int c = 0; c < grid && c < 160; ++c) {
                if (sp[c][0] < t0) t0 = sp[c][0];
                if (sp[c][3] > t3) t3 = sp[c][3];
                if (sp[c][0] > last_start) last_start = sp[c][0];
                if (sp[c][1] - sp[c][0] > setup_max) setup_max = sp[c][1] - sp[c][0];
                const unsigned long long lp = sp[c][2] - sp[c][1];
                if (lp < loop_min) loop_min = lp;
                if (lp > loop_max) loop_max = lp;
                if (sp[c][3] - sp[c][2] > tail_max) tail_max = sp[c][3] - sp[c][2];
            }
            if (getenv("RLG_TF_DEBUG")[0] == '2')
                for (int c = 0; c < grid && c < 160; ++c)
                    fprintf(stderr, "  cta %3d tasks %2llu from %4llu  start +%5llu  loop %6llu ns\n", c, sp[c][4], sp[c][5], sp[c][0] - t0, sp[c][2] - sp[c][1]);
            fprintf(stderr, "tcfilter span (ns): first entry -> last exit %llu | entries spread %llu | setup max %llu | MMA loop min %llu max %llu | loop end -> exit max %llu\n",
                    t3 - t0, last_start - t0, setup_max, loop_min, loop_max, tail_max);
        }
        fprintf(stderr, "tcfilter CTA0 cycles: mma total %llu wait_afull %llu wait_bfull %llu wait_accempty %llu | producer total %llu "
                "wait_bempty %llu wait_aempty %llu convert+store %llu | epilogue total %llu wait_accfull %llu ld+min %llu flush %llu\n",
                c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[8], c[9], c[10], c[11]);
    }
    return check_launch("chamfer_tcfilter_kernel");
}

}  // namespace rlg
#endif  // RLG_EXPERIMENTS
