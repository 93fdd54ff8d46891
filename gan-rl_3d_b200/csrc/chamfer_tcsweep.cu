// chamfer_tcsweep.cu -- batched Chamfer distance forward in ONE launch: the pair sweep on the tensor cores with the exact
// refinement, the few ambiguous points, the per-pair means and the batch loss all fused in.
//
// Replaces utils/losses.py:29-37 of the reference (torch.cdist -> min over both axes -> mean); the (B,N,M) matrix never
// exists.  Filter and refine, as in chamfer_filter.cu: a cheap filter value t~(i,j) ~ |x_i - y_j|^2 for EVERY pair
// decides which 32-candidate group can hold a query's nearest neighbour; the direct-form arithmetic
// (t = d0*d0; t = fma(d1,d1,t); t = fma(d2,d2,t); sqrtf -- bit-identical to ATen's direct cdist) runs only there.
//
// The filter is a K = 16 tcgen05.mma kind::tf32 contraction with error-compensated operands:
//       A row (query x)      x0h x0h x0l | x1h x1h x1l | x2h x2h x2l | nxh nxm nxl | 1   1   1   | 0
//       B row (candidate y)  Y0h Y0l Y0h | Y1h Y1l Y1h | Y2h Y2l Y2h | 1   1   1   | nyh nym nyl | 0      (Y = -2y)
//   vh = tf32(v), vl = tf32(v - vh); squared norms split three ways (exactly); only the vl x vl products are dropped.
//   D[128 queries x 128 candidates] (fp32, TMEM) = A . B^T; a query sits on one TMEM lane, so its minimum over the
//   candidates is a per-thread reduction of tcgen05.ld registers.  Both directions are the same problem with the clouds
//   swapped.
//
// Error of the filter (u = 2^-24, a = |x_i|, b = |y_j|, S = a^2 + b^2): dropped split terms <= 12 u S, accumulation
// inside the tensor core (modelled as fp32 truncation at each of 18 additions) <= 36 u S, computed norms <= 3 u S:
// |t~ - t| <= 51 u S; the direct form adds <= 10 u S.  So |t~_j - D_j| <= 61 u S_j for the direct-form value D_j.
// The refinement needs no knowledge of the candidate cloud's extent: b <= a + sqrt(t) gives
//       S_j <= Sigma(t_j) := 2 a^2 + 2 a sqrt(t_j) + t_j      (increasing in t_j),
// and from t_j <= t~_j + 51 u S_j <= t~_j + 51 u (3 a^2 + 2 t_j):   t_j <= t+(t~_j) := 1.0001 max(t~_j, 0) + 256 u a^2.
// With val = smallest group minimum (group g*, attained by candidate c) and sv = smallest minimum of any other group,
// every candidate j outside g* has D_j >= L(t~_j) := t~_j - 61 u Sigma(t+(t~_j)) with t~_j >= sv, L increasing for
// t~ >= 1e-9 a^2, and D_c <= H(val) := val + 61 u Sigma(t+(val)).  If  sv >= 1e-9 a^2  and  L(sv) > H(val) (1 + 8 u)
// (evaluated with 64 u instead of 61 u to cover the rounding of the test itself; the 8 u cover the <= 4 u relative window
// in which two squared distances share one sqrtf), the exact nearest neighbour -- also under the reference's tie rule,
// which compares the square-rooted values -- lies in group g*: 32 direct-form evaluations.  Otherwise the point is
// AMBIGUOUS (about 0.6 % of the points at N = M = 2048 on the unit sphere): with the runner-up's group and the
// third-smallest group minimum known (TOP3, large clouds) it is refined on two groups when the third is out of reach,
// else its index goes to a per-CTA list and, once the segment's queries are refined, the CTA's warps scan the whole
// candidate cloud for it (one warp per point).  Either way the outputs are independent of how the filter rounded.
//
// Tie rule: torch.min runs on the sqrt-ed matrix, and sqrtf maps up to three adjacent fp32 values onto one, so
// candidates whose squared distances differ in the last bits tie there and the LOWEST INDEX wins.  The refinement finds
// m = min t, s = sqrtf(m), the largest h with sqrtf(h) == s, and returns the lowest index with t <= h (oracle:
// ORC_TIE_FAITHFUL).
//
// Roles (one CTA per SM, persistent over (direction, cloud, 128-query block) tasks; a CTA owns a contiguous task range
// and walks it in SEGMENTS of up to 8 query blocks of one (direction, cloud) that share every converted candidate tile;
// a "visit" = one query block x one 256-candidate tile = two 128-column halves):
//   warps 0-15  epilogue: four groups of four warps; half-visit h (in issue order) goes to TMEM accumulator h % 4 and to
//               group h % 4; thread = query row = TMEM lane, 128 candidates per half-visit as four tcgen05.ld of 32
//               columns in two register sets (the load of chunk c+1 is in flight while chunk c is reduced).  Running
//               (best, second, third group minimum; best / second group) per query block live in shared memory; at the
//               end of a segment the four groups' partial states of a query are merged and the merging thread refines it.
//   warps 16-17 producers: convert 256 candidates per tile (and the segment's queries) to the split-tf32 rows, written
//               straight into K-major SWIZZLE_128B operand tiles (rows are 64 bytes: two query blocks share one tile,
//               the descriptor's start address picks the half).  Clouds of up to 2048 points are also kept raw in shared
//               memory, so the fused refinement reads its 32 candidates with LDS.128 instead of going to L2.
//   warp 18     MMA issuer (one elected lane), TMEM owner.
// Work split: with at most as many (direction, cloud) units as CTAs every unit is cut into equal chunks, one per CTA, so
// a CTA runs ONE segment (its fixed costs -- operand conversion, pipeline fill, merge, refinement -- are paid once);
// otherwise the task list is split evenly and contiguously.
// Means and loss: after a segment's queries are refined the CTA sums their distances per query block in a fixed order,
// publishes the partial sums and counts the block as done; whoever completes a (direction, cloud) unit reduces its
// partials to the mean, and whoever completes the last unit reduces the 2B means to the loss -- all in fixed orders, so
// the results are bit-reproducible.  The counters start and end all-ones (the workspace invariant).
#include "common.cuh"
#include "tcgen05.cuh"
#include <math.h>

namespace rlg {

// experiment knob (tools/ A-B builds only; the default is the product configuration): 1 or 2 MMA issuer warps.  Measured
// at B=32, N=M=2048: no difference (28.6 vs 28.7 us for the filter sweep); packing both candidate tiles into one
// 128-byte-row tile was measured too and costs 2-2.7 us (operand fetches and producer stores collide), so it is gone.
#ifndef RLG_TS_ISSUERS
#define RLG_TS_ISSUERS 1
#endif
// experiments only (tools/ A-B builds, wrong results): knock out parts of the tail to time them -- 1: the per-query
// refinement, 2: the ambiguous scans, 4: the means / loss reduction
#if !defined(RLG_EXPERIMENTS) || !defined(RLG_TS_KO)
#undef RLG_TS_KO
#define RLG_TS_KO 0
#endif
constexpr int kTQ = 128;                       // queries per block (UMMA M)
constexpr int kTC = 256;                       // candidates per tile (two UMMA N = 128 halves)
constexpr int kQmax = 8;                       // query blocks that share one pass over the candidate tiles
constexpr int kEpWarps = 16, kPrWarps = 2, kMmaWarps = RLG_TS_ISSUERS;
constexpr int kEpThreads = kEpWarps * 32;
constexpr int kMmaWarp0 = kEpWarps + kPrWarps;
constexpr int kThreads = (kEpWarps + kPrWarps + kMmaWarps) * 32;
constexpr int kPrThreads = kPrWarps * 32;
constexpr int kARows = kTQ / kPrThreads, kBRows = kTC / kPrThreads;            // rows per producer thread
constexpr uint32_t kABytes = kTQ * 128, kBBytes = kTC * 128;
// A rows carry K = 16 tf32 = 64 bytes, half of a SWIZZLE_128B row: two query blocks share one 16 KB tile (block q sits in
// 16-byte chunks 4*(q&1) .. 4*(q&1)+3 of tile q>>1; the descriptor's start address selects the half)
constexpr uint32_t kOffB = (kQmax / 2) * kABytes;                             // two candidate tiles follow the A tiles
constexpr int kSlots = 4;                                                     // partial states per query: one per epilogue group
// raw candidate copy: a 32-candidate group takes 96 floats; a stride of 100 floats (25 x 16 bytes, odd) spreads the
// groups over the shared-memory banks -- with 96 every thread's LDS.128 would start in the same bank quad (8-way conflict)
constexpr int kRawStride = 100;
constexpr uint32_t kOffState = kOffB + 2 * kBBytes;                           // [slot][kQmax][kStW][128]
// state words per (slot, query block, row): best, second, best group [, third, second group]
constexpr int kWBest = 0, kWSecond = 1, kWBgrp = 2, kWThird = 3, kWSgrp = 4;
template <bool TOP3> struct SweepCfg {
    static constexpr int kStW = TOP3 ? 5 : 3;
    static constexpr uint32_t kStateBytes = (uint32_t)kSlots * kQmax * kTQ * kStW * 4u;
    // raw copy of the segment's candidate cloud for the fused refinement (clouds of up to kRawMax points).  With TOP3
    // (clouds beyond 4096 points) there is nothing to stage: the refinement reads its candidates from L2.
    static constexpr int kRawMax = TOP3 ? 0 : 2048;
    static constexpr uint32_t kOffRaw = kOffState + kStateBytes;
    static constexpr uint32_t kRawBytes = (uint32_t)(kRawMax / kGroup) * kRawStride * 4u;
    static constexpr uint32_t kOffDist = kOffRaw + kRawBytes;                 // float [kQmax*128]: the segment's distances
    static constexpr uint32_t kOffAmb = kOffDist + kQmax * kTQ * 4u;          // u16 [kQmax*128]: the segment's ambiguous queries
    static constexpr uint32_t kOffBar = kOffAmb + kQmax * kTQ * 2u + 2u * kEpWarps * 4u;   // + scratch of scan_query_cta
    static constexpr uint32_t kSmem = kOffBar + 256 + 1024;                   // + barriers + alignment slack
};
constexpr float kBig = 1.0e30f;
constexpr float kU = 1.0f / 16777216.0f;                                      // 2^-24
constexpr float kErr = 64.0f * kU;                                             // filter-vs-direct error per unit of Sigma (header)

__device__ __forceinline__ float tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void split2(float v, float &h, float &l) {
    h = tf32_rn(v);
    l = tf32_rn(v - h);                                // v - h is exact
}
__device__ __forceinline__ void split3(float v, float &h, float &m, float &l) {
    h = tf32_rn(v);
    const float r = v - h;                             // exact
    m = tf32_rn(r);
    l = r - m;                                         // exact, at most 3 significant bits
}
// row `row` of a K-major SW128 tile: 16 floats as four 16-byte chunks starting at chunk `chunk0`
__device__ __forceinline__ void store_row(unsigned char *tile, int row, const float (&e)[16], uint32_t chunk0) {
    const uint32_t base = (uint32_t)row * 128u, x = (uint32_t)row & 7u;
#pragma unroll
    for (uint32_t c = 0; c < 4; ++c)
        *reinterpret_cast<float4 *>(tile + base + (((chunk0 + c) ^ x) << 4)) = make_float4(e[4 * c], e[4 * c + 1], e[4 * c + 2], e[4 * c + 3]);
}
__device__ __forceinline__ float nrm2(float x, float y, float z) { return fmaf(z, z, fmaf(y, y, x * x)); }

// minimum of 32 values, as a tree of 3-input minima (NaN operands are dropped)
__device__ __forceinline__ float min32(const float *v) {
    float m[11];
#pragma unroll
    for (int i = 0; i < 10; ++i) m[i] = min3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
    m[10] = fminf(v[30], v[31]);
    const float n0 = min3(m[0], m[1], m[2]), n1 = min3(m[3], m[4], m[5]), n2 = min3(m[6], m[7], m[8]);
    return min3(min3(n0, n1, n2), m[9], m[10]);
}

// ---- exact arithmetic shared by the fused refinement and the tail kernel --------------------------------------------
// direct-form squared distances of (qx,qy,qz) to candidates base .. base+31 of cloud c (indices past the end repeat the
// last point: a real candidate with the highest index, so it can never win a tie it should not)
__device__ __forceinline__ void eval_group(const float *__restrict__ c, int nc, int base, float qx, float qy, float qz,
                                           float (&t)[32]) {
    const float *p = c + 3 * (size_t)base;
    if (base + 32 <= nc && (reinterpret_cast<uintptr_t>(p) & 15u) == 0) {
        const float4 *p4 = reinterpret_cast<const float4 *>(p);
#pragma unroll
        for (int h = 0; h < 4; ++h) {                  // 8 candidates = 24 floats = six 16-byte loads in flight
            float4 v[6];
#pragma unroll
            for (int e = 0; e < 6; ++e) v[e] = __ldg(p4 + h * 6 + e);
            float f[24];
#pragma unroll
            for (int e = 0; e < 6; ++e) { f[4 * e] = v[e].x; f[4 * e + 1] = v[e].y; f[4 * e + 2] = v[e].z; f[4 * e + 3] = v[e].w; }
#pragma unroll
            for (int k = 0; k < 8; ++k) t[h * 8 + k] = sqdist(qx, qy, qz, f[3 * k], f[3 * k + 1], f[3 * k + 2]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const int j = min(base + k, nc - 1);
            t[k] = sqdist(qx, qy, qz, __ldg(c + 3 * (size_t)j), __ldg(c + 3 * (size_t)j + 1), __ldg(c + 3 * (size_t)j + 2));
        }
    }
}
// the same from the raw shared-memory copy of the candidate cloud (group g at raw + g * kRawStride, always 16-byte
// aligned; rows past the end of the cloud hold copies of the last point)
__device__ __forceinline__ void eval_group_smem(const float *raw, int g, float qx, float qy, float qz, float (&t)[32]) {
    const float4 *p4 = reinterpret_cast<const float4 *>(raw + g * kRawStride);
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        float4 v[6];
#pragma unroll
        for (int e = 0; e < 6; ++e) v[e] = p4[h * 6 + e];
        float f[24];
#pragma unroll
        for (int e = 0; e < 6; ++e) { f[4 * e] = v[e].x; f[4 * e + 1] = v[e].y; f[4 * e + 2] = v[e].z; f[4 * e + 3] = v[e].w; }
#pragma unroll
        for (int k = 0; k < 8; ++k) t[h * 8 + k] = sqdist(qx, qy, qz, f[3 * k], f[3 * k + 1], f[3 * k + 2]);
    }
}
// lowest k with t[k] <= h (32 if none)
__device__ __forceinline__ int first_le(const float (&t)[32], float h) {
    int kk = 32;
#pragma unroll
    for (int k = 31; k >= 0; --k) kk = (t[k] <= h) ? k : kk;
    return kk;
}

struct SweepOut {
    float *d1, *d2;                 // (B,N), (B,M)
    int32_t *i1, *i2;
    float *z1, *z2;                 // optional (B,N,3)/(B,M,3) gradient buffers to zero-fill
    float *mean1, *mean2, *loss;    // (B), (B), scalar; mean1/mean2 nullable as a pair, loss nullable
    float w1, w2;                   // loss = sum_b w1 * mean1[b] + w2 * mean2[b]
    double *partial;                // (2B, pq): sums of the distances per 128-query block
    unsigned *unit_cnt;             // (2B): finished query blocks per (direction, cloud), all-ones = none
    unsigned *global_cnt;           // (1):  finished units, all-ones = none
    int pq;                         // query blocks per unit in `partial` (the longer direction's count)
    // diagnostic (RLG_CHAMFER_FILTER_ONLY): the merged filter results per query instead of the refinement
    u64 *key1, *key2;               // (B,N), (B,M): smallest group minimum << 32 | its group
    unsigned *sec1, *sec2;          // (B,N), (B,M): second-smallest group minimum (+inf: none)
};

struct Seg { int dir, b, qb0, Q, nq, nc, n_ct, next; };

// ---- fused refinement of one query ------------------------------------------------------------------------------
// k1 = (smallest group minimum, its group), k2 = (second smallest, its group), t3 = third smallest group minimum
template <bool TOP3>
__device__ __forceinline__ void refine_query(const float *__restrict__ qc, const float *__restrict__ cc, const float *raw,
                                             int nc, int i, u64 k1, u64 k2, float t3, float *__restrict__ dout,
                                             int32_t *__restrict__ iout, float *__restrict__ zero, int local,
                                             float *s_dist, unsigned short *s_amb, unsigned *s_namb) {
    const float qx = __ldg(qc + 3 * (size_t)i), qy = __ldg(qc + 3 * (size_t)i + 1), qz = __ldg(qc + 3 * (size_t)i + 2);
    if (zero != nullptr) { zero[3 * (size_t)i] = 0.0f; zero[3 * (size_t)i + 1] = 0.0f; zero[3 * (size_t)i + 2] = 0.0f; }
    const float val = __uint_as_float((unsigned)(k1 >> 32));
    const int grp = (int)(unsigned)(k1 & 0xffffffffu);
    const bool has2 = k2 != kKeyInit;                                   // a single candidate group: nothing to confuse
    const float sv = has2 ? __uint_as_float((unsigned)(k2 >> 32)) : INFINITY;
    const int sgrp = (int)(unsigned)(k2 & 0xffffffffu);
    // the margin test of the header; any NaN (non-finite input) makes a comparison false -> ambiguous -> full scan
    const float a2 = nrm2(qx, qy, qz);
    auto sigma = [&](float tt) {
        const float a = sqrtf(a2);
        const float tp = fmaf(fmaxf(tt, 0.0f), 1.0001f, 256.0f * kU * a2);
        return fmaf(2.0f * a, sqrtf(tp), fmaf(2.0f, a2, tp));
    };
    // cheap sufficient test first (2 a sqrt(t) <= a^2 + t, so Sigma <= 3 a^2 + 2 t: no square roots); the sharper bound
    // only for the few queries that fail it
    const float hi_cheap = fmaf(kErr, fmaf(2.02f, fmaxf(val, 0.0f), 3.01f * a2), val) * (1.0f + 8.0f * kU);
    float hi_val = 0.0f;
    bool have_hi = false;
    auto out_of_reach = [&](float v) {                                  // no candidate with a filter value >= v can win
        if (v == INFINITY) return true;
        if (!(v >= 1.0e-9f * a2)) return false;
        if (fmaf(-kErr, fmaf(2.02f, v, 3.01f * a2), v) > hi_cheap) return true;
        if (!have_hi) { hi_val = fmaf(kErr, sigma(val), val) * (1.0f + 8.0f * kU); have_hi = true; }
        return fmaf(-kErr, sigma(v), v) > hi_val;
    };
    const bool clear = out_of_reach(sv);
    const bool two = TOP3 && !clear && has2 && out_of_reach(t3);
    bool amb = !clear && !two;
    float m = INFINITY;
    float t[32];
    auto eval = [&](int g) {
        if (raw != nullptr) eval_group_smem(raw, g, qx, qy, qz, t);
        else eval_group(cc, nc, g * kGroup, qx, qy, qz, t);
    };
    if (!amb) {
        eval(grp);
        m = min32(t);
        if (TOP3 && two) {
            eval(sgrp);
            m = fminf(m, min32(t));
        }
        amb = !(m < INFINITY);                                          // non-finite input: the full scan mirrors torch.min
    }
    if (amb) {
        s_amb[atomicAdd(s_namb, 1u)] = (unsigned short)local;
        return;
    }
    const float s = __fsqrt_rn(m);
    const float h = sqrt_window_top(m, s);
    int bj = 0x7fffffff;
    if (TOP3 && two) {
        const int k2nd = first_le(t, h);                                // t holds the second group
        if (k2nd < 32) bj = sgrp * kGroup + k2nd;
        eval(grp);
    }
    const int k = first_le(t, h);
    if (k < 32) bj = min(bj, grp * kGroup + k);
    dout[i] = s;
    iout[i] = min(bj, nc - 1);
    s_dist[local] = s;
}

// Full exact scan of the candidate cloud for one ambiguous query by one warp, mirroring torch.min on the sqrt-ed row:
// the first NaN wins, otherwise the lowest index among the candidates that share the smallest sqrtf.  `at(p, x, y, z)`
// fetches candidate position p (positions past the end of the cloud repeat the last point).
template <typename At>
__device__ __forceinline__ void scan_query(At at, int npos, int nc, int lane, float qx, float qy, float qz, float &dist, int &bj) {
    float lm = INFINITY;
    int nanj = 0x7fffffff;
    int j = lane;
    for (; j + 96 < npos; j += 128) {                                   // four candidates in flight per lane
        float t[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { float x, y, z; at(j + 32 * e, x, y, z); t[e] = sqdist(qx, qy, qz, x, y, z); }
#pragma unroll
        for (int e = 3; e >= 0; --e) if (t[e] != t[e]) nanj = min(nanj, j + 32 * e);
        lm = fminf(fminf(lm, t[0]), fminf(t[1], fminf(t[2], t[3])));
    }
    for (; j < npos; j += 32) {
        float x, y, z; at(j, x, y, z);
        const float t = sqdist(qx, qy, qz, x, y, z);
        if (t != t) nanj = min(nanj, j);
        lm = fminf(lm, t);
    }
    nanj = __reduce_min_sync(0xffffffffu, nanj);
    const float m = __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(lm)));   // t >= 0: orders as unsigned
    if (nanj != 0x7fffffff) {
        dist = __uint_as_float(0x7fc00000u);
        bj = min(nanj, nc - 1);
        return;
    }
    dist = __fsqrt_rn(m);
    const float h = m < INFINITY ? sqrt_window_top(m, dist) : m;
    int fj = 0x7fffffff;
    j = lane;
    for (; j + 96 < npos; j += 128) {
        float t[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) { float x, y, z; at(j + 32 * e, x, y, z); t[e] = sqdist(qx, qy, qz, x, y, z); }
#pragma unroll
        for (int e = 3; e >= 0; --e) if (t[e] <= h) fj = min(fj, j + 32 * e);
    }
    for (; j < npos; j += 32) {
        float x, y, z; at(j, x, y, z);
        if (sqdist(qx, qy, qz, x, y, z) <= h) fj = min(fj, j);
    }
    bj = __reduce_min_sync(0xffffffffu, fj);
    bj = bj == 0x7fffffff ? 0 : min(bj, nc - 1);
}

// The same scan by ALL epilogue threads for one query, candidates read from global memory (clouds too large to stage):
// a single warp walking a 16384-point cloud through L2 is latency-bound for ~40 us; 512 threads with eight loads in
// flight each take ~3 us.  s_red: 2 x kEpWarps words of scratch.  Every epilogue thread must call it (named barrier 1).
__device__ __forceinline__ void scan_query_cta(const float *__restrict__ cc, int nc, int etid, float qx, float qy, float qz,
                                               unsigned *s_red, float &dist, int &bj) {
    const int lane = etid & 31, w = etid >> 5;
    auto block_min = [&](unsigned v0, unsigned v1, unsigned &r0, unsigned &r1) {
        v0 = __reduce_min_sync(0xffffffffu, v0);
        v1 = __reduce_min_sync(0xffffffffu, v1);
        asm volatile("bar.sync 1, %0;" ::"n"(kEpThreads) : "memory");   // previous readers of s_red are done
        if (lane == 0) { s_red[w] = v0; s_red[kEpWarps + w] = v1; }
        asm volatile("bar.sync 1, %0;" ::"n"(kEpThreads) : "memory");
        r0 = s_red[0]; r1 = s_red[kEpWarps];
#pragma unroll
        for (int k = 1; k < kEpWarps; ++k) { r0 = min(r0, s_red[k]); r1 = min(r1, s_red[kEpWarps + k]); }
    };
    auto dist2 = [&](int j) {
        const float *c = cc + 3 * (size_t)j;
        return sqdist(qx, qy, qz, __ldg(c), __ldg(c + 1), __ldg(c + 2));
    };
    float lm = INFINITY;
    int nanj = 0x7fffffff;
    int j = etid;
    for (; j + 7 * kEpThreads < nc; j += 8 * kEpThreads) {
        float t[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) t[e] = dist2(j + e * kEpThreads);
#pragma unroll
        for (int e = 7; e >= 0; --e) { if (t[e] != t[e]) nanj = min(nanj, j + e * kEpThreads); lm = fminf(lm, t[e]); }
    }
    for (; j < nc; j += kEpThreads) {
        const float t = dist2(j);
        if (t != t) nanj = min(nanj, j);
        lm = fminf(lm, t);
    }
    unsigned mb, nb;
    block_min(__float_as_uint(lm), (unsigned)nanj, mb, nb);             // t >= 0: orders as unsigned
    if (nb != 0x7fffffffu) {
        dist = __uint_as_float(0x7fc00000u);
        bj = (int)nb;
        return;
    }
    const float m = __uint_as_float(mb);
    dist = __fsqrt_rn(m);
    const float h = m < INFINITY ? sqrt_window_top(m, dist) : m;
    int fj = 0x7fffffff;
    j = etid;
    for (; j + 7 * kEpThreads < nc; j += 8 * kEpThreads) {
        float t[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) t[e] = dist2(j + e * kEpThreads);
#pragma unroll
        for (int e = 7; e >= 0; --e) if (t[e] <= h) fj = min(fj, j + e * kEpThreads);
    }
    for (; j < nc; j += kEpThreads)
        if (dist2(j) <= h) fj = min(fj, j);
    unsigned fb, dummy;
    block_min((unsigned)fj, 0u, fb, dummy);
    bj = fb == 0x7fffffffu ? 0 : (int)fb;
}

// TOP3: also track the runner-up's group and the third-smallest group minimum (5 more issue slots per 32 candidates;
// pays off when ambiguous points would otherwise rescan a large candidate cloud -- see launch_tcsweep)
template <bool TOP3>
__global__ void __launch_bounds__(kThreads, 1)
chamfer_tcsweep_kernel(const float *__restrict__ pc1, const float *__restrict__ pc2, int B, int N, int M, int n_tasks,
                       int qb1, int qb2, int split, SweepOut o, int refine) {
    using Cfg = SweepCfg<TOP3>;
    constexpr int kStW = Cfg::kStW;
    extern __shared__ unsigned char ts_smem_raw[];
    __shared__ unsigned s_namb;                                         // ambiguous queries of the current segment
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t pad = (1024u - (smem_u32(ts_smem_raw) & 1023u)) & 1023u;
    unsigned char *smem = ts_smem_raw + pad;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bars = sbase + Cfg::kOffBar;
    const uint32_t bar_afull = bars, bar_aempty = bars + 8 * kQmax;                     // [kQmax], [1]
    const uint32_t bar_bfull = bar_aempty + 8, bar_bempty = bar_bfull + 16;             // [2] each
    const uint32_t bar_accfull = bar_bempty + 16, bar_accempty = bar_accfull + 32;      // [4] each
    const uint32_t bar_rawfull = bar_accempty + 32, bar_rawempty = bar_rawfull + 8;     // [1] each
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + Cfg::kOffBar + 8 * (kQmax + 15));

    if (tid == 0) {
        for (int q = 0; q < kQmax; ++q) mbar_init(bar_afull + 8 * q, kPrThreads);
        mbar_init(bar_aempty, kMmaWarps);
        for (int k = 0; k < 2; ++k) { mbar_init(bar_bfull + 8 * k, kPrThreads); mbar_init(bar_bempty + 8 * k, kMmaWarps); }
        for (int k = 0; k < 4; ++k) { mbar_init(bar_accfull + 8 * k, 1); mbar_init(bar_accempty + 8 * k, (kEpWarps / 4) * 32); }
        mbar_init(bar_rawfull, kPrThreads);
        mbar_init(bar_rawempty, kEpThreads);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_launch_dependents();
    pdl_wait();                                   // clouds and workspace may come from the kernels right before this one

    // Tasks (direction, cloud, 128-query block) in direction-major order; a CTA owns a contiguous range and walks it in
    // SEGMENTS of up to kQmax query blocks of one (direction, cloud): they share every converted candidate tile.
    //   split 0: even contiguous split of the task list (a CTA may straddle two clouds: two segments)
    //   split 1: at most as many (direction, cloud) units as CTAs: every unit is cut into equal chunks, one per CTA
    const int G = (int)gridDim.x, cta = (int)blockIdx.x, U = 2 * B;
    int t_begin, t_end;
    if (split == 1) {
        const int c_lo = G / U, extra = G - c_lo * U;                // the first `extra` units get one chunk more
        int u, ci, c;
        if (cta < extra * (c_lo + 1)) { u = cta / (c_lo + 1); ci = cta - u * (c_lo + 1); c = c_lo + 1; }
        else { const int r = cta - extra * (c_lo + 1); u = extra + r / c_lo; ci = r - (u - extra) * c_lo; c = c_lo; }
        const int T = u < B ? qb1 : qb2, s0 = u < B ? u * qb1 : B * qb1 + (u - B) * qb2;
        t_begin = s0 + (int)((long long)ci * T / c);
        t_end = s0 + (int)((long long)(ci + 1) * T / c);
    } else {
        const int per = n_tasks / G, rem = n_tasks - per * G;
        t_begin = cta * per + min(cta, rem);
        t_end = t_begin + per + (cta < rem ? 1 : 0);
    }
    auto seg_at = [&](int task) {
        Seg sg;
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
    // the raw candidate copy of a segment is kept only for clouds that fit (and only when the refinement runs)
    auto stages_raw = [&](const Seg &sg) { return Cfg::kRawMax > 0 && refine && sg.nc <= Cfg::kRawMax; };

    if (warp >= kMmaWarp0) {
        // =========================== MMA issuer(s) ===========================
        // half-visits in issue order: (tile k, query block q, half hf), running index gt = 2 * visit + hf; accumulator
        // gt % 4 == epilogue group gt % 4 (with two issuers, issuer p issues the halves hf == p)
        const uint32_t me = (uint32_t)(warp - kMmaWarp0);
        const bool leader = elect_one();
        const uint32_t idesc = umma_idesc_tf32(kTQ, kTC / 2);
        const uint64_t ad0 = umma_desc(sbase), bd0 = umma_desc(sbase + kOffB);
        const uint64_t a_inc = (uint64_t)(kABytes >> 4), b_inc = (uint64_t)(kBBytes >> 4);   // 16-byte units
        uint32_t vis = 0, bt = 0, sn = 0;
        for (int task = t_begin; task < t_end;) {
            const Seg sg = seg_at(task);
            task = sg.next;
            for (int k = 0; k < sg.n_ct; ++k, ++bt) {
                const uint32_t sb = bt & 1u;
                mbar_wait_spin(bar_bfull + 8 * sb, (bt >> 1) & 1u);
                const uint64_t bt0 = bd0 + (uint64_t)sb * b_inc;
                for (int q = 0; q < sg.Q; ++q, ++vis) {
                    if (k == 0) mbar_wait_spin(bar_afull + 8 * q, sn & 1u);
#pragma unroll
                    for (uint32_t hf = 0; hf < 2; ++hf) {
                        if (kMmaWarps == 2 && hf != me) continue;
                        const uint32_t gt = 2u * vis + hf, ab = gt & 3u;
                        mbar_wait_spin(bar_accempty + 8 * ab, ((gt >> 2) & 1u) ^ 1u);
                        tc_fence_after();
                        if (leader) {
                            const uint64_t ad = ad0 + (uint64_t)(q >> 1) * a_inc + (uint64_t)((q & 1) * 4);
                            const uint64_t bh = bt0 + (uint64_t)hf * (b_inc >> 1);
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
    } else if (warp >= kEpWarps) {
        // =========================== producers ===========================
        const int ptid = tid - kEpWarps * 32;                           // 0..kPrThreads-1
        unsigned char *const tileA0 = smem, *const tileB0 = smem + kOffB;
        float *const raw = reinterpret_cast<float *>(smem + Cfg::kOffRaw);
        auto load3 = [&](const float *src, float &o0, float &o1, float &o2) {
            // volatile: the loads stay where they are written (ahead of the barrier wait that hides their latency)
            asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(o0) : "l"(src));
            asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(o1) : "l"(src + 1));
            asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(o2) : "l"(src + 2));
        };
        uint32_t bt = 0, sn = 0;
        for (int task = t_begin; task < t_end;) {
            const Seg sg = seg_at(task);
            task = sg.next;
            const float *qc = (sg.dir ? pc2 : pc1) + (size_t)sg.b * sg.nq * 3;
            const float *cc = (sg.dir ? pc1 : pc2) + (size_t)sg.b * sg.nc * 3;
            const bool keep_raw = stages_raw(sg);
            // 128 queries of block qb0+q -> A tile q (rows past the end are all-zero)
            auto load_a = [&](int q, float (&x)[kARows][3]) {
#pragma unroll
                for (int h = 0; h < kARows; ++h) {
                    const int i = (sg.qb0 + q) * kTQ + h * kPrThreads + ptid;
                    x[h][0] = 0.f; x[h][1] = 0.f; x[h][2] = 0.f;
                    if (i < sg.nq) load3(qc + (size_t)i * 3, x[h][0], x[h][1], x[h][2]);
                }
            };
            auto store_a = [&](int q, const float (&x)[kARows][3]) {
#pragma unroll
                for (int h = 0; h < kARows; ++h) {
                    const int r = h * kPrThreads + ptid;
                    const int i = (sg.qb0 + q) * kTQ + r;
                    float e[16];
#pragma unroll
                    for (int k = 0; k < 16; ++k) e[k] = 0.0f;
                    if (i < sg.nq) {
                        float hh, l;
                        split2(x[h][0], hh, l); e[0] = hh; e[1] = hh; e[2] = l;
                        split2(x[h][1], hh, l); e[3] = hh; e[4] = hh; e[5] = l;
                        split2(x[h][2], hh, l); e[6] = hh; e[7] = hh; e[8] = l;
                        split3(nrm2(x[h][0], x[h][1], x[h][2]), e[9], e[10], e[11]);
                        e[12] = 1.0f; e[13] = 1.0f; e[14] = 1.0f;
                    }
                    store_row(tileA0 + (uint32_t)(q >> 1) * kABytes, r, e, 4u * (uint32_t)(q & 1));
                }
                fence_async_proxy();
                mbar_arrive(bar_afull + 8 * q);
            };
            // 256 candidates of tile k -> B tile (four rows per thread); the loads are issued BEFORE the wait for the buffer
            auto produce_b = [&](int k) {
                const uint32_t sb = bt & 1u;
                float y[kBRows][3];
#pragma unroll
                for (int h = 0; h < kBRows; ++h) {
                    const int j = min(k * kTC + h * kPrThreads + ptid, sg.nc - 1);
                    load3(cc + (size_t)j * 3, y[h][0], y[h][1], y[h][2]);
                }
                mbar_wait_spin(bar_bempty + 8 * sb, ((bt >> 1) & 1u) ^ 1u);
#pragma unroll
                for (int h = 0; h < kBRows; ++h) {
                    const int r = h * kPrThreads + ptid;
                    const bool valid = k * kTC + r < sg.nc;
                    const float y0 = y[h][0], y1 = y[h][1], y2 = y[h][2];
                    if (Cfg::kRawMax > 0 && keep_raw) {         // rows past the end hold copies of the last point
                        const int jj = k * kTC + r;
                        float *rr = raw + (jj >> 5) * kRawStride + 3 * (jj & 31);
                        rr[0] = y0; rr[1] = y1; rr[2] = y2;
                    }
                    float e[16];
                    float hh, l;
                    split2(valid ? -2.0f * y0 : 0.0f, hh, l); e[0] = hh; e[1] = l; e[2] = hh;
                    split2(valid ? -2.0f * y1 : 0.0f, hh, l); e[3] = hh; e[4] = l; e[5] = hh;
                    split2(valid ? -2.0f * y2 : 0.0f, hh, l); e[6] = hh; e[7] = l; e[8] = hh;
                    e[9] = 1.0f; e[10] = 1.0f; e[11] = 1.0f;
                    split3(valid ? nrm2(y0, y1, y2) : kBig, e[12], e[13], e[14]);
                    e[15] = 0.0f;
                    store_row(tileB0 + sb * kBBytes, r, e, 0u);
                }
                fence_async_proxy();
                mbar_arrive(bar_bfull + 8 * sb);
                ++bt;
            };
            // query loads run one block ahead of their conversion; the first two are in flight under the first tile
            float ax[2][kARows][3];
            load_a(0, ax[0]);
            if (sg.Q > 1) load_a(1, ax[1]);
            // the previous segment's refinement must be done with the raw copy (one buffer: a CTA rarely has two segments)
            if (Cfg::kRawMax > 0) mbar_wait_spin(bar_rawempty, (sn & 1u) ^ 1u);
            produce_b(0);
            mbar_wait_spin(bar_aempty, (sn & 1u) ^ 1u);                 // the previous segment's MMAs are done with the A tiles
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
            if (Cfg::kRawMax > 0) mbar_arrive(bar_rawfull);             // the segment's raw copy is complete
            ++sn;
        }
    } else {
        // =========================== epilogue ===========================
        const int slot = warp >> 2;                                     // epilogue group == accumulator == partial-state slot
        const int row = (warp & 3) * 32 + lane;                         // query of the block == TMEM lane
        const uint32_t lane_base = ((uint32_t)(warp & 3) * 32u) << 16;
        const uint32_t taddr = tmem + lane_base + (uint32_t)slot * (uint32_t)(kTC / 2);
        const uint32_t my_full = bar_accfull + 8 * slot, my_empty = bar_accempty + 8 * slot;
        const int hf = slot & 1;                                        // this group's half of a visit
        float *st = reinterpret_cast<float *>(smem + kOffState) + (slot * kQmax) * kTQ * kStW + row;
        float *const s_dist = reinterpret_cast<float *>(smem + Cfg::kOffDist);
        unsigned short *const s_amb = reinterpret_cast<unsigned short *>(smem + Cfg::kOffAmb);
        const float *const raw_base = reinterpret_cast<const float *>(smem + Cfg::kOffRaw);
        constexpr int kStQ = kTQ * kStW;                                // words per query block of one slot
        uint32_t vis0 = 0, sn = 0;                                      // visits issued before this segment, segment index
        for (int task = t_begin; task < t_end;) {
            const Seg sg = seg_at(task);
            task = sg.next;
            for (int q = 0; q < sg.Q; ++q) {
                float *sq = st + q * kStQ;
                sq[kWBest * kTQ] = INFINITY; sq[kWSecond * kTQ] = INFINITY;
                reinterpret_cast<int *>(sq)[kWBgrp * kTQ] = 0;
                if (TOP3) { sq[kWThird * kTQ] = INFINITY; reinterpret_cast<int *>(sq)[kWSgrp * kTQ] = 0; }
            }
            // this group's half-visits of the segment: half hf of the visits v with (vis0 + v) % 2 == slot / 2
            const uint32_t n_vis = (uint32_t)sg.n_ct * (uint32_t)sg.Q;
            uint32_t v = (((uint32_t)slot >> 1) - vis0) & 1u;
            int k = 0, q = (int)v;
            while (q >= sg.Q) { q -= sg.Q; ++k; }
            for (; v < n_vis; v += 2u) {
                const uint32_t use = (2u * (vis0 + v) + (uint32_t)hf) >> 2;        // how often this accumulator was used before
                float *sq = st + q * kStQ;
                float best = sq[kWBest * kTQ], second = sq[kWSecond * kTQ], third = TOP3 ? sq[kWThird * kTQ] : INFINITY;
                int bgrp = reinterpret_cast<int *>(sq)[kWBgrp * kTQ], sgrp = TOP3 ? reinterpret_cast<int *>(sq)[kWSgrp * kTQ] : 0;
                // running three smallest group minima (strict <: the earliest group keeps a tie) and the groups of the first two
                auto group_done = [&](const float *vv, int Gc) {
                    const float m = min32(vv);
                    if (!TOP3) {
                        second = fminf(second, fmaxf(best, m));
                        if (m < best) { best = m; bgrp = Gc; }
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
                    sgrp = p2 ? (p1 ? bgrp : Gc) : sgrp;
                    best = fminf(best, m);
                    bgrp = p1 ? Gc : bgrp;
                };
                mbar_wait_spin(my_full, use & 1u);
                tc_fence_after();
                // 4 chunks of 32 columns, two register sets: the load of chunk c+1 is in flight while c is reduced; the
                // accumulator goes back to the issuer as soon as its last load has landed
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
                mbar_arrive(my_empty);
                group_done(vb, G0 + 3);
                sq[kWBest * kTQ] = best; sq[kWSecond * kTQ] = second;
                reinterpret_cast<int *>(sq)[kWBgrp * kTQ] = bgrp;
                if (TOP3) { sq[kWThird * kTQ] = third; reinterpret_cast<int *>(sq)[kWSgrp * kTQ] = sgrp; }
                q += 2;
                while (q >= sg.Q) { q -= sg.Q; ++k; }
            }
            vis0 += n_vis;
            // ---- merge the four partial states per query; the merging thread refines the query (every candidate of
            // these queries was seen by this CTA in this segment).  Group s merges query blocks s and s + 4.
            asm volatile("bar.sync 1, %0;" ::"n"(kEpThreads) : "memory");
            const float *all_st = reinterpret_cast<const float *>(smem + kOffState) + row;
            u64 mk1[2], mk2[2];
            float mt3[2];
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int qq = slot + 4 * r;
                // three smallest (value, group) keys over the four slots' top-two lists, third value also over their thirds
                u64 k1 = kKeyInit, k2 = kKeyInit;
                float t3 = INFINITY;
                auto insert = [&](u64 key) {
                    const u64 c1 = key > k1 ? key : k1;
                    k1 = key < k1 ? key : k1;
                    const u64 c2 = c1 > k2 ? c1 : k2;
                    k2 = c1 < k2 ? c1 : k2;
                    t3 = fminf(t3, __uint_as_float((unsigned)(c2 >> 32) & 0x7fffffffu));    // all-ones init -> NaN: ignored
                };
                if (qq < sg.Q) {
#pragma unroll
                    for (int g = 0; g < kSlots; ++g) {
                        const float *sp = all_st + (g * kQmax + qq) * kStQ;
                        const int *ip = reinterpret_cast<const int *>(sp);
                        insert(((u64)__float_as_uint(fmaxf(sp[kWBest * kTQ], 0.0f)) << 32) | (unsigned)ip[kWBgrp * kTQ]);   // +inf if nothing seen
                        if (TOP3) {
                            insert(((u64)__float_as_uint(fmaxf(sp[kWSecond * kTQ], 0.0f)) << 32) | (unsigned)ip[kWSgrp * kTQ]);
                            t3 = fminf(t3, fmaxf(sp[kWThird * kTQ], 0.0f));
                        } else {
                            // no runner-up group is tracked: only its value takes part (group field unused)
                            insert(((u64)__float_as_uint(fmaxf(sp[kWSecond * kTQ], 0.0f)) << 32) | 0xffffffffull);
                        }
                    }
                    // a slot that saw no second group reports +inf there: "no runner-up"
                    if ((unsigned)(k2 >> 32) == 0x7f800000u) k2 = kKeyInit;
                }
                mk1[r] = k1; mk2[r] = k2; mt3[r] = t3;
            }
            if (tid == 0) s_namb = 0;
            asm volatile("bar.sync 1, %0;" ::"n"(kEpThreads) : "memory");   // the state may be re-initialised; s_namb is zero
            const float *qc = (sg.dir ? pc2 : pc1) + (size_t)sg.b * sg.nq * 3;
            const float *cc = (sg.dir ? pc1 : pc2) + (size_t)sg.b * sg.nc * 3;
            const size_t qoff = (size_t)sg.b * sg.nq;
            if (!refine) {
                // diagnostic: publish what the filter found (tests measure its error against float64 with this)
                u64 *keys = (sg.dir ? o.key2 : o.key1) + qoff;
                unsigned *secs = (sg.dir ? o.sec2 : o.sec1) + qoff;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    const int qq = slot + 4 * r;
                    const int i = (sg.qb0 + qq) * kTQ + row;
                    if (qq < sg.Q && i < sg.nq) {
                        keys[i] = mk1[r];
                        secs[i] = mk2[r] != kKeyInit ? (unsigned)(mk2[r] >> 32) : 0x7f800000u;
                    }
                }
                if (Cfg::kRawMax > 0) { mbar_wait_spin(bar_rawfull, sn & 1u); mbar_arrive(bar_rawempty); }
                ++sn;
                continue;
            }
            const float *raw = nullptr;
            if (Cfg::kRawMax > 0) {
                mbar_wait_spin(bar_rawfull, sn & 1u);                   // the producers' raw copy of this segment's candidates
                if (stages_raw(sg)) raw = raw_base;
            }
            float *dout = (sg.dir ? o.d2 : o.d1) + qoff;
            int32_t *iout = (sg.dir ? o.i2 : o.i1) + qoff;
            float *zb = sg.dir ? o.z2 : o.z1;
            float *zero = zb ? zb + 3 * qoff : nullptr;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int qq = slot + 4 * r;
                const int i = (sg.qb0 + qq) * kTQ + row;
                if (qq < sg.Q) {
                    if (i < sg.nq && !(RLG_TS_KO & 1)) refine_query<TOP3>(qc, cc, raw, sg.nc, i, mk1[r], mk2[r], mt3[r], dout, iout, zero,
                                                                          qq * kTQ + row, s_dist, s_amb, &s_namb);
                    else s_dist[qq * kTQ + row] = 0.0f;                 // rows past the end of the cloud add nothing
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kEpThreads) : "memory");   // the ambiguous list is complete
            // ---- ambiguous queries: the whole candidate cloud is scanned exactly -- one warp per query from the staged
            // copy, all epilogue threads per query when the cloud is read from global memory
            {
                const int n_amb = (RLG_TS_KO & 2) ? 0 : (int)s_namb;
                if (raw != nullptr) {
                    const int npos = ((sg.nc + kGroup - 1) / kGroup) * kGroup;
                    for (int a = warp; a < n_amb; a += kEpWarps) {
                        const int local = (int)s_amb[a];
                        const int i = (sg.qb0 + (local >> 7)) * kTQ + (local & 127);
                        const float qx = __ldg(qc + 3 * (size_t)i), qy = __ldg(qc + 3 * (size_t)i + 1), qz = __ldg(qc + 3 * (size_t)i + 2);
                        float dist;
                        int bj;
                        scan_query([&](int p, float &x, float &y, float &z) {
                            const float *c = raw + (p >> 5) * kRawStride + 3 * (p & 31);
                            x = c[0]; y = c[1]; z = c[2];
                        }, npos, sg.nc, lane, qx, qy, qz, dist, bj);
                        if (lane == 0) { dout[i] = dist; iout[i] = bj; s_dist[local] = dist; }
                    }
                } else {
                    unsigned *s_red = reinterpret_cast<unsigned *>(smem + Cfg::kOffAmb) + (kQmax * kTQ) / 2;   // behind the list
                    for (int a = 0; a < n_amb; ++a) {
                        const int local = (int)s_amb[a];
                        const int i = (sg.qb0 + (local >> 7)) * kTQ + (local & 127);
                        const float qx = __ldg(qc + 3 * (size_t)i), qy = __ldg(qc + 3 * (size_t)i + 1), qz = __ldg(qc + 3 * (size_t)i + 2);
                        float dist;
                        int bj;
                        scan_query_cta(cc, sg.nc, tid, qx, qy, qz, s_red, dist, bj);
                        if (tid == 0) { dout[i] = dist; iout[i] = bj; s_dist[local] = dist; }
                    }
                }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(kEpThreads) : "memory");   // every distance of the segment is in s_dist
            if (Cfg::kRawMax > 0) mbar_arrive(bar_rawempty);                // done with this segment's raw copy
            // ---- means and loss: warp w sums query block w of the segment; whoever completes the unit reduces its
            // partial sums to the mean; whoever completes the last unit reduces the means to the loss (fixed orders)
            if (o.mean1 != nullptr && warp < sg.Q && !(RLG_TS_KO & 4)) {
                const float *dv = s_dist + warp * kTQ;
                double acc = ((double)dv[lane] + (double)dv[lane + 32]) + ((double)dv[lane + 64] + (double)dv[lane + 96]);
#pragma unroll
                for (int sft = 16; sft > 0; sft >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, sft);
                const int unit = (sg.dir ? B : 0) + sg.b;
                const int nblk = sg.dir ? qb2 : qb1;
                double *parts = o.partial + (size_t)unit * o.pq;
                unsigned last = 0;
                if (lane == 0) {
                    parts[sg.qb0 + warp] = acc;
                    __threadfence();
                    // counters start at 0xffffffff (the workspace's all-ones state): the k-th arrival reads k-2
                    last = atomicAdd(o.unit_cnt + unit, 1u) == (unsigned)(nblk - 2);
                }
                last = __shfl_sync(0xffffffffu, last, 0);
                if (last) {
                    __threadfence();
                    double tot = 0.0;
                    for (int e = lane; e < nblk; e += 32) tot += __ldcg(parts + e);
#pragma unroll
                    for (int sft = 16; sft > 0; sft >>= 1) tot += __shfl_down_sync(0xffffffffu, tot, sft);
                    unsigned last_unit = 0;
                    if (lane == 0) {
                        o.unit_cnt[unit] = 0xffffffffu;
                        (sg.dir ? o.mean2 : o.mean1)[sg.b] = (float)(tot / (double)sg.nq);
                        if (o.loss != nullptr) {
                            __threadfence();
                            last_unit = atomicAdd(o.global_cnt, 1u) == (unsigned)(2 * B - 2);
                        }
                    }
                    last_unit = __shfl_sync(0xffffffffu, last_unit, 0);
                    if (last_unit) {
                        __threadfence();
                        double ls = 0.0;
                        for (int e = lane; e < B; e += 32)
                            ls += (double)o.w1 * (double)__ldcg(o.mean1 + e) + (double)o.w2 * (double)__ldcg(o.mean2 + e);
#pragma unroll
                        for (int sft = 16; sft > 0; sft >>= 1) ls += __shfl_down_sync(0xffffffffu, ls, sft);
                        if (lane == 0) { *o.global_cnt = 0xffffffffu; *o.loss = (float)ls; }
                    }
                }
            }
            ++sn;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

// workspace of the fused forward behind the shared key / second-value arrays: counters (all-ones invariant) + partial sums
static size_t tcsweep_counter_bytes(int B) { return align_up(sizeof(unsigned) * (1 + 2 * (size_t)B), 256); }
static int tcsweep_pq(int N, int M) { return ((N > M ? N : M) + kTQ - 1) / kTQ; }
size_t tcsweep_ws_bytes(int B, int N, int M) {
    return tcsweep_counter_bytes(B) + align_up(sizeof(double) * 2 * (size_t)B * tcsweep_pq(N, M), 256);
}

// The fused forward (one launch).
//   fin_ws    tcsweep_ws_bytes(B,N,M): [global counter | 2B unit counters] (all-ones on entry and on exit) | partial sums
// filter_only: diagnostic -- the filter sweep alone, its per-query results published to w.rowkey/colkey/rowsec/colsec
int launch_tcsweep(const float *pc1, const float *pc2, int B, int N, int M, const FwdWs &w, void *fin_ws,
                   float *d1, float *d2, int32_t *i1, int32_t *i2, float *mean1, float *mean2, float *loss, float w1,
                   float w2, float *zero1, float *zero2, bool filter_only, bool force_top3, int reserve_sms, cudaStream_t st) {
    const int qb1 = (N + kTQ - 1) / kTQ, qb2 = (M + kTQ - 1) / kTQ;
    const long long n_tasks = (long long)B * (qb1 + qb2);
    if (n_tasks > 0x3fffffffLL) return fail(RLG_ERR_TOO_LARGE, "rlg_chamfer_fwd: too many query blocks");
    if ((long long)qb1 * (long long)((M + kTC - 1) / kTC) * B > 0x3fffffffLL || (long long)qb2 * (long long)((N + kTC - 1) / kTC) * B > 0x3fffffffLL)
        return fail(RLG_ERR_TOO_LARGE, "rlg_chamfer_fwd: too many tile visits for 32-bit counters");
    int sms = sm_count();
    if (sms <= 0) return fail((int)cudaErrorNoDevice, "rlg_chamfer_fwd: no CUDA device");
    if (reserve_sms > 0) sms = sms - reserve_sms > 1 ? sms - reserve_sms : 1;    // leave SMs to co-running communication kernels
    // Ambiguous points (runner-up group within the margin) cost a scan of the whole candidate cloud unless the sweep
    // tracks three groups; their share grows with the point density.  Break-even is around 4096 candidates.
    const bool top3 = force_top3 || N > 4096 || M > 4096;
    // per launch (a host-side attribute write, no device work): the setting is per device and this library keeps no
    // per-device state of its own
    {
        cudaError_t e = top3 ? cudaFuncSetAttribute(chamfer_tcsweep_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SweepCfg<true>::kSmem)
                             : cudaFuncSetAttribute(chamfer_tcsweep_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SweepCfg<false>::kSmem);
        if (e != cudaSuccess) { cudaGetLastError(); return fail((int)e, "chamfer_tcsweep_kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); }
    }
    SweepOut o;
    o.d1 = d1; o.d2 = d2; o.i1 = i1; o.i2 = i2; o.z1 = zero1; o.z2 = zero2;
    o.mean1 = mean1; o.mean2 = mean2; o.loss = loss; o.w1 = w1; o.w2 = w2;
    o.global_cnt = (unsigned *)fin_ws;
    o.unit_cnt = o.global_cnt + 1;
    o.partial = (double *)((char *)fin_ws + tcsweep_counter_bytes(B));
    o.pq = tcsweep_pq(N, M);
    o.key1 = w.rowkey; o.key2 = w.colkey; o.sec1 = w.rowsec; o.sec2 = w.colsec;
    const int grid = (int)(n_tasks < sms ? n_tasks : sms);
    // unit-aligned chunks when every (direction, cloud) unit can have a CTA of its own and no chunk needs a second segment
    const int units = 2 * B;
    int split = 0;
    if (units <= grid) {
        const int c_lo = grid / units;
        const int longest = ((qb1 > qb2 ? qb1 : qb2) + c_lo - 1) / c_lo;
        if (longest <= kQmax) split = 1;
    }
    const int refine = filter_only ? 0 : 1;
    cudaError_t le = top3 ? launch_pdl(chamfer_tcsweep_kernel<true>, dim3((unsigned)grid), dim3(kThreads), (size_t)SweepCfg<true>::kSmem, st, pc1, pc2,
                                       B, N, M, (int)n_tasks, qb1, qb2, split, o, refine)
                          : launch_pdl(chamfer_tcsweep_kernel<false>, dim3((unsigned)grid), dim3(kThreads), (size_t)SweepCfg<false>::kSmem, st, pc1, pc2,
                                       B, N, M, (int)n_tasks, qb1, qb2, split, o, refine);
    if (le != cudaSuccess) { cudaGetLastError(); return fail((int)le, "chamfer_tcsweep_kernel: %s", cudaGetErrorString(le)); }
    return check_launch("chamfer_tcsweep_kernel");
}

}  // namespace rlg
