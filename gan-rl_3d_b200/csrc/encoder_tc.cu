// encoder_tc.cu -- PointNet encoder trunk on the 5th-generation tensor cores (tcgen05 + TMEM), bf16 x bf16 -> fp32.
//
// Replaces models/autoencoder.py:65-71 of the reference (transpose, [Conv1d(k=1)+BN+ReLU] x L, max over points)
// in eval mode for the wide configurations, e.g. 3->64->128->1024 (BASELINE config 3).  Per 128-point tile:
//
//   layer 0 (3 -> C1)        CUDA cores, fp32; the result is written as bf16 straight into the K-major,
//                            128-byte-swizzled shared-memory tile the next MMA reads as its A operand.
//   hidden layers            D[128 points x C_out] = A[128 x C_in] . W^T : tcgen05.mma (M=128, N=C_out, K=16 per
//                            instruction), accumulator in TMEM; the epilogue pulls it back with tcgen05.ld, adds
//                            bias, applies ReLU, converts to bf16 and writes the NEXT layer's operand tile.
//   last layer               roles swapped so channels sit on the TMEM lanes:  D[128 channels x 128 points] =
//                            W_blk[128 x C_in] . A^T.  The epilogue is a per-thread running max over the columns,
//                            so the (B, C_last, N) activation never exists anywhere -- not even in shared memory.
//                            Two TMEM accumulators alternate, so block k+1's MMAs run under block k's epilogue.
//
// Weights are packed once (rlg_encoder_pack_bf16) into the exact swizzled shared-memory image and pulled in
// with 1-D TMA bulk copies (cp.async.bulk -> UBLKCP) onto an mbarrier; they then stay RESIDENT in shared memory:
// a CTA owns one slice of the last layer's channels (C_last / split) and loops over (cloud, point-chunk) tasks,
// so weight traffic is one load per CTA per launch.  bias + ReLU of the last layer commute with the max and are
// applied once per task; partial maxima of the point chunks of a cloud are merged with an integer atomicMax
// (post-ReLU values are >= 0, so their bit patterns order like ints).
#include "common.cuh"
#include "tcgen05.cuh"
#include <cuda_bf16.h>
#include <type_traits>

namespace rlg {

static constexpr int kTcMaxLayers = 8;
static constexpr int kTcThreads = 128;
static constexpr int kTileP = 128;          // points per tile (UMMA M for hidden layers, N for the last)
static constexpr int kMaxBlk = 8;           // last-layer 128-channel blocks per CTA
static constexpr int kTmemCols = 512;
static constexpr uint32_t kColHidden = 0, kColLast0 = 128, kColLast1 = 256;

struct TcPlan {
    int L;                              // layers including layer 0
    int c[kTcMaxLayers + 1];            // widths, c[0] = 3
    const float *w0, *b0;               // layer 0, fp32 (c[1] x 3), (c[1])
    const float *bias[kTcMaxLayers];    // fp32 biases of layers >= 1
    const unsigned char *img[kTcMaxLayers];   // packed bf16 images of layers >= 1 (global)
    uint32_t smem_w[kTcMaxLayers];      // shared-memory byte offsets (1024-B aligned)
    uint32_t smem_act[2], smem_x[2], smem_bias[kTcMaxLayers], smem_w0, smem_bar, smem_total;
    int split, nblk;                    // last layer: channel groups over CTAs, 128-channel blocks per CTA
    int n_pchunks, tiles_per_chunk;     // point chunks per cloud, 128-point tiles per chunk
};

// D[M x N] (+)= A[M x K] . B[N x K]^T over K, A/B K-major SW128 tiles at a_smem/b_smem with a_rows/b_rows rows
__device__ __forceinline__ void issue_gemm(uint32_t d_tmem, uint32_t a_smem, int a_rows, uint32_t b_smem, int b_rows,
                                           int M, int N, int K) {
    const uint32_t idesc = umma_idesc(M, N);
    uint32_t acc = 0;
    for (int kb = 0; kb < K / 64; ++kb) {
        const uint64_t ad = umma_desc(a_smem + kb * a_rows * 128);
        const uint64_t bd = umma_desc(b_smem + kb * b_rows * 128);
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {                  // 16 bf16 = 32 bytes per MMA along K
            tc_mma_bf16(d_tmem, ad + (uint64_t)(k4 * 2), bd + (uint64_t)(k4 * 2), idesc, acc);
            acc = 1;
        }
    }
}

// Same GEMM as issue_gemm, written so that every operand is warp-uniform (descriptors advance by constant
// increments): called by ALL lanes of the MMA warp, the elected lane issues.  Keeping the descriptor arithmetic on
// the uniform datapath avoids a register->uniform-register move per tcgen05.mma operand.
__device__ __forceinline__ void issue_gemm_uniform(bool leader, uint32_t d_tmem, uint32_t a_smem, int a_rows,
                                                   uint32_t b_smem, int b_rows, int M, int N, int K) {
    const uint32_t idesc = umma_idesc(M, N);
    uint64_t ad = umma_desc(a_smem), bd = umma_desc(b_smem);
    const uint64_t a_inc = (uint64_t)((uint32_t)a_rows * 128u >> 4), b_inc = (uint64_t)((uint32_t)b_rows * 128u >> 4);
    uint32_t acc = 0;
    for (int kb = 0; kb < K / 64; ++kb) {
        if (leader) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {              // 16 bf16 = 32 bytes per MMA along K
                tc_mma_bf16(d_tmem, ad + (uint64_t)(k4 * 2), bd + (uint64_t)(k4 * 2), idesc, acc);
                acc = 1;
            }
        }
        acc = 1;
        ad += a_inc;
        bd += b_inc;
    }
}

// ---- warp-specialised pipeline ------------------------------------------------------------------------------
//   warps 0-7   FE   two threads per point of the 128-point tile (each half of the channels): layer 0 on the CUDA
//                    cores, then the bias+ReLU+bf16 epilogue of every hidden layer (TMEM -> registers -> swizzled
//                    operand tile of the next GEMM)
//   warps 8,17  MMA  two issuers, one elected thread each.  A tcgen05.mma blocks its issuing thread until the tensor
//                    pipe takes it (measured with tools/umma_bench.cu: the queue is about one instruction deep, so every
//                    cycle the issuer spends on barrier waits, fences and commits is a tensor-pipe bubble).  Two issuers
//                    alternate last-layer blocks -- each owns one TMEM accumulator -- so one's bookkeeping runs under
//                    the other's MMAs.  Warp 8 also owns the TMEM allocation, the weight TMA loads and the hidden GEMMs.
//   warps 9-16  EP   last-layer epilogue: tcgen05.ld of a 128-channel x 128-point accumulator (two warps per TMEM lane
//                    quarter, 64 columns each), running max per channel
// Hand-offs are mbarriers; the MMA thread interleaves the hidden GEMM of tile n+1 between the last-layer blocks of
// tile n, so the front end of the next tile runs under the tensor-core time of the current one.
//   TMEM columns: [0,128) the hidden accumulator H, [128,512) three last-layer accumulators (a ring shared by the two
//   issuers: block kb goes to accumulator kb % 3 and is issued by issuer kb % 2).
static constexpr int kFeThreads = 256, kEpThreads = 256;
static constexpr int kMmaWarp = kFeThreads / 32;                        // issuer A (+ TMEM owner, weight loads, hidden GEMMs)
static constexpr int kMmaWarpB = (kFeThreads + 32 + kEpThreads) / 32;   // issuer B: the last warp
static constexpr int kTcThreads2 = kFeThreads + 32 + kEpThreads + 32;
static constexpr int kAccBufs = 3;
static constexpr uint32_t kColH = 0, kColAcc = 128;      // H at columns [0,128), three accumulators at 128 / 256 / 384

// two fp32 -> packed bf16x2 with ReLU folded into the conversion (round-to-nearest-even, negative -> +0): the same
// bits as fmaxf(.,0) followed by the conversion, in one instruction per pair
__device__ __forceinline__ uint32_t relu_pack_bf16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__global__ void __launch_bounds__(kTcThreads2, 1) encoder_tc_kernel(const float *__restrict__ x, int B, int N, TcPlan p,
                                                                   float *__restrict__ pooled) {
    extern __shared__ unsigned char smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // SWIZZLE_128B operand tiles need a 1024-byte aligned base: realign inside the 1 KB of slack the launch adds
    const uint32_t pad = (1024u - (smem_u32(smem_raw) & 1023u)) & 1023u;
    unsigned char *smem = smem_raw + pad;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bars = sbase + p.smem_bar;
    // every pair below is indexed by tile parity (n & 1) resp. accumulator buffer
    const uint32_t bar_w = bars, bar_fe = bars + 8, bar_h = bars + 24;
    const uint32_t bar_xfull = bars + 40, bar_xempty = bars + 56;
    const uint32_t bar_accfull = bars + 72, bar_accempty = bars + 96;           // [3] each
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + p.smem_bar + 120);
    const int L = p.L, c_last = p.c[L], k_last = p.c[L - 1];
    const int g = blockIdx.x % p.split, cta_in_group = blockIdx.x / p.split, ctas_per_group = gridDim.x / p.split;
    const int blk0 = g * p.nblk;
    const int nblk = min(p.nblk, c_last / 128 - blk0);
    const int n_tasks = (nblk > 0) ? B * p.n_pchunks : 0;
    const int n_hidden = L - 2;                                                 // tensor-core layers before the last
    // With <= 2 hidden layers the front end computes layer 0 of tile n+1 BEFORE the epilogue of tile n (the layer-0 tile
    // is free once the first hidden GEMM of tile n is done), so issuer A can start the next tile's hidden GEMM the
    // moment this tile's epilogue has drained H.  Deeper chains would overwrite a live operand tile: plain order.
    const bool early = n_hidden >= 1 && n_hidden <= 2;

    if (tid == 0) {
        mbar_init(bar_w, 1);
        for (int k = 0; k < 2; ++k) {
            mbar_init(bar_fe + 8 * k, kFeThreads); mbar_init(bar_h + 8 * k, 1);
            mbar_init(bar_xfull + 8 * k, kFeThreads); mbar_init(bar_xempty + 8 * k, 2);   // both issuers commit
        }
        for (int k = 0; k < kAccBufs; ++k) { mbar_init(bar_accfull + 8 * k, 1); mbar_init(bar_accempty + 8 * k, kEpThreads); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // tiles this CTA will see (same count in every role)
    int T = 0;
    for (int task = cta_in_group; task < n_tasks; task += ctas_per_group) {
        const int pc = task % p.n_pchunks;
        for (int t = 0; t < p.tiles_per_chunk; ++t)
            if ((pc * p.tiles_per_chunk + t) * kTileP < N) ++T;
    }

    if (warp == kMmaWarp || warp == kMmaWarpB) {
        // =========================== MMA issuers ===========================
        // All 32 lanes run this (warp-uniform) code; one elected lane issues the tcgen05 instructions.
        const int me = (warp == kMmaWarp) ? 0 : 1;
        if (T > 0) {
            const bool leader = elect_one();
            if (me == 0 && leader) {
                // resident weights: one TMA bulk load per layer (the last layer: this CTA's channel slice)
                uint32_t total = 0;
                for (int l = 1; l < L - 1; ++l) total += (uint32_t)p.c[l + 1] * p.c[l] * 2;
                total += (uint32_t)nblk * 128 * k_last * 2;
                mbar_arrive_expect_tx(bar_w, total);
                for (int l = 1; l < L; ++l) {
                    const bool last = (l == L - 1);
                    uint32_t bytes = last ? (uint32_t)nblk * 128 * k_last * 2 : (uint32_t)p.c[l + 1] * p.c[l] * 2;
                    const unsigned char *src = p.img[l] + (last ? (size_t)blk0 * 128 * k_last * 2 : 0);
                    uint32_t dst = sbase + p.smem_w[l];
                    while (bytes) {                                  // <= 32 KB per bulk copy
                        const uint32_t n = bytes < 32768u ? bytes : 32768u;
                        bulk_g2s(dst, src, n, bar_w);
                        dst += n; src += n; bytes -= n;
                    }
                }
            }
            __syncwarp();
            mbar_wait_wd(bar_w, 0);
            uint32_t fe_ph[2] = {0, 0};
            int kb = 0;                                           // running last-layer block counter (both issuers count)
            const uint32_t wl = sbase + p.smem_w[L - 1];
            auto hidden_step = [&](int n, int l) {                // hidden GEMM l of tile n into H[n & 1]  (issuer A)
                const int par = n & 1;
                mbar_wait_wd(bar_fe + 8 * par, par ? fe_ph[1] : fe_ph[0]);
                if (par) fe_ph[1] ^= 1; else fe_ph[0] ^= 1;
                tc_fence_after();
                const uint32_t in = sbase + (((l - 1) & 1) ? p.smem_act[1] : p.smem_act[0]);
                issue_gemm_uniform(leader, tmem + kColH, in, kTileP, sbase + p.smem_w[l], p.c[l + 1], kTileP, p.c[l + 1], p.c[l]);
                if (leader) tc_commit(bar_h + 8 * par);
                __syncwarp();
            };
            // block kb goes to accumulator kb % 3 and is issued by issuer kb & 1
            auto last_block = [&](int blk, uint32_t act) {
                const int a = kb % kAccBufs, use = kb / kAccBufs;
                if ((kb & 1) == me) {
                    mbar_wait_wd(bar_accempty + 8 * a, (uint32_t)((use & 1) ^ 1));
                    tc_fence_after();
                    issue_gemm_uniform(leader, tmem + kColAcc + (uint32_t)a * 128u, wl + (uint32_t)blk * 128 * k_last * 2, 128,
                                       act, kTileP, 128, kTileP, k_last);
                    if (leader) tc_commit(bar_accfull + 8 * a);
                    __syncwarp();
                }
                ++kb;
            };
            if (me == 0 && n_hidden >= 1) hidden_step(0, 1);      // prologue: first hidden GEMM of the first tile
            for (int n = 0; n < T; ++n) {
                const int s = n & 1;
                if (me == 0)
                    for (int l = 2; l <= n_hidden; ++l) hidden_step(n, l);
                mbar_wait_wd(bar_xfull + 8 * s, (uint32_t)((n >> 1) & 1));    // X[s] is ready, and H has been drained
                tc_fence_after();
                // first hidden GEMM of the next tile ahead of this tile's blocks: the front end gets H back early
                if (me == 0 && n_hidden >= 1 && n + 1 < T) hidden_step(n + 1, 1);
                for (int blk = 0; blk < nblk; ++blk) last_block(blk, sbase + p.smem_x[s]);
                if (leader) tc_commit(bar_xempty + 8 * s);        // X[s] may be overwritten once BOTH issuers' MMAs are done
                __syncwarp();
            }
        }
    } else if (warp < kMmaWarp) {
        // =========================== front end ===========================
        float *w0s = reinterpret_cast<float *>(smem + p.smem_w0);
        const int c1 = p.c[1];
        // layer-0 weights per channel PAIR (2k, 2k+1): (wx,wx', wy,wy' | wz,wz', b,b') -- FFMA2 operands
        for (int e = tid; e < c1 * 4; e += kFeThreads) {
            const int pr = e >> 3, j = e & 7, o = 2 * pr + (j & 1), k = j >> 1;
            w0s[e] = (k < 3) ? __ldg(p.w0 + o * 3 + k) : __ldg(p.b0 + o);
        }
        for (int l = 1; l < L - 1; ++l) {
            float *bs = reinterpret_cast<float *>(smem + p.smem_bias[l]);
            for (int e = tid; e < p.c[l + 1]; e += kFeThreads) bs[e] = __ldg(p.bias[l] + e);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kFeThreads) : "memory");
        const int row = (warp & 3) * 32 + lane;                        // point of the tile == TMEM lane
        const uint32_t row_off = (uint32_t)row * 128u, row_x = (uint32_t)(row & 7);   // swizzled store address parts
        const int half = warp >> 2;                                    // which half of the channels this thread does
        const uint32_t lane_base = ((uint32_t)(warp & 3) * 32u) << 16; // this warp's TMEM lane quarter
        uint32_t h_ph[2] = {0, 0};

        // tile cursor with one tile of look-ahead
        struct Cursor { int task, t, b, n0; bool valid; };
        auto first = [&]() {
            Cursor c{cta_in_group, 0, 0, 0, cta_in_group < n_tasks};
            if (c.valid) { c.b = c.task / p.n_pchunks; c.n0 = (c.task - c.b * p.n_pchunks) * p.tiles_per_chunk * kTileP; }
            return c;
        };
        auto advance = [&](Cursor c) {
            ++c.t;
            c.n0 += kTileP;
            if (c.t >= p.tiles_per_chunk || c.n0 >= N) {
                c.task += ctas_per_group;
                c.t = 0;
                c.valid = c.task < n_tasks;
                if (c.valid) { c.b = c.task / p.n_pchunks; c.n0 = (c.task - c.b * p.n_pchunks) * p.tiles_per_chunk * kTileP; }
            }
            return c;
        };
        auto load_point = [&](const Cursor &c, float &ox, float &oy, float &oz) {
            const float *src = x + ((size_t)c.b * N + min(c.n0 + row, N - 1)) * 3;    // rows past the end repeat the last
            // volatile: keeps the loads where they are written (one tile ahead of their use)
            asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(ox) : "l"(src));
            asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(oy) : "l"(src + 1));
            asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(oz) : "l"(src + 2));
        };
        // 16-byte chunk k8 (8 channels) of this thread's row in a [128 x K] K-major SW128 tile
        auto chunk_ptr = [&](unsigned char *tile, int k8) {
            return reinterpret_cast<uint4 *>(tile + (uint32_t)(k8 >> 3) * (kTileP * 128u) + row_off +
                                             ((((uint32_t)k8 & 7u) ^ row_x) << 4));
        };
        // layer 0 on the CUDA cores for this thread's point and half of the channels -> bf16 operand tile.
        // Packed FFMA2: two channels per instruction (same IEEE fma per lane as the scalar form).
        auto layer0 = [&](float px, float py, float pz, unsigned char *dst) {
            const u64 px2 = pack2(px, px), py2 = pack2(py, py), pz2 = pack2(pz, pz);
            const int k8n = c1 / 16;                                   // 16-byte chunks (8 channels) per thread
#pragma unroll 2
            for (int kk = 0; kk < k8n; ++kk) {
                const int k8 = half * k8n + kk;
                const ulonglong2 *wp = reinterpret_cast<const ulonglong2 *>(w0s) + k8 * 8;   // 4 pairs x 2 vectors
                uint32_t pk[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const ulonglong2 wa = wp[2 * q], wb = wp[2 * q + 1];
                    const u64 r = fma2(wa.x, px2, fma2(wa.y, py2, fma2(wb.x, pz2, wb.y)));
                    float lo, hi;
                    unpack2(r, lo, hi);
                    pk[q] = relu_pack_bf16(lo, hi);
                }
                *chunk_ptr(dst, k8) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
            fence_async_proxy();
        };
        // epilogue of hidden layer l of tile n: H[n & 1] -> bias + ReLU -> bf16 -> dst (NCOL = 32 or 64 channels)
        auto epilogue_cols = [&](auto ncol_tag, int n, int l, unsigned char *dst) {
            constexpr int NCOL = decltype(ncol_tag)::value;
            const int col0 = half * NCOL;
            const ulonglong2 *bs = reinterpret_cast<const ulonglong2 *>(reinterpret_cast<const float *>(smem + p.smem_bias[l]) + col0);
            const uint32_t taddr = tmem + lane_base + kColH + (uint32_t)col0;
            float v[NCOL];
            if constexpr (NCOL == 64) tc_ld64(taddr, v); else tc_ld32(taddr, v);
#pragma unroll
            for (int q8 = 0; q8 < NCOL / 8; ++q8) {
                const ulonglong2 b0 = bs[2 * q8], b1 = bs[2 * q8 + 1];        // 8 biases as 4 pairs
                const u64 s0 = add2(pack2(v[8 * q8], v[8 * q8 + 1]), b0.x), s1 = add2(pack2(v[8 * q8 + 2], v[8 * q8 + 3]), b0.y);
                const u64 s2 = add2(pack2(v[8 * q8 + 4], v[8 * q8 + 5]), b1.x), s3 = add2(pack2(v[8 * q8 + 6], v[8 * q8 + 7]), b1.y);
                float a0, a1, a2, a3, a4, a5, a6, a7;
                unpack2(s0, a0, a1); unpack2(s1, a2, a3); unpack2(s2, a4, a5); unpack2(s3, a6, a7);
                *chunk_ptr(dst, (col0 >> 3) + q8) = make_uint4(relu_pack_bf16(a0, a1), relu_pack_bf16(a2, a3),
                                                               relu_pack_bf16(a4, a5), relu_pack_bf16(a6, a7));
            }
            fence_async_proxy();
            tc_fence_before();
        };
        auto hidden_epilogue = [&](int n, int l, unsigned char *dst) {
            if (p.c[l + 1] == 128) epilogue_cols(std::integral_constant<int, 64>{}, n, l, dst);
            else epilogue_cols(std::integral_constant<int, 32>{}, n, l, dst);
        };

        Cursor cur = first();
        float px = 0.f, py = 0.f, pz = 0.f;
        if (cur.valid) load_point(cur, px, py, pz);
        if (n_hidden >= 1 && cur.valid) {                              // prologue: layer 0 of the first tile
            layer0(px, py, pz, smem + p.smem_act[0]);
            mbar_arrive(bar_fe);
        }
        for (int n = 0; cur.valid; ++n) {
            const int s = n & 1;
            const uint32_t xempty_par = (uint32_t)(((n >> 1) & 1) ^ 1);
            const Cursor nxt = advance(cur);
            float qx = 0.f, qy = 0.f, qz = 0.f;
            if (nxt.valid) load_point(nxt, qx, qy, qz);                // next tile's point, one tile ahead
            if (n_hidden == 0) {
                mbar_wait_wd(bar_xempty + 8 * s, xempty_par);
                layer0(px, py, pz, smem + p.smem_x[s]);
                mbar_arrive(bar_xfull + 8 * s);
            } else {
                // hidden GEMM 1 of this tile is done: H[s] holds it and the layer-0 tile A0 is free again
                mbar_wait_wd(bar_h + 8 * s, s ? h_ph[1] : h_ph[0]);
                if (s) h_ph[1] ^= 1; else h_ph[0] ^= 1;
                tc_fence_after();
                if (early && nxt.valid) {
                    layer0(qx, qy, qz, smem + p.smem_act[0]);
                    mbar_arrive(bar_fe + 8 * (s ^ 1));
                }
                for (int l = 1; l <= n_hidden; ++l) {
                    const bool last_hidden = (l == n_hidden);
                    if (l > 1) {
                        mbar_wait_wd(bar_h + 8 * s, s ? h_ph[1] : h_ph[0]);
                        if (s) h_ph[1] ^= 1; else h_ph[0] ^= 1;
                        tc_fence_after();
                    }
                    if (last_hidden) mbar_wait_wd(bar_xempty + 8 * s, xempty_par);
                    hidden_epilogue(n, l, smem + (last_hidden ? p.smem_x[s] : ((l & 1) ? p.smem_act[1] : p.smem_act[0])));
                    mbar_arrive(last_hidden ? bar_xfull + 8 * s : bar_fe + 8 * s);
                }
                if (!early && nxt.valid) {
                    layer0(qx, qy, qz, smem + p.smem_act[0]);
                    mbar_arrive(bar_fe + 8 * (s ^ 1));
                }
            }
            px = qx; py = qy; pz = qz;
            cur = nxt;
        }
    } else {
        // =========================== last-layer epilogue ===========================
        const int q = warp & 3;                                        // TMEM lane quarter this warp may read
        const int half = (warp - kMmaWarp - 1) >> 2;                   // which 64 of the 128 points (columns)
        const uint32_t lane_base = ((uint32_t)q * 32u) << 16;
        const int ch_in_blk = q * 32 + lane;
        const float *bl = p.bias[L - 1];
        int kb = 0;
        for (int task = cta_in_group; task < n_tasks; task += ctas_per_group) {
            const int b = task / p.n_pchunks, pc = task - b * p.n_pchunks;
            float runmax[kMaxBlk];
#pragma unroll
            for (int k = 0; k < kMaxBlk; ++k) runmax[k] = -INFINITY;
            for (int t = 0; t < p.tiles_per_chunk; ++t) {
                if ((pc * p.tiles_per_chunk + t) * kTileP >= N) break;
#pragma unroll
                for (int blk = 0; blk < kMaxBlk; ++blk) {
                    if (blk < nblk) {
                        const int a = kb % kAccBufs, use = kb / kAccBufs;
                        mbar_wait_wd(bar_accfull + 8 * a, (uint32_t)(use & 1));
                        tc_fence_after();
                        float v[64];
                        tc_ld64(tmem + lane_base + kColAcc + (uint32_t)a * 128u + (uint32_t)half * 64u, v);
                        tc_fence_before();
                        mbar_arrive(bar_accempty + 8 * a);             // the accumulator is free as soon as it is in registers
                        float m0 = runmax[blk], m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
                        for (int i = 0; i < 64; i += 8) {
                            m0 = max3(m0, v[i], v[i + 1]);
                            m1 = max3(m1, v[i + 2], v[i + 3]);
                            m2 = max3(m2, v[i + 4], v[i + 5]);
                            m3 = max3(m3, v[i + 6], v[i + 7]);
                        }
                        runmax[blk] = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                        ++kb;
                    }
                }
            }
            // task result: bias + ReLU commute with the max; merge the point chunks of a cloud with atomicMax
#pragma unroll
            for (int blk = 0; blk < kMaxBlk; ++blk) {
                if (blk < nblk) {
                    const int ch = (blk0 + blk) * 128 + ch_in_blk;
                    const float v = fmaxf(runmax[blk] + __ldg(bl + ch), 0.0f);
                    atomicMax(reinterpret_cast<int *>(pooled) + (size_t)b * c_last + ch, __float_as_int(v));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
    }
}

// fp32 (c_out, c_in) row-major -> bf16 K-major SW128 image; last layer: consecutive 128-row blocks
__global__ void __launch_bounds__(256) encoder_pack_kernel(const float *__restrict__ w, int cout, int cin, int rows_per_tile,
                                                          unsigned char *__restrict__ img) {
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= cout * cin) return;
    const int o = e / cin, k = e - o * cin;
    const int tile = o / rows_per_tile, r = o - tile * rows_per_tile;
    const size_t off = (size_t)tile * rows_per_tile * cin * 2 + sw128_chunk_off(rows_per_tile, r, k >> 3) + (size_t)(k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16 *>(img + off) = __float2bfloat16_rn(w[e]);
}

static int tc_plan(const rlg_layer *layers, int L, int N, TcPlan &p, size_t img_off[kTcMaxLayers], size_t &img_total) {
    if (!layers || L < 2 || L > kTcMaxLayers)
        return fail(RLG_ERR_UNSUPPORTED, "rlg_encoder bf16: need 2..%d layers, got %d", kTcMaxLayers, L);
    if (layers[0].c_in != 3) return fail(RLG_ERR_UNSUPPORTED, "rlg_encoder bf16: layer 0 must have c_in == 3");
    p.L = L;
    p.c[0] = 3;
    for (int l = 0; l < L; ++l) {
        if (!layers[l].w || !layers[l].b) return fail(RLG_ERR_NULL_POINTER, "rlg_encoder bf16: layer %d has null weights", l);
        if (l > 0 && layers[l].c_in != layers[l - 1].c_out)
            return fail(RLG_ERR_BAD_SHAPE, "rlg_encoder bf16: layer %d c_in %d != previous c_out %d", l, layers[l].c_in,
                        layers[l - 1].c_out);
        p.c[l + 1] = layers[l].c_out;
        if (l >= 1) p.bias[l] = layers[l].b;
    }
    for (int l = 1; l < L; ++l)
        if (p.c[l] % 64 != 0 || p.c[l] > 128)
            return fail(RLG_ERR_UNSUPPORTED, "rlg_encoder bf16: hidden width %d must be 64 or 128 (use the fp32 path)", p.c[l]);
    if (p.c[L] % 128 != 0) return fail(RLG_ERR_UNSUPPORTED, "rlg_encoder bf16: last width %d must be a multiple of 128", p.c[L]);
    p.w0 = layers[0].w;
    p.b0 = layers[0].b;
    // packed image offsets (global)
    size_t off = 0;
    for (int l = 1; l < L; ++l) { img_off[l] = off; off += align_up((size_t)p.c[l + 1] * p.c[l] * 2, 256); }
    img_total = off;
    // shared-memory plan
    const int k_last = p.c[L - 1], nblk_total = p.c[L] / 128;
    uint32_t fixed = 0;
    for (int l = 1; l < L - 1; ++l) { p.smem_w[l] = fixed; fixed += (uint32_t)align_up((size_t)p.c[l + 1] * p.c[l] * 2, 1024); }
    // operand tiles: A0/A1 alternate along the hidden chain (only as wide as their users), X[2] feeds the last layer
    uint32_t wa0 = 0, wa1 = 0;
    for (int l = 1; l <= L - 2; ++l) { uint32_t &wdt = ((l - 1) & 1) ? wa1 : wa0; if ((uint32_t)p.c[l] > wdt) wdt = (uint32_t)p.c[l]; }
    p.smem_act[0] = fixed; fixed += kTileP * wa0 * 2;
    p.smem_act[1] = fixed; fixed += kTileP * wa1 * 2;
    p.smem_x[0] = fixed; fixed += kTileP * (uint32_t)p.c[L - 1] * 2;
    p.smem_x[1] = fixed; fixed += kTileP * (uint32_t)p.c[L - 1] * 2;
    p.smem_w[L - 1] = fixed;
    const uint32_t blk_bytes = 128u * k_last * 2u;
    uint32_t tail = 0;
    for (int l = 1; l < L - 1; ++l) { p.smem_bias[l] = tail; tail += (uint32_t)align_up((size_t)p.c[l + 1] * 4, 16); }
    const uint32_t w0_bytes = (uint32_t)p.c[1] * 16, bar_bytes = 128;
    const uint32_t budget = 226u * 1024u;          // 227 KB per CTA minus the 1 KB alignment slack
    if (fixed + tail + w0_bytes + bar_bytes + blk_bytes > budget)
        return fail(RLG_ERR_UNSUPPORTED, "rlg_encoder bf16: resident weights do not fit in shared memory");
    int nblk = (int)((budget - fixed - tail - w0_bytes - bar_bytes) / blk_bytes);
    if (nblk > kMaxBlk) nblk = kMaxBlk;
    if (nblk > nblk_total) nblk = nblk_total;
    p.split = (nblk_total + nblk - 1) / nblk;
    p.nblk = (nblk_total + p.split - 1) / p.split;
    uint32_t o2 = fixed + (uint32_t)p.nblk * blk_bytes;
    for (int l = 1; l < L - 1; ++l) p.smem_bias[l] += o2;
    o2 += tail;
    p.smem_w0 = o2; o2 += w0_bytes;
    p.smem_bar = (uint32_t)align_up(o2, 16); o2 = p.smem_bar + bar_bytes;
    p.smem_total = o2;
    // point chunks: ~4 per SM-worth of tasks keeps the persistent CTAs balanced; a chunk is >= 1 tile
    const int tiles = (N + kTileP - 1) / kTileP;
    p.tiles_per_chunk = tiles >= 8 ? 4 : tiles;
    p.n_pchunks = (tiles + p.tiles_per_chunk - 1) / p.tiles_per_chunk;
    return 0;
}

}  // namespace rlg

using namespace rlg;

extern "C" {

size_t rlg_encoder_pack_bytes(const rlg_layer *layers, int L) {
    TcPlan p;
    size_t off[kTcMaxLayers], total = 0;
    if (tc_plan(layers, L, 128, p, off, total) != 0) return 0;
    return total;
}

int rlg_encoder_pack_bf16(const rlg_layer *layers, int L, void *packed, size_t packed_bytes, void *stream) {
    TcPlan p;
    size_t off[kTcMaxLayers], total = 0;
    int rc = tc_plan(layers, L, 128, p, off, total);
    if (rc) return rc;
    if (!packed || packed_bytes < total || ((uintptr_t)packed & 255u))
        return fail(RLG_ERR_WORKSPACE, "rlg_encoder_pack_bf16: buffer %p/%zu bytes, need %zu bytes 256-B aligned", packed,
                    packed_bytes, total);
    cudaStream_t st = (cudaStream_t)stream;
    for (int l = 1; l < L; ++l) {
        const int cout = p.c[l + 1], cin = p.c[l];
        const int rows = (l == L - 1) ? 128 : cout;
        encoder_pack_kernel<<<(cout * cin + 255) / 256, 256, 0, st>>>(layers[l].w, cout, cin, rows,
                                                                     (unsigned char *)packed + off[l]);
    }
    return check_launch("encoder_pack_kernel");
}

int rlg_encoder_fwd_bf16(const float *x, int B, int N, const rlg_layer *layers, int L, const void *packed,
                         size_t packed_bytes, float *pooled, void *stream) {
    if (B < 0 || N < 1) return fail(RLG_ERR_BAD_SHAPE, "rlg_encoder_fwd_bf16: bad shape B=%d N=%d", B, N);
    if (B == 0) return 0;
    if (!x || !pooled || !packed) return fail(RLG_ERR_NULL_POINTER, "rlg_encoder_fwd_bf16: null pointer");
    TcPlan p;
    size_t off[kTcMaxLayers], total = 0;
    int rc = tc_plan(layers, L, N, p, off, total);
    if (rc) return rc;
    if (packed_bytes < total || ((uintptr_t)packed & 255u))
        return fail(RLG_ERR_WORKSPACE, "rlg_encoder_fwd_bf16: packed weights %zu bytes, need %zu (256-B aligned)", packed_bytes, total);
    for (int l = 1; l < L; ++l) p.img[l] = (const unsigned char *)packed + off[l];
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(pooled, 0, sizeof(float) * (size_t)B * p.c[L], st);
    if (e != cudaSuccess) { cudaGetLastError(); return fail((int)e, "rlg_encoder_fwd_bf16: cudaMemsetAsync: %s", cudaGetErrorString(e)); }
    const size_t smem_bytes = p.smem_total + 1024;       // slack for the 1024-B alignment of the dynamic base
    e = cudaFuncSetAttribute(encoder_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) { cudaGetLastError(); return fail((int)e, "rlg_encoder_fwd_bf16: cudaFuncSetAttribute(%zu): %s", smem_bytes, cudaGetErrorString(e)); }
    const int sms = sm_count();
    if (sms <= 0) return fail((int)cudaErrorNoDevice, "rlg_encoder_fwd_bf16: no CUDA device");
    long long tasks = (long long)B * p.n_pchunks;
    long long per_group = sms / p.split;
    if (per_group < 1) per_group = 1;
    if (per_group > tasks) per_group = tasks;
    const unsigned grid = (unsigned)(per_group * p.split);
    encoder_tc_kernel<<<grid, kTcThreads2, smem_bytes, st>>>(x, B, N, p, pooled);
    return check_launch("encoder_tc_kernel");
}

}  // extern "C"
