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
#include <cuda_bf16.h>

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

// ---- PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of warp w gets lane 32*(w%4)+t
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t"
                 "tcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 lanes x 64 consecutive fp32 columns in one instruction (the wait is part of the same statement, so no use of
// the registers can be scheduled ahead of it)
__device__ __forceinline__ void tc_ld64(uint32_t taddr, float (&v)[64]) {
    uint32_t r[64];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
                 "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,"
                 "%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];\n\t"
                 "tcgen05.wait::ld.sync.aligned;"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
                   "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
                   "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
                   "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
                   "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address and offsets
// in 16-byte units, LBO = 1 (unused for swizzled K-major), SBO = 1024 B between 8-row groups, version 1.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) |
           (2ull << 61);
}
// cute::UMMA::InstrDescriptor: D fp32, A/B bf16, both K-major
__device__ __forceinline__ uint32_t umma_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// byte offset of 8 consecutive K elements (one 16-byte chunk) of row r in a [rows x K] K-major SW128 tile
__host__ __device__ __forceinline__ uint32_t sw128_chunk_off(int rows, int r, int k8) {
    return (uint32_t)((k8 >> 3) * rows * 128 + r * 128 + (((k8 & 7) ^ (r & 7)) << 4));
}

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

// one lane of a converged warp (the same one every time)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
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
//   warp  8     MMA  one elected thread issues every tcgen05.mma; owns the TMEM allocation and the weight TMA loads
//   warps 9-16  EP   last-layer epilogue: tcgen05.ld of a 128-channel x 128-point accumulator (two warps per TMEM lane
//                    quarter, 64 columns each), running max per channel
// Hand-offs are mbarriers; the MMA thread interleaves the hidden GEMM of tile n+1 between the last-layer blocks of
// tile n, so the front end of the next tile runs under the tensor-core time of the current one.
//   TMEM columns: [0,128) hidden accumulator H, [128,512) three last-layer accumulators.
static constexpr int kFeThreads = 256, kEpThreads = 256;
static constexpr int kMmaWarp = kFeThreads / 32;
static constexpr int kTcThreads2 = kFeThreads + 32 + kEpThreads;
static constexpr int kAccBufs = 3;
static constexpr uint32_t kColH = 0, kColAcc = 128;

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a protocol bug must trap, not hang the GPU
__device__ __forceinline__ void mbar_wait_wd(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    for (uint32_t spin = 0;; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (spin > (1u << 26)) __trap();
    }
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
    const uint32_t bar_w = bars, bar_fe = bars + 8, bar_h = bars + 16;
    const uint32_t bar_xfull = bars + 24, bar_xempty = bars + 40;               // [2] each
    const uint32_t bar_accfull = bars + 56, bar_accempty = bars + 80;           // [3] each
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + p.smem_bar + 104);
    const int L = p.L, c_last = p.c[L], k_last = p.c[L - 1];
    const int g = blockIdx.x % p.split, cta_in_group = blockIdx.x / p.split, ctas_per_group = gridDim.x / p.split;
    const int blk0 = g * p.nblk;
    const int nblk = min(p.nblk, c_last / 128 - blk0);
    const int n_tasks = (nblk > 0) ? B * p.n_pchunks : 0;
    const int n_hidden = L - 2;                                                 // tensor-core layers before the last

    if (tid == 0) {
        mbar_init(bar_w, 1); mbar_init(bar_fe, kFeThreads); mbar_init(bar_h, 1);
        for (int k = 0; k < 2; ++k) { mbar_init(bar_xfull + 8 * k, kFeThreads); mbar_init(bar_xempty + 8 * k, 1); }
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

    if (warp == kMmaWarp) {
        // =========================== MMA issuer ===========================
        // All 32 lanes run this (warp-uniform) code; one elected lane issues the tcgen05 instructions.
        if (n_tasks > 0) {
            const bool leader = elect_one();
            if (leader) {
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
            // how many tiles this CTA will see
            int T = 0;
            for (int task = cta_in_group; task < n_tasks; task += ctas_per_group) {
                const int pc = task % p.n_pchunks;
                for (int t = 0; t < p.tiles_per_chunk; ++t)
                    if ((pc * p.tiles_per_chunk + t) * kTileP < N) ++T;
            }
            mbar_wait_wd(bar_w, 0);
            uint32_t fe_ph = 0;
            int kb = 0;                                           // running last-layer block counter -> accumulator ring
            const uint32_t wl = sbase + p.smem_w[L - 1];
            auto hidden_step = [&](int l) {
                mbar_wait_wd(bar_fe, fe_ph);
                fe_ph ^= 1;
                tc_fence_after();
                const uint32_t in = sbase + (((l - 1) & 1) ? p.smem_act[1] : p.smem_act[0]);
                issue_gemm_uniform(leader, tmem + kColH, in, kTileP, sbase + p.smem_w[l], p.c[l + 1], kTileP, p.c[l + 1], p.c[l]);
                if (leader) tc_commit(bar_h);
                __syncwarp();
            };
            auto last_block = [&](int blk, uint32_t act) {
                const int a = kb % kAccBufs, use = kb / kAccBufs;
                mbar_wait_wd(bar_accempty + 8 * a, (uint32_t)((use & 1) ^ 1));
                tc_fence_after();
                issue_gemm_uniform(leader, tmem + kColAcc + (uint32_t)a * 128u, wl + (uint32_t)blk * 128 * k_last * 2, 128,
                                   act, kTileP, 128, kTileP, k_last);
                if (leader) tc_commit(bar_accfull + 8 * a);
                __syncwarp();
                ++kb;
            };
            if (T > 0)
                for (int l = 1; l <= n_hidden; ++l) hidden_step(l);   // prologue: the first tile's hidden chain
            for (int n = 0; n < T; ++n) {
                const int s = n & 1;
                mbar_wait_wd(bar_xfull + 8 * s, (uint32_t)((n >> 1) & 1));
                tc_fence_after();
                int blk = 0;
                last_block(blk++, sbase + p.smem_x[s]);
                if (n + 1 < T) {
                    for (int l = 1; l <= n_hidden; ++l) {         // next tile's hidden GEMMs ride between the blocks
                        hidden_step(l);
                        if (blk < nblk && l < n_hidden) last_block(blk++, sbase + p.smem_x[s]);
                    }
                }
                while (blk < nblk) last_block(blk++, sbase + p.smem_x[s]);
                if (leader) tc_commit(bar_xempty + 8 * s);        // X[s] may be overwritten once these MMAs are done
                __syncwarp();
            }
        }
    } else if (warp < kMmaWarp) {
        // =========================== front end ===========================
        float *w0s = reinterpret_cast<float *>(smem + p.smem_w0);
        const int c1 = p.c[1];
        for (int e = tid; e < c1 * 4; e += kFeThreads) {
            const int o = e >> 2, k = e & 3;
            w0s[e] = (k < 3) ? __ldg(p.w0 + o * 3 + k) : __ldg(p.b0 + o);
        }
        for (int l = 1; l < L - 1; ++l) {
            float *bs = reinterpret_cast<float *>(smem + p.smem_bias[l]);
            for (int e = tid; e < p.c[l + 1]; e += kFeThreads) bs[e] = __ldg(p.bias[l] + e);
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kFeThreads) : "memory");
        const int row = (warp & 3) * 32 + lane;                        // point of the tile == TMEM lane
        const int half = warp >> 2;                                    // which half of the channels this thread does
        const uint32_t lane_base = ((uint32_t)(warp & 3) * 32u) << 16; // this warp's TMEM lane quarter
        uint32_t h_ph = 0;
        int n = 0;
        // coordinates of the first tile's point; every later tile's are fetched one tile ahead
        float px = 0.f, py = 0.f, pz = 0.f;
        if (cta_in_group < n_tasks) {
            const int b = cta_in_group / p.n_pchunks, pc = cta_in_group - b * p.n_pchunks;
            const int pt = min(pc * p.tiles_per_chunk * kTileP + row, N - 1);
            const float *src = x + ((size_t)b * N + pt) * 3;
            px = __ldg(src); py = __ldg(src + 1); pz = __ldg(src + 2);
        }
        for (int task = cta_in_group; task < n_tasks; task += ctas_per_group) {
            const int b = task / p.n_pchunks, pc = task - b * p.n_pchunks;
            for (int t = 0; t < p.tiles_per_chunk; ++t) {
                const int n0 = (pc * p.tiles_per_chunk + t) * kTileP;
                if (n0 >= N) break;
                const int s = n & 1;
                const uint32_t xempty_par = (uint32_t)(((n >> 1) & 1) ^ 1);
                // prefetch the next tile's point (same task, or the first tile of this CTA's next task)
                float qx = 0.f, qy = 0.f, qz = 0.f;
                {
                    int nb = b, nn0 = n0 + kTileP;
                    bool have = (t + 1 < p.tiles_per_chunk) && (nn0 < N);
                    if (!have && task + ctas_per_group < n_tasks) {
                        const int nt = task + ctas_per_group;
                        nb = nt / p.n_pchunks;
                        nn0 = (nt - nb * p.n_pchunks) * p.tiles_per_chunk * kTileP;
                        have = true;
                    }
                    if (have) {
                        const float *src = x + ((size_t)nb * N + min(nn0 + row, N - 1)) * 3;
                        qx = __ldg(src); qy = __ldg(src + 1); qz = __ldg(src + 2);
                    }
                }
                // ---- layer 0 on CUDA cores; rows past the end repeat the last valid point (the max is unaffected)
                {
                    if (n_hidden == 0) mbar_wait_wd(bar_xempty + 8 * s, xempty_par);
                    unsigned char *dst = smem + (n_hidden == 0 ? p.smem_x[s] : p.smem_act[0]);
                    const int k8n = c1 / 16;                           // 16-byte chunks (8 channels) per thread
#pragma unroll 2
                    for (int kk = 0; kk < k8n; ++kk) {
                        const int k8 = half * k8n + kk;
                        uint32_t pk[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            float v[2];
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                const float4 w = *reinterpret_cast<const float4 *>(w0s + (k8 * 8 + q * 2 + h) * 4);
                                v[h] = fmaxf(fmaf(w.x, px, fmaf(w.y, py, fmaf(w.z, pz, w.w))), 0.0f);
                            }
                            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[0], v[1]);
                            pk[q] = *reinterpret_cast<uint32_t *>(&h2);
                        }
                        *reinterpret_cast<uint4 *>(dst + sw128_chunk_off(kTileP, row, k8)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                    fence_async_proxy();
                    mbar_arrive(n_hidden == 0 ? bar_xfull + 8 * s : bar_fe);
                }
                // ---- hidden layers: accumulator H (points on the TMEM lanes) -> bias + ReLU -> bf16 operand tile
                for (int l = 1; l <= n_hidden; ++l) {
                    const int cout = p.c[l + 1];
                    const int ncol = cout >> 1, col0 = half * ncol;    // this thread's channels: 32 or 64 of them
                    const bool last_hidden = (l == n_hidden);
                    mbar_wait_wd(bar_h, h_ph);
                    h_ph ^= 1;
                    tc_fence_after();
                    if (last_hidden) mbar_wait_wd(bar_xempty + 8 * s, xempty_par);
                    const float *bs = reinterpret_cast<const float *>(smem + p.smem_bias[l]) + col0;
                    unsigned char *dst = smem + (last_hidden ? p.smem_x[s] : ((l & 1) ? p.smem_act[1] : p.smem_act[0]));
                    float v[64];
                    if (ncol == 64) {
                        tc_ld64(tmem + lane_base + kColH + col0, v);
                    } else {
                        float v32[32];
                        tc_ld32(tmem + lane_base + kColH + col0, v32);
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = v32[i];
                    }
#pragma unroll
                    for (int q8 = 0; q8 < 8; ++q8) {
                        if (q8 * 8 < ncol) {
                            uint32_t pk[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const int ch = q8 * 8 + q * 2;
                                const float a = fmaxf(v[ch] + bs[ch], 0.0f), bb = fmaxf(v[ch + 1] + bs[ch + 1], 0.0f);
                                __nv_bfloat162 h2 = __floats2bfloat162_rn(a, bb);
                                pk[q] = *reinterpret_cast<uint32_t *>(&h2);
                            }
                            *reinterpret_cast<uint4 *>(dst + sw128_chunk_off(kTileP, row, (col0 >> 3) + q8)) =
                                make_uint4(pk[0], pk[1], pk[2], pk[3]);
                        }
                    }
                    fence_async_proxy();
                    tc_fence_before();
                    mbar_arrive(last_hidden ? bar_xfull + 8 * s : bar_fe);
                }
                px = qx; py = qy; pz = qz;
                ++n;
            }
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
