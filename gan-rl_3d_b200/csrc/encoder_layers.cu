// encoder_layers.cu -- PointNet encoder trunk, layer by layer on the tensor cores: TMA-fed tcgen05 GEMMs for ANY layer
// list whose widths are multiples of 64 up to 256 (last layer: any multiple of 64) -- in particular the reference's own
// encoder_dims 3->64->128->128->256->128 (configs/config.yaml:9-11), which the single fused kernel of encoder_tc.cu cannot
// hold on one SM (180 KB of bf16 weights next to the activation tiles).
//
// Replaces models/autoencoder.py:65-71 (transpose, [Conv1d(k=1) + BatchNorm1d + ReLU] x L, max over the points) in eval
// mode; BatchNorm is folded into the convolution by the caller (include/rlg_b200.h).
//
//   layer 0 (3 -> C1)   CUDA cores, fp32, one pass over the cloud: x (B,N,3) -> A1 [B*N x C1] in the operand format
//   layer l >= 1        Y[128 points x 128 channels] = X[128 x K] . W_slice[128 x K]^T  per CTA step:
//                         - a CTA owns one 128-channel slice of the layer and keeps its weights RESIDENT in shared memory
//                           (one TMA load per launch); it walks over the 128-point tiles of the batch
//                         - X tiles arrive by TMA (cp.async.bulk.tensor.3d, 128-byte swizzle, tensor map over
//                           (channel, point, cloud) so tiles never straddle clouds and rows past the end come in as zeros)
//                           through a ring of 64-channel K blocks; the elected MMA thread issues tcgen05.mma from the
//                           landed blocks into TMEM; two groups of four epilogue warps pull the accumulators back
//                           (tcgen05.ld), apply scale + bias + ReLU and either write the next layer's operand rows (64
//                           contiguous bytes per thread and piece, staged in shared memory and written by TMA tensor
//                           stores) or, for the last layer, take the max over the tile's points: there the MMA operands
//                           are swapped (channels on the TMEM lanes, points along the columns), so it is a per-thread
//                           running max merged into the pooled output with one atomicMax per channel and tile -- the
//                           (B, C_last, N) activation never exists.
//
// Two operand formats (mode):
//   RLG_ENC_BF16   bf16 activations and weights, fp32 accumulation: 2e-2 class (north_star's bf16 clause)
//   RLG_ENC_FP32X  fp32-grade: every activation and (pre-scaled) weight is carried as TWO fp16 numbers hi + lo
//                  (hi = fp16(v), lo = fp16(v - hi): 22 significant bits, like split-tf32 but at the full f16 MMA rate and
//                  half the bytes); three MMAs per K step (hi.hi + hi.lo + lo.hi, fp32 accumulation in TMEM), the lo.lo
//                  term (2^-22 relative) is dropped.  The tensor core's fp32 accumulation truncates, so its error is biased
//                  and grows with the number of additions into one accumulator (measured on the reference's dims: 3.6e-5 with
//                  one accumulator).  So the K steps are dealt out over FOUR accumulators (a quarter of the hi.hi steps each,
//                  at most four additions at K = 256); the cross terms (2^-11 of the result, so their own truncation is
//                  harmless) all go to the fourth one BEFORE its hi.hi steps; the epilogue adds the four in
//                  round-to-nearest (accumulators 1 and 3 carry negated products and are subtracted).  That takes all 512
//                  TMEM columns, so this mode is single-buffered: MMAs and epilogue of consecutive tiles alternate, the X
//                  tiles of the next tile are prefetched meanwhile, and both epilogue groups drain the tile together (every
//                  other 32-column chunk each, the four accumulator loads of a chunk in flight at once).  Measured: 64-column
//                  slices with two buffers overlap MMA and epilogue but read every X tile twice as often from L2 -- no gain.
//                  Weights are scaled by a power of two per layer so that max|w| lands near 2^14 (the lo parts stay normal fp16 numbers); the epilogue undoes the scale exactly.  Activations
//                  are stored unscaled: values in [2^-3, 65504] keep all 22 bits, smaller ones an absolute error <= 2^-25,
//                  larger ones saturate (post-ReLU activations of a BatchNorm-folded network are nowhere near 6.5e4).
// Between layers the activations live in HBM in exactly the format the next TMA load wants (row-major [points x C],
// 2 bytes per element and piece): the path is HBM-bound by those round trips, 6-12x faster than the stock
// torch kernels on the same GPU, and the only way the deep-narrow configuration fits the SM.
#include "common.cuh"
#include "tcgen05.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

namespace rlg {

static constexpr int kLMaxLayers = 8;
static constexpr int kLT = 128;                 // points per tile (UMMA M)
static constexpr int kLN = 128;                 // output channels per CTA slice (UMMA N)
__host__ __device__ constexpr int slice_width(int) { return kLN; }
static constexpr int kLKB = 64;                 // channels per K block: 128 bytes = one swizzle row
static constexpr uint32_t kLBlk = kLT * 128;    // bytes of one K block of one piece of an X tile (128 rows x 128 B)
static constexpr int kLThreads = 320;           // warp 0 TMA, warp 1 MMA + TMEM owner, warps 2-5 / 6-9 epilogue of buffer 0 / 1
static constexpr int kLMaxStages = 8;

struct LayerArgs {
    int B, N, K, C_out;
    int tiles_per_cloud, n_slices, stages;
    const float *bias;             // fp32 [C_out] (folded); nullable in the RAW epilogue
    float out_scale;               // 1 / weight scale (a power of two); 1 for bf16
    const float *dscale0, *dscale1;  // optional device scalars multiplied into out_scale (train mode: the operand scales are
                                     // found on the device -- no host synchronisation -- and undone here)
    void *y0, *y1;                 // EPI_ACT: next layer's operand pieces [B*N x C_out] (2-byte elements)
                                   // EPI_RAW: y0 = fp32 [B*N x C_out] (scale + bias only, no ReLU)
    float *pooled;                 // EPI_POOL: (B, C_out) fp32, zero-filled before the launch
    int tma_store;                 // EPI_ACT / EPI_RAW: staging buffers per epilogue group (0 = none: per-thread stores; 1; 2):
                                   // outputs leave through shared memory + TMA tensor stores
};
static constexpr uint32_t kStgBuf = 128 * 128;   // one staging buffer: 128 rows x 128 bytes (SWIZZLE_128B box of a TMA store)
static constexpr uint32_t kStgBytes = 4 * kStgBuf;   // two epilogue groups x two buffers
enum { EPI_ACT = 0, EPI_POOL = 1, EPI_RAW = 2 };

__device__ __forceinline__ void tma_store_3d(const CUtensorMap *tm, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void group_bar(uint32_t id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}

// cute::UMMA::InstrDescriptor: D fp32, A/B both bf16 (fmt 1) or both f16 (fmt 0), K-major
__device__ __forceinline__ uint32_t umma_idesc_16(int M, int Nn, uint32_t fmt) {
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(Nn >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// PIECES: 1 = bf16, 2 = fp16 hi/lo.  EPI: what the epilogue does with scale * acc + bias:
//   EPI_ACT ReLU and store as the next layer's operand rows; EPI_POOL ReLU and max over the points (last layer, eval);
//   EPI_RAW store as fp32 (train mode: pre-BatchNorm outputs and gradient GEMMs)
template <int PIECES, int EPI>
__global__ void __launch_bounds__(kLThreads, 1)
encoder_layer_kernel(const __grid_constant__ CUtensorMap tmx0, const __grid_constant__ CUtensorMap tmx1,
                     const __grid_constant__ CUtensorMap tmw0, const __grid_constant__ CUtensorMap tmw1,
                     const __grid_constant__ CUtensorMap tmy0, const __grid_constant__ CUtensorMap tmy1, LayerArgs a) {
    extern __shared__ unsigned char el_smem_raw[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t pad = (1024u - (smem_u32(el_smem_raw) & 1023u)) & 1023u;
    unsigned char *smem = el_smem_raw + pad;
    const uint32_t sbase = smem_u32(smem);
    const int kblocks = a.K / kLKB;
    // layout: W [PIECES][kblocks] blocks | X ring [stages][PIECES] blocks | output staging (4 buffers, with tma_store) | barriers
    constexpr int LN = slice_width(PIECES);                       // output channels of this CTA's slice
    constexpr uint32_t kWBlk = (uint32_t)LN * 128u;               // bytes of one K block of one piece of the weight slice
    const uint32_t w_off = 0, x_off = (uint32_t)PIECES * kblocks * kWBlk;
    const uint32_t stg_off = x_off + (uint32_t)a.stages * PIECES * kLBlk;
    const uint32_t bars_off = stg_off + (uint32_t)a.tma_store * 2u * kStgBuf;
    const uint32_t bars = sbase + bars_off;
    const uint32_t bar_w = bars, bar_full = bars + 8, bar_empty = bar_full + 8 * kLMaxStages;
    const uint32_t bar_accfull = bar_empty + 8 * kLMaxStages, bar_accempty = bar_accfull + 16;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + bars_off + 8 * (2 * kLMaxStages + 5));

    if (tid == 0) {
        mbar_init(bar_w, 1);
        for (int s = 0; s < a.stages; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        // bf16: two accumulator buffers, each drained by one epilogue group (128 threads); fp16 hi+lo: one buffer (its
        // four accumulators take all 512 columns) drained by both groups (256 threads), each taking every other 32-column chunk
        for (int k = 0; k < 2; ++k) { mbar_init(bar_accfull + 8 * k, 1); mbar_init(bar_accempty + 8 * k, PIECES == 1 ? 128 : 256); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    pdl_launch_dependents();
    pdl_wait();                                   // the operand rows come from the previous layer's kernel

    // this CTA: channel slice `slice`, point tiles t = first, first + step, ...   (tile = (cloud, 128-point chunk))
    const int slice = (int)blockIdx.x % a.n_slices;
    const int first = (int)blockIdx.x / a.n_slices, step = (int)gridDim.x / a.n_slices;
    const int n_tiles = a.B * a.tiles_per_cloud;

    if (warp == 0) {
        // =========================== TMA producer ===========================
        if (elect_one()) {
            mbar_arrive_expect_tx(bar_w, (uint32_t)PIECES * kblocks * kWBlk);
            for (int p = 0; p < PIECES; ++p)
                for (int kb = 0; kb < kblocks; ++kb)
                    tma_load_2d(sbase + w_off + (uint32_t)(p * kblocks + kb) * kWBlk, p ? &tmw1 : &tmw0, kb * kLKB, slice * LN, bar_w);
            uint32_t it = 0;
            for (int t = first; t < n_tiles; t += step) {
                const int b = t / a.tiles_per_cloud, n0 = (t - b * a.tiles_per_cloud) * kLT;
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const uint32_t s = it % (uint32_t)a.stages, use = it / (uint32_t)a.stages;
                    mbar_wait_wd(bar_empty + 8 * s, (use & 1u) ^ 1u);
                    mbar_arrive_expect_tx(bar_full + 8 * s, (uint32_t)PIECES * kLBlk);
                    for (int p = 0; p < PIECES; ++p)
                        tma_load_3d(sbase + x_off + (s * PIECES + p) * kLBlk, p ? &tmx1 : &tmx0, kb * kLKB, n0, b, bar_full + 8 * s);
                }
            }
        }
    } else if (warp == 1) {
        // =========================== MMA issuer ===========================
        const bool leader = elect_one();
        const uint32_t idesc = umma_idesc_16(kLT, LN, PIECES == 1 ? 1u : 0u);
        // fp16 hi+lo: the accumulators 1 and 3 collect NEGATED products (a_negate, bit 13) and are subtracted in the
        // epilogue.  The tensor core's fp32 accumulation rounds toward minus infinity, a bias that adds up coherently over
        // the points in the column sums of the training backward; with half of every sum carried negated it cancels.
        const uint32_t idesc_neg = idesc | (1u << 13);
        mbar_wait_wd(bar_w, 0);
        uint32_t it = 0, ti = 0;
        for (int t = first; t < n_tiles; t += step, ++ti) {
            // bf16: two 128-column accumulator buffers alternate between tiles (the MMAs of tile t+1 run under the epilogue of
            // tile t).  fp16 hi+lo: ONE set of four 128-column accumulators: the hi.hi steps in four contiguous quarters, the
            // cross terms into the fourth before its own hi.hi steps.
            const uint32_t acc = PIECES == 1 ? (ti & 1u) : 0u;
            const uint32_t par = PIECES == 1 ? ((ti >> 1) & 1u) : (ti & 1u);
            mbar_wait_wd(bar_accempty + 8 * acc, par ^ 1u);
            tc_fence_after();
            const uint32_t base = tmem + acc * (uint32_t)kLN;
            const uint32_t d = base, dx = base + 3u * (uint32_t)LN;
            const uint32_t n_steps = (uint32_t)a.K / 16u, q_steps = (n_steps + 3u) / 4u;   // hi.hi steps per accumulator
            uint32_t accum = 0, used = 0, ks = 0;
            for (int kb = 0; kb < kblocks; ++kb, ++it) {
                const uint32_t s = it % (uint32_t)a.stages, use = it / (uint32_t)a.stages;
                mbar_wait_wd(bar_full + 8 * s, use & 1u);
                tc_fence_after();
                if (leader) {
                    const uint64_t x0 = umma_desc(sbase + x_off + (s * PIECES) * kLBlk);
                    const uint64_t w0 = umma_desc(sbase + w_off + (uint32_t)kb * kWBlk);
                    const uint64_t xp = (uint64_t)(kLBlk >> 4), wp = (uint64_t)(((uint32_t)kblocks * kWBlk) >> 4);   // piece strides
                    if (PIECES == 2) {
                        // this block's cross terms first: accumulator 3 must have all of them before its own hi.hi steps
#pragma unroll
                        for (int k4 = 0; k4 < 4; ++k4) {
                            const uint64_t ko = (uint64_t)(k4 * 2);
                            if (EPI == EPI_POOL) {                                            // channels on the TMEM lanes
                                tc_mma_bf16(dx, w0 + ko, x0 + xp + ko, idesc_neg, accum);
                                tc_mma_bf16(dx, w0 + wp + ko, x0 + ko, idesc_neg, 1u);
                            } else {
                                tc_mma_bf16(dx, x0 + xp + ko, w0 + ko, idesc_neg, accum);     // -(lo . hi)
                                tc_mma_bf16(dx, x0 + ko, w0 + wp + ko, idesc_neg, 1u);        // -(hi . lo)
                            }
                            accum = 1;
                        }
                        used |= 8u;
                    }
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4, ++ks) {    // 16 elements = 32 bytes per MMA along K
                        const uint64_t ko = (uint64_t)(k4 * 2);
                        // EPI_POOL: operands swapped (A = weights): the accumulator holds channels on the TMEM lanes and the
                        // tile's points along the columns, so the max over the points is a per-thread running max
                        const uint64_t da = EPI == EPI_POOL ? w0 + ko : x0 + ko, db = EPI == EPI_POOL ? x0 + ko : w0 + ko;
                        if (PIECES == 1) {
                            tc_mma_bf16(d, da, db, idesc, accum);
                            accum = 1;
                        } else {
                            const uint32_t m = min(3u, ks / q_steps);
                            tc_mma_bf16(base + m * (uint32_t)LN, da, db, (m & 1u) ? idesc_neg : idesc, (used >> m) & 1u);   // +-(hi . hi)
                            used |= 1u << m;
                        }
                    }
                    tc_commit(bar_empty + 8 * s);
                }
                __syncwarp();
            }
            if (leader) tc_commit(bar_accfull + 8 * acc);
            __syncwarp();
        }
    } else {
        // =========================== epilogue ===========================
        const int q = warp & 3;                                        // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;                                 // point of the tile
        const uint32_t lane_base = ((uint32_t)q * 32u) << 16;
        const int c_base = slice * LN;
        // two epilogue groups of four warps (two warps per SM sub-partition: a lone epilogue warp per scheduler cannot hide
        // its own latencies).  bf16: group g drains accumulator buffer g, i.e. the CTA's tiles ti = g, g + 2, ...;
        // fp16 hi+lo: both groups drain every tile, group g taking the 64-column half g (chunks 2g, 2g + 1).
        // Outputs (EPI_ACT / EPI_RAW) leave through two 16 KB staging buffers per group, written in the 128-byte-swizzled box
        // layout and pushed out by TMA tensor stores: per-thread row stores would reach HBM as 16-byte partial sectors
        // (measured: more than half of the kernel's time).
        const uint32_t grp = (uint32_t)(warp - 2) >> 2;
        const int gtid = tid - 64 - (int)grp * 128;                    // thread of the group
        constexpr uint32_t kTileStride = PIECES == 1 ? 2u : 1u;
        constexpr bool POOL = EPI == EPI_POOL;
        const bool staged = !POOL && a.tma_store != 0;
        const uint32_t nbuf = (uint32_t)a.tma_store;                   // staging buffers of this group
        const uint32_t stg = sbase + stg_off + grp * nbuf * kStgBuf;
        const uint32_t bar_id = 1u + grp;                              // named barrier of the group
        const uint32_t swz = (uint32_t)row * 128u, rx = (uint32_t)row & 7u;
        const float sc = a.out_scale * (a.dscale0 ? __ldg(a.dscale0) : 1.0f) * (a.dscale1 ? __ldg(a.dscale1) : 1.0f);
        uint32_t ti = PIECES == 1 ? grp : 0u;
        for (int t = first + (int)ti * step; t < n_tiles; t += (int)kTileStride * step, ti += kTileStride) {
            const uint32_t acc = PIECES == 1 ? (ti & 1u) : 0u;
            const uint32_t par = PIECES == 1 ? ((ti >> 1) & 1u) : (ti & 1u);
            const uint32_t tbase = tmem + lane_base + acc * (uint32_t)kLN;
            const int b = t / a.tiles_per_cloud, n0 = (t - b * a.tiles_per_cloud) * kLT;
            const bool valid = n0 + row < a.N;
            const size_t grow = (size_t)b * a.N + n0 + row;
            // a store unit = what fills the group's staging buffers once: with two buffers the group's whole share of the tile,
            // with one buffer a single box (32 fp32 columns, or 64 bf16 columns)
            auto unit_begin = [&]() {
                if (gtid == 0) tma_store_wait_read();                  // the previous unit's stores have read their buffers
                group_bar(bar_id);
            };
            auto unit_end = [&](int cfirst, int n_boxes) {             // boxes: consecutive buffers, consecutive column blocks
                fence_async_proxy();                                   // the staged rows become visible to the TMA engine
                group_bar(bar_id);
                if (gtid == 0) {
                    const int cg = c_base + cfirst * 32;
                    if (EPI != EPI_RAW && PIECES == 2) {               // hi -> buffer 0, lo -> buffer 1, same columns
                        if (cg < a.C_out) { tma_store_3d(&tmy0, stg, cg, n0, b); tma_store_3d(&tmy1, stg + kStgBuf, cg, n0, b); }
                    } else {
                        const int wcols = EPI == EPI_RAW ? 32 : 64;
                        for (int k = 0; k < n_boxes; ++k)
                            if (cg + k * wcols < a.C_out) tma_store_3d(&tmy0, stg + (uint32_t)k * kStgBuf, cg + k * wcols, n0, b);
                    }
                    tma_store_commit();
                }
            };
            if (staged && nbuf == 2) unit_begin();
            mbar_wait_wd(bar_accfull + 8 * acc, par);
            tc_fence_after();
            const int c_begin = PIECES == 1 ? 0 : 2 * (int)grp, c_end = PIECES == 1 ? LN / 32 : c_begin + 2;
            if (POOL) {
                // this thread = output channel c_base + row; columns = the tile's points (rows past the end of the cloud
                // came in as zeros and are masked out); bias and ReLU commute with the max and are applied once
                const int n_valid = min(kLT, a.N - n0);
                float mx = -INFINITY;
#pragma unroll 1
                for (int c = c_begin; c < c_end; ++c) {
                    const int nv = n_valid - c * 32;
                    if (nv <= 0) break;                                // warp-uniform
                    float v[32];
                    if (PIECES == 1) {
                        tc_ld32(tbase + (uint32_t)(c * 32), v);
                    } else {
                        float v1[32], v2[32], v3[32];
                        tc_ld32_nowait(tbase + (uint32_t)(c * 32), v);
                        tc_ld32_nowait(tbase + (uint32_t)LN + (uint32_t)(c * 32), v1);
                        tc_ld32_nowait(tbase + 2u * (uint32_t)LN + (uint32_t)(c * 32), v2);
                        tc_ld32_nowait(tbase + 3u * (uint32_t)LN + (uint32_t)(c * 32), v3);
                        tc_wait_ld(v); tc_wait_ld(v1); tc_wait_ld(v2); tc_wait_ld(v3);
#pragma unroll
                        for (int e = 0; e < 32; ++e) v[e] = ((v[e] - v1[e]) + v2[e]) - v3[e];
                    }
                    if (nv >= 32) {
#pragma unroll
                        for (int e = 0; e < 32; e += 2) mx = max3(mx, v[e], v[e + 1]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; ++e) mx = fmaxf(mx, e < nv ? v[e] : -INFINITY);
                    }
                }
                tc_fence_before();
                mbar_arrive(bar_accempty + 8 * acc);
                const int ch = c_base + row;
                if (ch < a.C_out && mx > -INFINITY) {
                    const float val = fmaxf(fmaf(mx, sc, __ldg(a.bias + ch)), 0.0f);
                    if (val > 0.0f) atomicMax(reinterpret_cast<unsigned *>(a.pooled) + (size_t)b * a.C_out + ch, __float_as_uint(val));
                }
                continue;
            }
#pragma unroll 1
            for (int c = c_begin; c < c_end; ++c) {
                const int col0 = c_base + c * 32;
                if (col0 >= a.C_out) break;                            // warp-uniform
                float v[32];
                if (PIECES == 1) {
                    tc_ld32(tbase + (uint32_t)(c * 32), v);
                } else {
                    // the four partial accumulators (K >= 64: all four are in use): their loads are in flight together,
                    // every addition in round-to-nearest, accumulators 1 and 3 hold negated sums
                    float v1[32], v2[32], v3[32];
                    tc_ld32_nowait(tbase + (uint32_t)(c * 32), v);
                    tc_ld32_nowait(tbase + (uint32_t)LN + (uint32_t)(c * 32), v1);
                    tc_ld32_nowait(tbase + 2u * (uint32_t)LN + (uint32_t)(c * 32), v2);
                    tc_ld32_nowait(tbase + 3u * (uint32_t)LN + (uint32_t)(c * 32), v3);
                    tc_wait_ld(v); tc_wait_ld(v1); tc_wait_ld(v2); tc_wait_ld(v3);
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = ((v[e] - v1[e]) + v2[e]) - v3[e];
                }
                if (EPI == EPI_RAW) {
                    if (a.bias != nullptr) {
                        const float4 *bp = reinterpret_cast<const float4 *>(a.bias + col0);
#pragma unroll
                        for (int e4 = 0; e4 < 8; ++e4) {
                            const float4 bb = __ldg(bp + e4);
                            v[4 * e4] = fmaf(v[4 * e4], sc, bb.x); v[4 * e4 + 1] = fmaf(v[4 * e4 + 1], sc, bb.y);
                            v[4 * e4 + 2] = fmaf(v[4 * e4 + 2], sc, bb.z); v[4 * e4 + 3] = fmaf(v[4 * e4 + 3], sc, bb.w);
                        }
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; ++e) v[e] *= sc;
                    }
                    if (staged) {
                        // a chunk of 32 fp32 columns is one 128-byte row of a box: buffer (c & 1) of the group (or its only one)
                        if (nbuf == 1) unit_begin();
                        const uint32_t dst = stg + (nbuf == 2 ? (uint32_t)(c & 1) : 0u) * kStgBuf + swz;
#pragma unroll
                        for (uint32_t j = 0; j < 8; ++j)
                            st_shared_v4(dst + ((j ^ rx) << 4), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                                         __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
                        if (nbuf == 1) unit_end(c, 1);
                    } else if (valid) {
                        float4 *dst = reinterpret_cast<float4 *>(reinterpret_cast<float *>(a.y0) + grow * a.C_out + col0);
#pragma unroll
                        for (int e4 = 0; e4 < 8; ++e4) dst[e4] = make_float4(v[4 * e4], v[4 * e4 + 1], v[4 * e4 + 2], v[4 * e4 + 3]);
                    }
                    continue;
                }
                const float4 *bp = reinterpret_cast<const float4 *>(a.bias + col0);
#pragma unroll
                for (int e4 = 0; e4 < 8; ++e4) {
                    const float4 bb = __ldg(bp + e4);
                    v[4 * e4] = fmaxf(fmaf(v[4 * e4], sc, bb.x), 0.0f);
                    v[4 * e4 + 1] = fmaxf(fmaf(v[4 * e4 + 1], sc, bb.y), 0.0f);
                    v[4 * e4 + 2] = fmaxf(fmaf(v[4 * e4 + 2], sc, bb.z), 0.0f);
                    v[4 * e4 + 3] = fmaxf(fmaf(v[4 * e4 + 3], sc, bb.w), 0.0f);
                }
                if (staged) {
                    // 32 columns of 2-byte elements = 64 bytes = half a box row: 16-byte chunks 4 * (c & 1) .. + 3.
                    // bf16: the tile's two 64-column halves go to buffers 0 / 1; fp16: hi -> buffer 0, lo -> buffer 1
                    const uint32_t j0 = 4u * (uint32_t)(c & 1);
                    if (PIECES == 1) {
                        if (nbuf == 1 && (c & 1) == 0) unit_begin();
                        const uint32_t dst = stg + (nbuf == 2 ? (uint32_t)(c >> 1) : 0u) * kStgBuf + swz;
#pragma unroll
                        for (uint32_t e8 = 0; e8 < 4; ++e8)
                            st_shared_v4(dst + (((j0 + e8) ^ rx) << 4), pack_bf16x2(v[8 * e8], v[8 * e8 + 1]), pack_bf16x2(v[8 * e8 + 2], v[8 * e8 + 3]),
                                         pack_bf16x2(v[8 * e8 + 4], v[8 * e8 + 5]), pack_bf16x2(v[8 * e8 + 6], v[8 * e8 + 7]));
                        if (nbuf == 1 && (c & 1) == 1) unit_end(c - 1, 1);
                    } else {
#pragma unroll
                        for (uint32_t e8 = 0; e8 < 4; ++e8) {
                            uint32_t h[4], l[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) split_f16x2(v[8 * e8 + 2 * k], v[8 * e8 + 2 * k + 1], h[k], l[k]);
                            const uint32_t o = swz + (((j0 + e8) ^ rx) << 4);
                            st_shared_v4(stg + o, h[0], h[1], h[2], h[3]);
                            st_shared_v4(stg + kStgBuf + o, l[0], l[1], l[2], l[3]);
                        }
                    }
                } else if (valid) {
                    if (PIECES == 1) {
                        uint4 *dst = reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(a.y0) + (grow * a.C_out + col0) * 2);
#pragma unroll
                        for (int e8 = 0; e8 < 4; ++e8)
                            dst[e8] = make_uint4(pack_bf16x2(v[8 * e8], v[8 * e8 + 1]), pack_bf16x2(v[8 * e8 + 2], v[8 * e8 + 3]),
                                                 pack_bf16x2(v[8 * e8 + 4], v[8 * e8 + 5]), pack_bf16x2(v[8 * e8 + 6], v[8 * e8 + 7]));
                    } else {
                        uint4 *dh = reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(a.y0) + (grow * a.C_out + col0) * 2);
                        uint4 *dl = reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(a.y1) + (grow * a.C_out + col0) * 2);
#pragma unroll
                        for (int e8 = 0; e8 < 4; ++e8) {
                            uint32_t h[4], l[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) split_f16x2(v[8 * e8 + 2 * k], v[8 * e8 + 2 * k + 1], h[k], l[k]);
                            dh[e8] = make_uint4(h[0], h[1], h[2], h[3]);
                            dl[e8] = make_uint4(l[0], l[1], l[2], l[3]);
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(bar_accempty + 8 * acc);                       // the accumulator goes back to the issuer
            if (staged && nbuf == 2) unit_end(c_begin, 2);
        }
        if (staged && gtid == 0) tma_store_wait_all();                 // the staging buffers must outlive their stores
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

// layer 0 on the CUDA cores: x (B,N,3) fp32 -> operand rows [B*N x C1]; one thread per (point, 8 channels)
template <int PIECES>
__global__ void __launch_bounds__(256) encoder_layer0_kernel(const float *__restrict__ x, long long P, int C1,
                                                            const float *__restrict__ w, const float *__restrict__ bias,
                                                            void *__restrict__ y0, void *__restrict__ y1) {
    pdl_launch_dependents();
    pdl_wait();
    const int chunks = C1 / 8;
    const long long e = (long long)blockIdx.x * 256 + threadIdx.x;
    if (e >= P * chunks) return;
    const long long p = e / chunks;
    const int c0 = (int)(e - p * chunks) * 8;
    const float px = __ldg(x + 3 * p), py = __ldg(x + 3 * p + 1), pz = __ldg(x + 3 * p + 2);
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float *wr = w + 3 * (c0 + k);
        // the operation order of the fp32 kernel (encoder_fp32.cu): bias + x*w0 + y*w1 + z*w2, then ReLU
        v[k] = fmaxf(fmaf(pz, __ldg(wr + 2), fmaf(py, __ldg(wr + 1), fmaf(px, __ldg(wr), __ldg(bias + c0 + k)))), 0.0f);
    }
    const size_t off = ((size_t)p * C1 + c0) * 2;
    if (PIECES == 1) {
        *reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(y0) + off) =
            make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
    } else {
        uint32_t h[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            h[k] = pack_f16x2_sat(v[2 * k], v[2 * k + 1]);
            const float2 hf = unpack_f16x2(h[k]);
            l[k] = pack_f16x2_sat(v[2 * k] - hf.x, v[2 * k + 1] - hf.y);
        }
        *reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(y0) + off) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4 *>(reinterpret_cast<unsigned char *>(y1) + off) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}

// folded fp32 weights (c_out, c_in) -> operand pieces, row-major [c_out x c_in], 2 bytes per element, scaled by `scale`
template <int PIECES>
__global__ void __launch_bounds__(256) encoder_layers_pack_kernel(const float *__restrict__ w, int n, float scale,
                                                                 unsigned short *__restrict__ p0, unsigned short *__restrict__ p1) {
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= n) return;
    const float v = w[e] * scale;
    if (PIECES == 1) {
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        p0[e] = *reinterpret_cast<const unsigned short *>(&h);
    } else {
        const __half h = __float2half_rn(v);
        const __half l = __float2half_rn(v - __half2float(h));
        p0[e] = *reinterpret_cast<const unsigned short *>(&h);
        p1[e] = *reinterpret_cast<const unsigned short *>(&l);
    }
}

// ---- host side -------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point: no link-time dependency on libcuda, so the library
// still loads (and exports its symbols) on a machine without a driver
static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    fn = (EncodeTiledFn)p;
    return fn;
}

int make_map(CUtensorMap *tm, int pieces_fmt, void *base, int rank, const cuuint64_t *dims, const cuuint64_t *strides,
                    const cuuint32_t *box) {
    EncodeTiledFn f = encode_tiled();
    if (!f) return fail(RLG_ERR_UNSUPPORTED, "rlg_encoder_gemm: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint32_t es[3] = {1, 1, 1};
    const CUtensorMapDataType dt = pieces_fmt == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                   : pieces_fmt == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUresult r = f(tm, dt, (cuuint32_t)rank, base, dims,
                   strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(RLG_ERR_UNSUPPORTED, "rlg_encoder_gemm: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

// one layer GEMM: Y = epi(X[B*N x K] . W[C_out x K]^T), operands as 2-byte piece arrays (see the file header)
int launch_layer_gemm(const GemmCall &g, int sms, cudaStream_t st) {
    const int pieces = g.pieces;
    if (g.K % kLKB != 0 || g.K < kLKB || g.K > 256 || g.C_out % 64 != 0 || g.C_out < 64)
        return fail(RLG_ERR_UNSUPPORTED, "encoder_layer_kernel: K=%d / C_out=%d not covered (multiples of 64, K <= 256)", g.K, g.C_out);
    LayerArgs a;
    a.B = g.B; a.N = g.N; a.K = g.K; a.C_out = g.C_out;
    a.tiles_per_cloud = (g.N + kLT - 1) / kLT;
    const int LN = slice_width(pieces);
    a.n_slices = (g.C_out + LN - 1) / LN;
    a.bias = g.bias;
    a.out_scale = g.out_scale;
    a.dscale0 = g.dscale0; a.dscale1 = g.dscale1;
    a.y0 = g.y0; a.y1 = g.y1; a.pooled = g.pooled;
    const int kblocks = g.K / kLKB;
    const size_t w_bytes = (size_t)pieces * kblocks * LN * 128, stage_bytes = (size_t)pieces * kLBlk;
    const size_t budget = 226u * 1024u - 256u;             // 227 KB per CTA minus the alignment slack and the barriers
    // outputs through shared-memory staging + TMA stores when the staging buffers leave room for a two-stage X ring
    a.tma_store = 0;
    if (g.epi != EPI_POOL) {
        const long long room = (long long)budget - (long long)w_bytes - 2 * (long long)stage_bytes;
        if (room >= (long long)kStgBytes) a.tma_store = 2;
        else if (room >= (long long)kStgBytes / 2 && !(g.epi == EPI_ACT && pieces == 2)) a.tma_store = 1;   // hi + lo need two
    }
    const size_t stg_bytes = (size_t)a.tma_store * 2 * kStgBuf;
    long long stages = ((long long)budget - (long long)w_bytes - (long long)stg_bytes) / (long long)stage_bytes;
    if (stages < 2) return fail(RLG_ERR_UNSUPPORTED, "encoder_layer_kernel: K=%d does not fit in shared memory", g.K);
    if (stages > kLMaxStages) stages = kLMaxStages;
    a.stages = (int)stages;
    const size_t smem_bytes = w_bytes + (size_t)stages * stage_bytes + stg_bytes + 256 + 1024;
    CUtensorMap tmx[2], tmw[2], tmy[2];
    for (int p = 0; p < 2; ++p) {
        const bool second = p == 1 && pieces == 2;
        const cuuint64_t xd[3] = {(cuuint64_t)g.K, (cuuint64_t)g.N, (cuuint64_t)g.B};
        const cuuint64_t xs[2] = {(cuuint64_t)g.K * 2, (cuuint64_t)g.N * g.K * 2};
        const cuuint32_t xb[3] = {(cuuint32_t)kLKB, (cuuint32_t)kLT, 1};
        int rc = make_map(&tmx[p], pieces, const_cast<void *>(second ? g.x1 : g.x0), 3, xd, xs, xb);
        if (rc) return rc;
        const cuuint64_t wd[2] = {(cuuint64_t)g.K, (cuuint64_t)g.C_out};
        const cuuint64_t wst[1] = {(cuuint64_t)g.K * 2};
        const cuuint32_t wb[2] = {(cuuint32_t)kLKB, (cuuint32_t)LN};
        rc = make_map(&tmw[p], pieces, const_cast<void *>(second ? g.w1 : g.w0), 2, wd, wst, wb);
        if (rc) return rc;
        tmy[p] = tmx[p];                                   // placeholder when nothing is stored through TMA
        if (a.tma_store) {
            // output boxes: 128 rows x 128 bytes = 64 two-byte columns (EPI_ACT) or 32 fp32 columns (EPI_RAW)
            const bool raw = g.epi == EPI_RAW;
            const size_t esz = raw ? 4 : 2;
            void *ybase = raw ? g.y0 : (second ? g.y1 : g.y0);
            const cuuint64_t yd[3] = {(cuuint64_t)g.C_out, (cuuint64_t)g.N, (cuuint64_t)g.B};
            const cuuint64_t ys[2] = {(cuuint64_t)g.C_out * esz, (cuuint64_t)g.N * g.C_out * esz};
            const cuuint32_t yb[3] = {(cuuint32_t)(raw ? 32 : 64), (cuuint32_t)kLT, 1};
            rc = make_map(&tmy[p], raw ? 3 : pieces, ybase, 3, yd, ys, yb);
            if (rc) return rc;
        }
    }
    const long long tiles = (long long)g.B * a.tiles_per_cloud;
    long long per_slice = sms / a.n_slices;
    if (per_slice < 1) per_slice = 1;
    if (per_slice > tiles) per_slice = tiles;
    const unsigned grid = (unsigned)(per_slice * a.n_slices);
    auto launch = [&](auto kernel) -> int {
        cudaError_t ae = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (ae != cudaSuccess) { cudaGetLastError(); return fail((int)ae, "encoder_layer_kernel: cudaFuncSetAttribute(%zu): %s", smem_bytes, cudaGetErrorString(ae)); }
        cudaError_t le = launch_pdl(kernel, dim3(grid), dim3(kLThreads), smem_bytes, st, tmx[0], tmx[1], tmw[0], tmw[1], tmy[0], tmy[1], a);
        if (le != cudaSuccess) { cudaGetLastError(); return fail((int)le, "encoder_layer_kernel: %s", cudaGetErrorString(le)); }
        return 0;
    };
    if (pieces == 1) {
        if (g.epi == EPI_POOL) return launch(encoder_layer_kernel<1, EPI_POOL>);
        if (g.epi == EPI_ACT) return launch(encoder_layer_kernel<1, EPI_ACT>);
        return fail(RLG_ERR_UNSUPPORTED, "encoder_layer_kernel: the raw epilogue exists for the fp16 hi+lo operands only");
    }
    if (g.epi == EPI_POOL) return launch(encoder_layer_kernel<2, EPI_POOL>);
    if (g.epi == EPI_ACT) return launch(encoder_layer_kernel<2, EPI_ACT>);
    return launch(encoder_layer_kernel<2, EPI_RAW>);
}

static int gemm_check(const rlg_layer *layers, int L, int mode) {
    if (mode != RLG_ENC_BF16 && mode != RLG_ENC_FP32X) return fail(RLG_ERR_UNSUPPORTED, "rlg_encoder_gemm: unknown mode %d", mode);
    if (!layers || L < 2 || L > kLMaxLayers) return fail(RLG_ERR_UNSUPPORTED, "rlg_encoder_gemm: need 2..%d layers, got %d", kLMaxLayers, L);
    if (layers[0].c_in != 3) return fail(RLG_ERR_UNSUPPORTED, "rlg_encoder_gemm: layer 0 must have c_in == 3");
    for (int l = 0; l < L; ++l) {
        if (!layers[l].w || !layers[l].b) return fail(RLG_ERR_NULL_POINTER, "rlg_encoder_gemm: layer %d has null weights", l);
        if (l > 0 && layers[l].c_in != layers[l - 1].c_out)
            return fail(RLG_ERR_BAD_SHAPE, "rlg_encoder_gemm: layer %d c_in %d != previous c_out %d", l, layers[l].c_in, layers[l - 1].c_out);
        if (layers[l].c_out % 64 != 0 || layers[l].c_out < 64)
            return fail(RLG_ERR_UNSUPPORTED, "rlg_encoder_gemm: width %d is not a multiple of 64 (use the fp32 path)", layers[l].c_out);
        if (l < L - 1 && layers[l].c_out > 256)
            return fail(RLG_ERR_UNSUPPORTED, "rlg_encoder_gemm: hidden width %d exceeds 256 (resident weights; use the fp32 path)", layers[l].c_out);
    }
    return 0;
}

static size_t pack_layer_off(const rlg_layer *layers, int l, int pieces) {     // byte offset of layer l's pieces (l >= 1)
    size_t off = 0;
    for (int k = 1; k < l; ++k) off += align_up((size_t)layers[k].c_out * layers[k].c_in * 2, 256) * pieces;
    return off;
}

}  // namespace rlg

using namespace rlg;

extern "C" {

size_t rlg_encoder_gemm_pack_bytes(const rlg_layer *layers, int L, int mode) {
    if (gemm_check(layers, L, mode) != 0) return 0;
    return pack_layer_off(layers, L, mode == RLG_ENC_BF16 ? 1 : 2);
}

int rlg_encoder_gemm_pack(const rlg_layer *layers, int L, int mode, const float *weight_scales, void *packed, size_t packed_bytes,
                          void *stream) {
    int rc = gemm_check(layers, L, mode);
    if (rc) return rc;
    const int pieces = mode == RLG_ENC_BF16 ? 1 : 2;
    const size_t total = pack_layer_off(layers, L, pieces);
    if (!packed || packed_bytes < total || ((uintptr_t)packed & 255u))
        return fail(RLG_ERR_WORKSPACE, "rlg_encoder_gemm_pack: buffer %p/%zu bytes, need %zu bytes 256-B aligned", packed, packed_bytes, total);
    if (!weight_scales) return fail(RLG_ERR_NULL_POINTER, "rlg_encoder_gemm_pack: null weight_scales");
    cudaStream_t st = (cudaStream_t)stream;
    for (int l = 1; l < L; ++l) {
        const int n = layers[l].c_out * layers[l].c_in;
        unsigned short *p0 = (unsigned short *)((char *)packed + pack_layer_off(layers, l, pieces));
        unsigned short *p1 = (unsigned short *)((char *)p0 + align_up((size_t)n * 2, 256));
        if (pieces == 1) encoder_layers_pack_kernel<1><<<(n + 255) / 256, 256, 0, st>>>(layers[l].w, n, 1.0f, p0, nullptr);
        else encoder_layers_pack_kernel<2><<<(n + 255) / 256, 256, 0, st>>>(layers[l].w, n, weight_scales[l], p0, p1);
    }
    return check_launch("encoder_layers_pack_kernel");
}

// two ping-pong activation buffers of B*N x max hidden width, 2 bytes per element and piece
size_t rlg_encoder_gemm_ws_bytes(int B, int N, const rlg_layer *layers, int L, int mode) {
    if (B < 0 || N < 1 || gemm_check(layers, L, mode) != 0) return 0;
    int cmax = 0;
    for (int l = 0; l < L - 1; ++l) cmax = layers[l].c_out > cmax ? layers[l].c_out : cmax;
    const int pieces = mode == RLG_ENC_BF16 ? 1 : 2;
    return 2 * (size_t)pieces * align_up((size_t)B * N * cmax * 2, 256);
}

int rlg_encoder_gemm_fwd(const float *x, int B, int N, const rlg_layer *layers, int L, int mode, const float *weight_scales,
                         const void *packed, size_t packed_bytes, float *pooled, void *ws, size_t ws_bytes, void *stream) {
    if (B < 0 || N < 1) return fail(RLG_ERR_BAD_SHAPE, "rlg_encoder_gemm_fwd: bad shape B=%d N=%d", B, N);
    int rc = gemm_check(layers, L, mode);
    if (rc) return rc;
    if (B == 0) return 0;
    if (!x || !pooled || !packed || !weight_scales) return fail(RLG_ERR_NULL_POINTER, "rlg_encoder_gemm_fwd: null pointer");
    if ((long long)B * N > 0x7fffffffLL / 256) return fail(RLG_ERR_TOO_LARGE, "rlg_encoder_gemm_fwd: B*N too large");
    const int pieces = mode == RLG_ENC_BF16 ? 1 : 2;
    const size_t need_pack = pack_layer_off(layers, L, pieces), need_ws = rlg_encoder_gemm_ws_bytes(B, N, layers, L, mode);
    if (packed_bytes < need_pack || ((uintptr_t)packed & 255u))
        return fail(RLG_ERR_WORKSPACE, "rlg_encoder_gemm_fwd: packed weights %zu bytes, need %zu (256-B aligned)", packed_bytes, need_pack);
    if (!ws || ws_bytes < need_ws || ((uintptr_t)ws & 255u))
        return fail(RLG_ERR_WORKSPACE, "rlg_encoder_gemm_fwd: workspace %p/%zu bytes, need %zu bytes 256-B aligned", ws, ws_bytes, need_ws);
    const int sms = sm_count();
    if (sms <= 0) return fail((int)cudaErrorNoDevice, "rlg_encoder_gemm_fwd: no CUDA device");
    cudaStream_t st = (cudaStream_t)stream;
    const long long P = (long long)B * N;
    int cmax = 0;
    for (int l = 0; l < L - 1; ++l) cmax = layers[l].c_out > cmax ? layers[l].c_out : cmax;
    const size_t buf_bytes = align_up((size_t)P * cmax * 2, 256);
    auto act = [&](int which, int piece) { return (char *)ws + ((size_t)which * pieces + piece) * buf_bytes; };

    cudaError_t e = cudaMemsetAsync(pooled, 0, sizeof(float) * (size_t)B * layers[L - 1].c_out, st);
    if (e != cudaSuccess) { cudaGetLastError(); return fail((int)e, "rlg_encoder_gemm_fwd: cudaMemsetAsync: %s", cudaGetErrorString(e)); }
    // layer 0
    {
        const int C1 = layers[0].c_out;
        const long long n = P * (C1 / 8);
        const unsigned blocks = (unsigned)((n + 255) / 256);
        cudaError_t le = pieces == 1 ? launch_pdl(encoder_layer0_kernel<1>, dim3(blocks), dim3(256), (size_t)0, st, x, P, C1, layers[0].w, layers[0].b,
                                                  (void *)act(0, 0), (void *)nullptr)
                                     : launch_pdl(encoder_layer0_kernel<2>, dim3(blocks), dim3(256), (size_t)0, st, x, P, C1, layers[0].w, layers[0].b,
                                                  (void *)act(0, 0), (void *)act(0, 1));
        if (le != cudaSuccess) { cudaGetLastError(); return fail((int)le, "encoder_layer0_kernel: %s", cudaGetErrorString(le)); }
    }
    int cur = 0;
    for (int l = 1; l < L; ++l) {
        const bool last = l == L - 1;
        const int K = layers[l].c_in, C = layers[l].c_out;
        GemmCall g;
        g.B = B; g.N = N; g.K = K; g.C_out = C; g.pieces = pieces;
        g.x0 = act(cur, 0); g.x1 = pieces == 2 ? act(cur, 1) : nullptr;
        g.w0 = (char *)packed + pack_layer_off(layers, l, pieces);
        g.w1 = pieces == 2 ? (char *)g.w0 + align_up((size_t)C * K * 2, 256) : nullptr;
        g.bias = layers[l].b;
        g.out_scale = pieces == 1 ? 1.0f : 1.0f / weight_scales[l];
        g.dscale0 = g.dscale1 = nullptr;
        g.epi = last ? EPI_POOL : EPI_ACT;
        g.y0 = last ? nullptr : act(cur ^ 1, 0);
        g.y1 = last || pieces == 1 ? nullptr : act(cur ^ 1, 1);
        g.pooled = last ? pooled : nullptr;
        rc = launch_layer_gemm(g, sms, st);
        if (rc) return rc;
        cur ^= 1;
    }
    return check_launch("encoder_layer_kernel");
}

}  // extern "C"
