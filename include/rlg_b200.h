/*
 * rlg_b200.h -- C ABI of librlg_b200.so: the B200 (sm_100a) hot path of RL-GAN-Net point-cloud
 * completion (phanich004/GAN-RL_3D): batched Chamfer distance and the PointNet encoder + max-pool.
 *
 * The reference is pure Python/PyTorch and has no FFI of its own; these entry points are what a
 * binding for its two choke points would call (citations into /root/reference):
 *
 *   rlg_chamfer_fwd   replaces  utils/losses.py:29-37   (torch.cdist -> min x2 -> mean) inside
 *                               chamfer_distance_l2 (utils/losses.py:13-39)
 *   rlg_chamfer_bwd   replaces  the autograd graph of the above (MeanBackward, MinBackward0 x2,
 *                               EuclideanDistBackward0) reached from loss.backward()
 *                               (train_rl_gan_net.py:239)
 *   rlg_encoder_*     replaces  models/autoencoder.py:65-71 (transpose, point_mlp, max over points)
 *                               inside PointNetEncoder.forward (models/autoencoder.py:56-76), eval mode
 *
 * Conventions
 *   - extern "C", C types only.  Every pointer marked "device" is a CUDA device pointer on the
 *     CURRENT device; the caller owns every buffer including the workspace.  The library never
 *     allocates or frees device memory, never synchronises, never copies host<->device, and
 *     enqueues all work on `stream` (a cudaStream_t passed as void*; NULL = legacy default
 *     stream), so every call is CUDA-graph capturable.
 *   - Return value: 0 = ok; negative = argument error detected on the host before any launch
 *     (RLG_ERR_*); positive = the cudaError_t of a failed launch.  rlg_last_error() returns a
 *     thread-local message for the last non-zero return.
 *   - All tensors are contiguous, row-major, fp32 unless noted; indices are int32.
 *   - There is no CPU implementation behind this ABI: without a CUDA device every compute call
 *     fails with a positive cudaError_t.
 */
#ifndef RLG_B200_H
#define RLG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RLG_ABI_VERSION 7

#define RLG_ERR_NULL_POINTER   (-1)
#define RLG_ERR_BAD_SHAPE      (-2)   /* B < 0, N < 1, M < 1 (the reference raises IndexError for empty clouds) */
#define RLG_ERR_WORKSPACE      (-3)   /* workspace missing, misaligned (256 B) or too small */
#define RLG_ERR_UNSUPPORTED    (-4)   /* layer shape / option the kernel does not cover */
#define RLG_ERR_TOO_LARGE      (-5)   /* index would not fit the packed 32-bit fields */

/* flags for rlg_chamfer_fwd / rlg_chamfer_loss_fwd.  Any other bit is rejected with RLG_ERR_UNSUPPORTED (timing
 * variants of the kernels exist only in the separate experiments build, never in librlg_b200.so). */
#define RLG_CHAMFER_WS_CLEAN   1u     /* workspace is known to hold the all-ones pattern the previous
                                         rlg_chamfer_fwd left behind on the same shape: skip the memset */
#define RLG_CHAMFER_ALGO_SIMPLE 2u    /* one thread per query point, every candidate in the direct form: the plain
                                         restatement the production kernels are cross-checked against */
#define RLG_CHAMFER_TILE_ONLY  4u     /* measurement aid: enqueue only the first launch (the pair sweep; with
                                         RLG_CHAMFER_ALGO_TENSOR that includes the fused refinement) so it can be timed
                                         alone: means/loss untouched, workspace left dirty */
/* 8u is reserved (ABI <= 3: RLG_CHAMFER_ALGO_DIRECT, removed) */
#define RLG_CHAMFER_ALGO_TENSOR 16u   /* pair sweep with the 3-term contraction on the tensor cores (tcgen05, split-tf32
                                         operands, fp32 accumulation in TMEM), minima and the exact refinement on the
                                         CUDA cores of the same kernel, then a small tail kernel (chamfer_tcsweep.cu).
                                         Without it: pair sweep on the FP32 pipe + refinement kernel (chamfer_filter.cu).
                                         Same outputs bit for bit. */
#define RLG_CHAMFER_TRACK_TWO  32u    /* with ALGO_TENSOR: also track the runner-up's group and the third-smallest group
                                         minimum, so an ambiguous point is refined on two groups instead of the whole
                                         candidate cloud.  The library switches this on by itself beyond 4096 points;
                                         the flag forces it for smaller clouds (same outputs). */
#define RLG_CHAMFER_FILTER_ONLY 64u   /* diagnostic, with ALGO_TENSOR: run the filter sweep alone and publish, per query,
                                         (smallest group minimum << 32 | group) as u64 at ws[0 .. 8*B*(N+M)) (pc1's
                                         queries first) and the second-smallest group minimum as u32 right after
                                         (256-B aligned); d/i/means untouched, workspace left dirty.  tests/ measure the
                                         filter's rounding error against float64 with it. */

#define RLG_CHAMFER_RESERVE_SMS(n) (((unsigned)(n) & 0xffu) << 16)
                                      /* with ALGO_TENSOR: size the persistent grid for (SM count - n) SMs.  One process per GPU
                                         with NCCL kernels in flight (the per-step loss / gradient all-reduce on a side stream):
                                         a communication kernel holds an SM while it waits for its peers, and a persistent grid
                                         of one CTA per SM then has a CTA waiting for that SM -- measured +6 us (2 GPUs) to
                                         +10 us (4 GPUs) on a 42 us forward.  Same outputs. */

int rlg_version(void);
const char *rlg_last_error(void);

/* Number of SMs of the current device (used by callers to size batches); <0 on error. */
int rlg_device_sm_count(void);

/* ---------------------------------------------------------------------------------------------
 * Chamfer distance, forward  (utils/losses.py:29-37)
 *
 *   pc1 (B,N,3), pc2 (B,M,3)                      device, fp32
 *   d1 (B,N)  min_j |pc1[b,i]-pc2[b,j]|_2         device, fp32   (NOT squared; = torch.min(cdist,2)[0])
 *   i1 (B,N)  argmin_j, lowest index on ties      device, int32  (= torch.min(cdist,2)[1])
 *   d2 (B,M), i2 (B,M)                            the same from pc2 to pc1 (torch.min(cdist,1))
 *   mean1 (B), mean2 (B)                          device, fp32, nullable as a pair: torch.mean(d, dim=1)
 *   ws                                            device, >= rlg_chamfer_ws_bytes(B,N,M), 256-B aligned
 *
 * Distances are computed in the direct-difference form  t=d0*d0; t=fma(d1,d1,t); t=fma(d2,d2,t);
 * sqrtf(min t)  and are bit-identical to ATen's direct-mode cdist.  The argmin follows the reference's rule:
 * torch.min runs on the SQRT-ED matrix, so candidates whose squared distances share one sqrtf (up to three
 * adjacent fp32 values do) tie, and the lowest index among them wins.  The kernels find the candidates with a
 * cheap filter (|x|^2 + |y|^2 - 2x.y, on the tensor cores or the FP32 pipe) and re-evaluate every candidate within
 * a rigorous rounding margin in the direct form, so the outputs do not depend on the filter's rounding; clouds far
 * from the origin relative to their extent (|p|^2 >> min distance^2 * 1e5) only lose speed (more candidates are
 * re-evaluated), not exactness.
 * Non-finite coordinates are outside the contract: a NaN/Inf QUERY point yields a NaN/Inf distance as in the
 * reference, but a NaN CANDIDATE may be skipped where torch.min would let it win (the Python
 * seam can check inputs with install(check_finite=True)).
 * --------------------------------------------------------------------------------------------- */
size_t rlg_chamfer_ws_bytes(int B, int N, int M);

int rlg_chamfer_fwd(const float *pc1, const float *pc2, int B, int N, int M,
                    float *d1, float *d2, int32_t *i1, int32_t *i2,
                    float *mean1, float *mean2,
                    void *ws, size_t ws_bytes, unsigned flags, void *stream);

/* Forward fused with the reference's loss reduction (utils/losses.py:54-59 and :75):
 *   loss[0] = sum_b ( w1 * mean1[b] + w2 * mean2[b] ),  accumulated in a fixed order (deterministic);
 *   ChamferLoss(bidirectional=True) is w1 = w2 = 0.5/B, bidirectional=False is w1 = 1/B, w2 = 0.
 * loss: device fp32 scalar (nullable); needs mean1/mean2.
 * gz1 (B,N,3), gz2 (B,M,3): device fp32, each nullable: buffers the forward zero-fills on the way (for free, in its
 * last launch) so that the backward can run as ONE launch with RLG_CHAMFER_BWD_ACCUMULATE into them. */
int rlg_chamfer_loss_fwd(const float *pc1, const float *pc2, int B, int N, int M,
                         float *d1, float *d2, int32_t *i1, int32_t *i2,
                         float *mean1, float *mean2, float *loss, float w1, float w2,
                         float *gz1, float *gz2,
                         void *ws, size_t ws_bytes, unsigned flags, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Chamfer distance, backward
 *
 *   g1 (B), g2 (B)      device, fp32: upstream gradients of mean1, mean2 (nullable individually = 0)
 *   gpc1 (B,N,3), gpc2 (B,M,3)   device, fp32, fully overwritten (or added to, with RLG_CHAMFER_BWD_ACCUMULATE):
 *       gpc1[b,i] = g1[b]/N * (pc1[b,i]-pc2[b,i1])/d1[b,i]  -  sum_{j: i2[b,j]==i} g2[b]/M * (pc2[b,j]-pc1[b,i])/d2[b,j]
 *       (and symmetrically for gpc2); a term is 0 where its distance is 0 (EuclideanDistBackward0).
 *   flags   RLG_CHAMFER_BWD_ACCUMULATE: gpc1/gpc2 already hold zeros (see rlg_chamfer_loss_fwd's gz1/gz2) or a
 *           gradient to add to: one launch of fire-and-forget float reductions (2- and 4-float wide when
 *           gpc1/gpc2 are 8- / 16-byte aligned; any alignment works).  Without it the library zero-fills
 *           them first (a memset node + the same launch).  Any other flag bit: RLG_ERR_UNSUPPORTED.
 *   B <= 65535 per call (RLG_ERR_TOO_LARGE beyond, like the forward).
 *   The L2 float reduction flushes subnormal sums to zero (gradient components below 1.2e-38).
 * --------------------------------------------------------------------------------------------- */
#define RLG_CHAMFER_BWD_ACCUMULATE 1u
int rlg_chamfer_bwd(const float *pc1, const float *pc2,
                    const float *d1, const float *d2, const int32_t *i1, const int32_t *i2,
                    const float *g1, const float *g2, int B, int N, int M,
                    float *gpc1, float *gpc2, unsigned flags, void *stream);

/* Backward of rlg_chamfer_loss_fwd: gloss is the device fp32 scalar upstream of loss[0]; the per-pair
 * upstreams are gloss[0]*w1 (mean1) and gloss[0]*w2 (mean2). */
int rlg_chamfer_loss_bwd(const float *pc1, const float *pc2,
                         const float *d1, const float *d2, const int32_t *i1, const int32_t *i2,
                         const float *gloss, float w1, float w2, int B, int N, int M,
                         float *gpc1, float *gpc2, unsigned flags, void *stream);

/* Run-to-run reproducible variants of the two calls above (what torch.use_deterministic_algorithms(True) asks of an
 * operator; the reference's own CUDA autograd scatters with float atomics as well).  Same arguments and results to
 * within rounding, but the partner terms are summed in 64-bit FIXED POINT -- every term of one direction of one pair has
 * Euclidean norm |g[b]|/n exactly, so a quantum of 2^-40 of that (less for clouds of more than 2^21 points) loses nothing
 * an fp32 sum would keep -- with integer atomics, which commute: the result does not depend on the order of arrival.
 *   ws   device workspace of rlg_chamfer_bwd_ws_bytes(B, N, M) bytes (24 per point), 16-byte aligned; content on entry
 *        is irrelevant.  A memset node + two launches; gpc1/gpc2 need no zero-fill (every row is written once;
 *        with RLG_CHAMFER_BWD_ACCUMULATE the finished row is added to what the buffer holds).
 *   A non-finite upstream weight makes every gradient row of that pair's partner cloud NaN. */
size_t rlg_chamfer_bwd_ws_bytes(int B, int N, int M);
int rlg_chamfer_bwd_det(const float *pc1, const float *pc2,
                        const float *d1, const float *d2, const int32_t *i1, const int32_t *i2,
                        const float *g1, const float *g2, int B, int N, int M,
                        float *gpc1, float *gpc2, void *ws, size_t ws_bytes, unsigned flags, void *stream);
int rlg_chamfer_loss_bwd_det(const float *pc1, const float *pc2,
                             const float *d1, const float *d2, const int32_t *i1, const int32_t *i2,
                             const float *gloss, float w1, float w2, int B, int N, int M,
                             float *gpc1, float *gpc2, void *ws, size_t ws_bytes, unsigned flags, void *stream);

/* ---------------------------------------------------------------------------------------------
 * PointNet encoder: shared per-point MLP + global max-pool  (models/autoencoder.py:65-71), eval mode.
 *
 * The caller folds each Conv1d(k=1)+BatchNorm1d pair into one affine layer
 *       W' = W * gamma/sqrt(var+eps),  b' = (b-mean)*gamma/sqrt(var+eps) + beta
 * and passes the folded layers; every layer is followed by ReLU (autoencoder.py:33-45).
 *
 *   x (B,N,3)                      device fp32
 *   layers[l].w (c_out,c_in)       device fp32 row-major;  layers[l].b (c_out) device fp32
 *   pooled (B, c_out of last)      device fp32 = torch.max(point_mlp(x^T), dim=2)[0]
 *   argmax (B, c_out of last)      device int32, nullable: a point index attaining the max
 *
 * rlg_encoder_fwd      fp32 CUDA-core path, any layer widths (c_in of layer 0 must be 3).
 * rlg_encoder_fwd_bf16 tcgen05/TMEM path, see below.
 * --------------------------------------------------------------------------------------------- */
typedef struct rlg_layer {
    const float *w;     /* device, (c_out, c_in) row-major, BatchNorm already folded */
    const float *b;     /* device, (c_out) */
    int32_t c_in;
    int32_t c_out;
} rlg_layer;

size_t rlg_encoder_ws_bytes(int B, int N, const rlg_layer *layers, int L);

int rlg_encoder_fwd(const float *x, int B, int N, const rlg_layer *layers, int L,
                    float *pooled, int32_t *argmax,
                    void *ws, size_t ws_bytes, void *stream);

/* tcgen05 / TMEM path (bf16 operands, fp32 accumulation).  Layer 0 (3 -> c1) stays fp32 on the CUDA cores; layers
 * >= 1 are tensor-core GEMMs whose accumulators live in TMEM; the last layer's bias+ReLU+max-pool is the epilogue,
 * so the (B, C_last, N) activation never leaves the SM.  Supported widths: every hidden width 64 or 128, last width
 * a multiple of 128, L >= 2 (anything else: RLG_ERR_UNSUPPORTED -> use rlg_encoder_fwd).
 *   rlg_encoder_pack_bytes  size of the packed bf16 weight image (0 if the layer list is unsupported)
 *   rlg_encoder_pack_bf16   folded fp32 layers -> bf16 images in the swizzled shared-memory layout (device, 256-B
 *                           aligned); repack whenever the weights change
 *   rlg_encoder_fwd_bf16    pooled (B, C_last) fp32, fully overwritten; `layers` still supplies layer 0 and biases */
size_t rlg_encoder_pack_bytes(const rlg_layer *layers, int L);
int rlg_encoder_pack_bf16(const rlg_layer *layers, int L, void *packed, size_t packed_bytes, void *stream);
int rlg_encoder_fwd_bf16(const float *x, int B, int N, const rlg_layer *layers, int L,
                         const void *packed, size_t packed_bytes, float *pooled, void *stream);

/* Layer-by-layer tensor-core path (encoder_layers.cu): TMA-fed tcgen05 GEMMs with the layer's weights resident in shared
 * memory and the activations in HBM between layers.  Any layer list with widths that are multiples of 64, hidden widths
 * <= 256 (the reference's own 3->64->128->128->256->128 included; anything else: RLG_ERR_UNSUPPORTED -> rlg_encoder_fwd).
 *   mode RLG_ENC_BF16   bf16 operands, fp32 accumulation (2e-2 class)
 *        RLG_ENC_FP32X  fp32-grade: operands carried as fp16 hi + lo pairs (22 bits), three MMAs per K step
 *   weight_scales       host array of L floats (entry 0 unused): per-layer power-of-two scale applied to the weights
 *                       when packing and undone in the epilogue; choose the largest power of two with
 *                       max|w| * scale <= 2^14 (RLG_ENC_FP32X; ignored for RLG_ENC_BF16).  Same values for pack and fwd.
 *   rlg_encoder_gemm_pack_bytes / rlg_encoder_gemm_pack   packed operand image of the weights (repack when they change)
 *   rlg_encoder_gemm_ws_bytes                             two ping-pong activation buffers (caller-owned)
 *   rlg_encoder_gemm_fwd                                  pooled (B, C_last) fp32, fully overwritten */
#define RLG_ENC_BF16  1
#define RLG_ENC_FP32X 2
size_t rlg_encoder_gemm_pack_bytes(const rlg_layer *layers, int L, int mode);
int rlg_encoder_gemm_pack(const rlg_layer *layers, int L, int mode, const float *weight_scales,
                          void *packed, size_t packed_bytes, void *stream);
size_t rlg_encoder_gemm_ws_bytes(int B, int N, const rlg_layer *layers, int L, int mode);
int rlg_encoder_gemm_fwd(const float *x, int B, int N, const rlg_layer *layers, int L, int mode,
                         const float *weight_scales, const void *packed, size_t packed_bytes,
                         float *pooled, void *ws, size_t ws_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * PointNet encoder trunk WITH its BatchNorm layers: forward and backward  (models/autoencoder.py:32-47,65-71 as
 * driven by the autoencoder training step, train_rl_gan_net.py:220-249: model.train(), loss.backward()).
 *
 * Each block is Conv1d(k=1) -> BatchNorm1d -> ReLU; the last block is followed by the max over the points.
 *   flags & RLG_ENC_BATCH_STATS   train mode: BatchNorm normalises with the statistics of this batch (biased variance
 *                                 over the B*N points) and updates running_mean / running_var in place
 *                                 (running = (1-momentum)*running + momentum*batch, unbiased variance), as
 *                                 torch.nn.BatchNorm1d does.  The caller increments num_batches_tracked.
 *   without it                    eval mode: BatchNorm uses running_mean / running_var (read only); the backward is that
 *                                 of the affine map (this is the encoder backward of SURVEY 8(b), rlg_encoder_bwd).
 * GEMMs (layers >= 1, forward, input-gradient and weight-gradient) run on the tensor cores at fp32 grade (tcgen05,
 * fp16 hi+lo operand pairs, fp32 accumulation in TMEM); statistics are accumulated in float64.
 * Widths: c_in of layer 0 is 3, every c_out a multiple of 64 and <= 256, 2 <= L <= 8 (else RLG_ERR_UNSUPPORTED: the
 * caller keeps the stock layers).  B*N >= 2 in train mode.
 *
 *   rlg_encoder_train_fwd   pooled (B, C_last) fp32 = max over points of the trunk's output; `saved` (caller-owned,
 *                           >= rlg_encoder_train_saved_bytes) receives what the backward needs (pre-BatchNorm outputs,
 *                           operand pieces of the activations, batch statistics, arg-max keys); `ws` is scratch.
 *   rlg_encoder_train_bwd   g_pooled (B, C_last) fp32 upstream gradient; grads[l] receives dL/d(conv weight) (c_out,c_in),
 *                           dL/d(conv bias), dL/d(gamma), dL/d(beta) (each nullable).  No gradient for x (the clouds
 *                           are data).  Must see the same layers (weights unchanged) and the `saved` of its forward.
 * --------------------------------------------------------------------------------------------- */
typedef struct rlg_bn_layer {
    const float *w;          /* device, (c_out, c_in) row-major: Conv1d weight with the kernel axis squeezed */
    const float *b;          /* device, (c_out) Conv1d bias, nullable */
    const float *gamma;      /* device, (c_out) BatchNorm weight, nullable (= 1) */
    const float *beta;       /* device, (c_out) BatchNorm bias, nullable (= 0) */
    float *running_mean;     /* device, (c_out); updated in train mode, read in eval mode */
    float *running_var;      /* device, (c_out) */
    float eps;
    float momentum;
    int32_t c_in;
    int32_t c_out;
} rlg_bn_layer;

typedef struct rlg_bn_grads {
    float *dw;               /* device, (c_out, c_in), nullable */
    float *db;               /* device, (c_out), nullable (exactly 0 in train mode: BatchNorm removes the bias) */
    float *dgamma;           /* device, (c_out), nullable */
    float *dbeta;            /* device, (c_out), nullable */
} rlg_bn_grads;

#define RLG_ENC_BATCH_STATS 1u
#define RLG_ENC_RESERVE_SMS(n) (((unsigned)(n) & 0xffu) << 16)   /* size the persistent GEMM grids for (SM count - n) SMs: leaves
                                                                     room for communication kernels running beside them */
size_t rlg_encoder_train_saved_bytes(int B, int N, const rlg_bn_layer *layers, int L);
size_t rlg_encoder_train_ws_bytes(int B, int N, const rlg_bn_layer *layers, int L);
/* Where things live inside `saved` (inspection / tests): offsets[5*l + k] for layer l with k = 0: z, fp32 (B*N, c_out)
 * pre-BatchNorm outputs; 1, 2: hi and lo fp16 pieces (B*N, c_out) of the post-ReLU activations (absent = 0 for the
 * last layer); 3: fp32 (4, c_out) = gamma*invstd, beta - mean*gamma*invstd, mean*invstd, invstd; 4: u32 (c_out) bit
 * pattern of max |zhat|;  offsets[5*L]: u64 (B, C_last) max-pool keys = (value bits << 32) | (0xFFFFFFFF - point). */
int rlg_encoder_train_saved_layout(int B, int N, const rlg_bn_layer *layers, int L, size_t *offsets, int n_offsets);
int rlg_encoder_train_fwd(const float *x, int B, int N, const rlg_bn_layer *layers, int L, unsigned flags,
                          float *pooled, void *saved, size_t saved_bytes, void *ws, size_t ws_bytes, void *stream);
int rlg_encoder_train_bwd(const float *x, int B, int N, const rlg_bn_layer *layers, int L, unsigned flags,
                          const float *g_pooled, const void *saved, size_t saved_bytes,
                          const rlg_bn_grads *grads, void *ws, size_t ws_bytes, void *stream);

/* ---------------------------------------------------------------------------------------------
 * Input pipeline on the device: the per-sample numeric work of the reference's DataLoader for a batch of B
 * items of a binary cache of complete clouds (items, N, 3) fp32 that lives in HBM:
 *   utils/dataset.py:252-276  _create_incomplete_pc   (random subset | points outside a sphere, np.percentile radius in float64)
 *   utils/dataset.py:278-297  _augment_point_cloud    (pc @ R^T, + clipped jitter, * scale)
 *   utils/data_utils.py:15-60 normalize_point_cloud   (centroid, largest norm)
 *   utils/dataset.py:393-421  shapenet_collate_fn     (pad the incomplete clouds to the batch's longest by repeating points)
 * Every random decision is an input (the host draws a few numbers per cloud); all arrays below are DEVICE pointers.
 *   complete_out   (B, N, 3)  augmented + normalised complete clouds
 *   incomplete_out (B, N, 3)  capacity; cloud b holds lengths[b] points followed by padding up to *max_len
 *   lengths (B), max_len (1)  int32; the caller slices incomplete_out[:, :max_len]
 * N <= 4096.
 * --------------------------------------------------------------------------------------------- */
typedef struct rlg_prepare_plan {
    const int32_t *item;      /* (B) row of the cache; NULL = rows 0..B-1 */
    const int32_t *method;    /* (B) 0 = random subset in drawn order (keep_idx), 1 = spatial removal */
    const int32_t *n_keep;    /* (B) method 0: number of kept points */
    const int32_t *keep_idx;  /* (B, N) method 0: kept point indices, first n_keep valid */
    const int32_t *center;    /* (B) method 1: index of the sphere's centre point */
    const int32_t *q_index;   /* (B) method 1: floor((N-1) * ratio), numpy's percentile index */
    const double  *q_gamma;   /* (B) method 1: its fractional part */
    const float   *rot;       /* (2, B, 9) row-major rotation matrices for the complete / incomplete cloud; NULL = none */
    const float   *scale;     /* (2, B) scale factors; NULL = 1 */
    const float   *jitter;    /* (2, B, N, 3) additive noise, already clipped; NULL = none */
    const int32_t *pad_idx;   /* (B, N) non-negative random integers: pad slot s repeats point pad_idx[b][s] % lengths[b] */
} rlg_prepare_plan;

int rlg_batch_prepare(const float *cache, int n_items, int N, int B, const rlg_prepare_plan *plan,
                      float *complete_out, float *incomplete_out, int32_t *lengths, int32_t *max_len, void *stream);

/* FP32 CUDA-core peak microbenchmark (measurement helper, not on the hot path; it synchronises).
 * Fills host array out[0..5]:
 *   [0] measured scalar FFMA TFLOP/s     [1] max SM clock (MHz)      [2] theoretical SMs*128*2*clock TFLOP/s
 *   [3] SM count                         [4] measured packed FFMA2 TFLOP/s
 *   [5] the Chamfer inner-loop instruction mix without memory traffic, as 8 flop per pair (TFLOP/s)
 * scratch_dev: any device buffer of >= 4 bytes. */
int rlg_fp32_peak(float *out_host, int n_out, void *scratch_dev, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RLG_B200_H */
