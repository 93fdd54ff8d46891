// chamfer_finalize.cu -- second (small) launch of the Chamfer forward.
//
// The tile kernel leaves, per query point, the packed key (min squared distance bits << 32 | winning group).
// A warp takes 32 points; its lanes each re-evaluate ONE candidate of a point's winning group (coalesced reads),
// a ballot picks the LOWEST index attaining the minimum, and the kernel writes sqrtf(min t) and the index and
// restores the key to the all-ones state the next forward expects.  Per-pair means (utils/losses.py:36-37)
// and the batch loss (:54-59, :75) are reduced in a fixed order so results are bit-reproducible run to run:
// CTA partial sums (double) -> the last CTA of each (cloud, direction) adds them in chunk order -> the last of
// those folds the means into the loss.  "Last" is decided by counters that live in the workspace, start at
// 0xffffffff (its all-ones state) and are put back to it.
#include "common.cuh"

namespace rlg {

static constexpr int kFinThreads = 256;

struct FinWs {
    unsigned *global_counter;   // 1
    unsigned *pair_counter;     // 2*B
    double *partial;            // 2*B*chunks_max
    int chunks_max;
};

static size_t fin_counter_bytes(int B) { return align_up(sizeof(unsigned) * (1 + 2 * (size_t)B), 256); }
static int fin_chunks(int n) { return (n + kFinThreads - 1) / kFinThreads; }

size_t finalize_ws_bytes(int B, int N, int M) {
    const int cm = fin_chunks(N > M ? N : M);
    return fin_counter_bytes(B) + align_up(sizeof(double) * 2 * (size_t)B * cm, 256);
}

__global__ void __launch_bounds__(kFinThreads) chamfer_finalize_kernel(
    const float *__restrict__ pc1, const float *__restrict__ pc2, int N, int M, int rows_per_group,
    u64 *__restrict__ rowkey, u64 *__restrict__ colkey, FinWs fw, float *__restrict__ d1, float *__restrict__ d2,
    int32_t *__restrict__ i1, int32_t *__restrict__ i2, float *__restrict__ mean1, float *__restrict__ mean2,
    float *__restrict__ loss, float loss_w1, float loss_w2) {
    __shared__ double red[kFinThreads / 32];
    __shared__ bool s_last;
    const int chunk = blockIdx.x, b = blockIdx.y, dir = blockIdx.z;
    const int B = gridDim.y;
    // dir 0: queries = pc1 rows, candidates = pc2 columns in groups of 32
    // dir 1: queries = pc2 columns, candidates = pc1 rows in groups of rows_per_group
    const int nq = dir ? M : N, nc = dir ? N : M;
    const int n_chunks = (nq + kFinThreads - 1) / kFinThreads;
    if (chunk >= n_chunks) return;                       // grid.x is sized for the longer direction
    const int gsz = dir ? rows_per_group : kGroup;
    const float *q = (dir ? pc2 : pc1) + (size_t)b * nq * 3;
    const float *c = (dir ? pc1 : pc2) + (size_t)b * nc * 3;
    u64 *keys = (dir ? colkey : rowkey) + (size_t)b * nq;

    // Each warp owns 32 consecutive query points.  Lane l loads point l's key and coordinates (coalesced);
    // then, 32/gsz points at a time, the warp's lanes each test ONE candidate of the winning group (coalesced
    // 12-byte reads) and a ballot picks the lowest matching index.
    const int lane = threadIdx.x & 31;
    const int i_mine = chunk * kFinThreads + threadIdx.x;
    u64 key_mine = kKeyInit;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (i_mine < nq) {
        key_mine = keys[i_mine];
        keys[i_mine] = kKeyInit;
        qx = q[3 * i_mine]; qy = q[3 * i_mine + 1]; qz = q[3 * i_mine + 2];
    }
    const int ppi = 32 / gsz;                 // points per iteration (gsz divides 32)
    const int sub = lane / gsz, cand = lane - sub * gsz;
    const unsigned sub_mask = (gsz == 32 ? 0xffffffffu : ((1u << gsz) - 1u)) << (sub * gsz);
    const int warp_base = chunk * kFinThreads + (threadIdx.x & ~31);
    double dist_d = 0.0;
#pragma unroll 4
    for (int it = 0; it < gsz; ++it) {        // 32 / ppi == gsz iterations
        const int src = it * ppi + sub;       // lane holding this point's key/coords
        const u64 key = __shfl_sync(0xffffffffu, key_mine, src);
        const float px = __shfl_sync(0xffffffffu, qx, src);
        const float py = __shfl_sync(0xffffffffu, qy, src);
        const float pz = __shfl_sync(0xffffffffu, qz, src);
        const int i = warp_base + src;
        float tmin = __uint_as_float((unsigned)(key >> 32));
        const int j = (int)(unsigned)(key & 0xffffffffu) * gsz + cand;
        const bool live = (i < nq) && (key != kKeyInit) && (j < nc);
        float t = 0.f;
        if (live) t = sqdist(px, py, pz, __ldg(c + 3 * j), __ldg(c + 3 * j + 1), __ldg(c + 3 * j + 2));
        const unsigned hits = __ballot_sync(0xffffffffu, live && t == tmin) & sub_mask;
        if (cand == 0 && i < nq) {
            int found;
            if (hits) {
                found = j + (__ffs((int)hits) - 1 - sub * gsz);
            } else {
                // only reachable with non-finite input (outside the contract): mirror torch.min, where the first
                // NaN wins -> candidate 0
                found = 0;
                tmin = sqdist(px, py, pz, c[0], c[1], c[2]);
            }
            const float dist = sqrtf(tmin);
            (dir ? d2 : d1)[(size_t)b * nq + i] = dist;
            (dir ? i2 : i1)[(size_t)b * nq + i] = found;
            dist_d += (double)dist;
        }
    }
    if (mean1 == nullptr) return;

    // CTA partial sum in a fixed order: lane tree, then warps 0..7 sequentially
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) dist_d += __shfl_down_sync(0xffffffffu, dist_d, s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dist_d;
    __syncthreads();
    if (threadIdx.x != 0) return;
    double part = 0.0;
#pragma unroll
    for (int w = 0; w < kFinThreads / 32; ++w) part += red[w];
    const int pair_slot = b * 2 + dir;
    volatile double *parts = fw.partial + (size_t)pair_slot * fw.chunks_max;
    parts[chunk] = part;
    __threadfence();
    if (atomicAdd(fw.pair_counter + pair_slot, 1u) != (unsigned)(n_chunks - 2)) return;
    // last CTA of this (cloud, direction)
    __threadfence();
    fw.pair_counter[pair_slot] = 0xffffffffu;
    double total = 0.0;
    for (int k = 0; k < n_chunks; ++k) total += parts[k];
    (dir ? mean2 : mean1)[b] = (float)(total / (double)nq);
    if (loss == nullptr) return;
    __threadfence();
    if (atomicAdd(fw.global_counter, 1u) != (unsigned)(2 * B - 2)) return;
    __threadfence();
    *fw.global_counter = 0xffffffffu;
    const volatile float *m1 = mean1, *m2 = mean2;
    double acc = 0.0;
    for (int k = 0; k < B; ++k) acc += (double)loss_w1 * (double)m1[k] + (double)loss_w2 * (double)m2[k];
    *loss = (float)acc;
}

int launch_finalize(const float *pc1, const float *pc2, int B, int N, int M, int rows_per_group, u64 *rowkey,
                    u64 *colkey, void *fin_ws, float *d1, float *d2, int32_t *i1, int32_t *i2, float *mean1,
                    float *mean2, float *loss, float w1, float w2, cudaStream_t st) {
    FinWs fw;
    fw.global_counter = (unsigned *)fin_ws;
    fw.pair_counter = fw.global_counter + 1;
    fw.partial = (double *)((char *)fin_ws + fin_counter_bytes(B));
    fw.chunks_max = fin_chunks(N > M ? N : M);
    dim3 grid(fw.chunks_max, B, 2);
    chamfer_finalize_kernel<<<grid, kFinThreads, 0, st>>>(pc1, pc2, N, M, rows_per_group, rowkey, colkey, fw, d1, d2,
                                                         i1, i2, mean1, mean2, loss, w1, w2);
    return check_launch("chamfer_finalize_kernel");
}

}  // namespace rlg
