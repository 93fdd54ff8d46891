// tools/umma_bench.cu -- how fast does ONE CTA per SM issue tcgen05.mma (kind::f16, M=128, K=16 per instruction)?
// Back-to-back MMAs from one elected thread, operands static (contents irrelevant), no epilogue.  Prints cycles per MMA
// for: both operands in shared memory (SS) or A in tensor memory (TS); N = 64 / 128 / 256; K-chain length 8 (one
// accumulator per 8 MMAs, round-robin over `nacc` accumulators).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/umma_bench tools/umma_bench.cu && build/umma_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t a) {
    return (uint64_t)((a & 0x3FFFFu) >> 4) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t umma_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <bool TS, int DELAY = 0>
__global__ void __launch_bounds__(128, 1) k_umma(int N, int nacc, int groups, long long *cycles_out) {
    extern __shared__ unsigned char raw[];
    const uint32_t pad = (1024u - (smem_u32(raw) & 1023u)) & 1023u;
    unsigned char *smem = raw + pad;
    const uint32_t sA = smem_u32(smem), sB = sA + 32768;       // A: 128 x 128 bf16 (32 KB), B: 256 x 128 bf16 (64 KB)
    __shared__ uint64_t bar, bar2, bar3;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < (32768 + 65536) / 4; e += 128) reinterpret_cast<uint32_t *>(smem)[e] = 0;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar2)), "r"(1));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar3)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (tid == 0) {
        const uint32_t idesc = umma_idesc(128, N);
        // accumulators first (nacc x N columns), then (TS) the A operand: 64 columns
        const uint32_t a_tmem = tmem + (uint32_t)(nacc * N);
        const long long t0 = clock64();
        for (int g = 0; g < groups; ++g) {
            const uint32_t d = tmem + (uint32_t)((g % nacc) * N);
            uint64_t ad = umma_desc(sA), bd = umma_desc(sB);
            uint32_t acc = 0;
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
                for (int k4 = 0; k4 < 4; ++k4) {
                    if (TS) mma_ts(d, a_tmem + (uint32_t)(kb * 32 + k4 * 8), bd + (uint64_t)(k4 * 2), idesc, acc);
                    else mma_ss(d, ad + (uint64_t)(k4 * 2), bd + (uint64_t)(k4 * 2), idesc, acc);
                    acc = 1;
                }
                ad += (128 * 128) >> 4;
                bd += (uint64_t)((N * 128) >> 4);
            }
            if (DELAY > 0) { const long long w0 = clock64(); while (clock64() - w0 < DELAY) {} }
            if (DELAY < 0) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2)) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        uint32_t ok = 0;
        while (!ok)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        const long long t1 = clock64();
        if (blockIdx.x == 0) *cycles_out = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}


// Same issue loop, but every group of 8 MMAs (one 128x128 accumulator, K=128) is handed to 4 epilogue warps through
// mbarriers (commit -> full; tcgen05.ld x64 of half the columns x2 -> arrive empty), nacc accumulators in flight.
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
template <int LDMODE>
__global__ void __launch_bounds__(160, 1) k_handshake(int nacc, int groups, long long *cycles_out) {
    extern __shared__ unsigned char raw[];
    const uint32_t pad = (1024u - (smem_u32(raw) & 1023u)) & 1023u;
    unsigned char *smem = raw + pad;
    const uint32_t sA = smem_u32(smem), sB = sA + 32768;
    __shared__ uint64_t full[4], empty[4];
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int e = tid; e < (32768 + 65536) / 4; e += 160) reinterpret_cast<uint32_t *>(smem)[e] = 0;
    if (tid == 0) {
        for (int k = 0; k < 4; ++k) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[k])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[k])), "r"(128));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (warp == 4) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(128, 128);
            const long long t0 = clock64();
            for (int g = 0; g < groups; ++g) {
                const int a = g % nacc, use = g / nacc;
                bar_wait(smem_u32(&empty[a]), (uint32_t)((use & 1) ^ 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem + (uint32_t)(a * 128);
                uint64_t ad = umma_desc(sA), bd = umma_desc(sB);
                uint32_t acc = 0;
#pragma unroll
                for (int kb = 0; kb < 2; ++kb) {
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) { mma_ss(d, ad + (uint64_t)(k4 * 2), bd + (uint64_t)(k4 * 2), idesc, acc); acc = 1; }
                    ad += (128 * 128) >> 4;
                    bd += (128 * 128) >> 4;
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&full[a])) : "memory");
            }
            // drain: wait until the epilogue released the last accumulator
            const int g = groups - 1;
            bar_wait(smem_u32(&empty[g % nacc]), (uint32_t)((g / nacc) & 1));
            const long long t1 = clock64();
            if (blockIdx.x == 0) *cycles_out = t1 - t0;
        }
    } else {
        const uint32_t lane_base = ((uint32_t)warp * 32u) << 16;
        float keep = 0.f;
        for (int g = 0; g < groups; ++g) {
            const int a = g % nacc, use = g / nacc;
            bar_wait(smem_u32(&full[a]), (uint32_t)(use & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (LDMODE == 1) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t r[64];
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 "
                                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
                                 "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];\n\t"
                                 "tcgen05.wait::ld.sync.aligned;"
                                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
                                   "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                                   "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
                                   "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]),
                                   "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]),
                                   "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
                                 : "r"(tmem + lane_base + (uint32_t)(a * 128 + h * 64)) : "memory");
                    keep += __uint_as_float(r[0]) + __uint_as_float(r[63]);
                }
            } else if (LDMODE == 2) {
#pragma unroll
                for (int h = 0; h < 8; ++h) {
                    uint32_t r[16];
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                                 : "r"(tmem + lane_base + (uint32_t)(a * 128 + h * 16)) : "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    keep += __uint_as_float(r[0]) + __uint_as_float(r[15]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[a])) : "memory");
        }
        if (keep == 123.456f) cycles_out[1] = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}


// TMEM -> register read bandwidth of tcgen05.ld shapes: 4 warps, each re-reading its own 32-lane quarter.
template <int SHAPE>
__global__ void __launch_bounds__(128, 1) k_tmem_read(int iters, long long *cycles_out) {
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot + (((uint32_t)warp * 32u) << 16);
    uint32_t keep = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < 512; c += 128) {           // 128 columns x 32 lanes per step = 16 KB per warp
            uint32_t r[64];
            if (SHAPE == 0) {                          // 32x32b.x64 twice
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 "
                                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
                                 "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
                                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
                                   "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                                   "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
                                   "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]),
                                   "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]),
                                   "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
                                 : "r"(tmem + (uint32_t)(c + h * 64)) : "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    keep ^= r[0] ^ r[63];
                }
            } else {                                   // 16x256b.x16: 16 lanes x 128 columns, twice (lanes 0-15, 16-31)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    asm volatile("tcgen05.ld.sync.aligned.16x256b.x16.b32 "
                                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
                                 "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
                                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
                                   "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                                   "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
                                   "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]),
                                   "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]),
                                   "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
                                 : "r"(tmem + (uint32_t)c + (((uint32_t)h * 16u) << 16)) : "memory");
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    keep ^= r[0] ^ r[63];
                }
            }
        }
    }
    const long long t1 = clock64();
    if (blockIdx.x == 0 && tid == 0) *cycles_out = t1 - t0;
    if (keep == 0x12345u) cycles_out[1] = 1;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512) : "memory");
}


// Handshake with warp-uniform issue: NI issuer warps (elect.sync), 4 epilogue warps (tcgen05.ld x64 x2 per block), nacc
// accumulators; block g goes to accumulator g % nacc and is issued by issuer (g % nacc) % NI.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P1;\n\telect.sync _|P1, 0xffffffff;\n\tselp.u32 %0, 1, 0, P1;\n\t}" : "=r"(pred));
    return pred != 0;
}
template <int NI, bool LOADS>
__global__ void __launch_bounds__(128 + 32 * NI, 1) k_pipe(int nacc, int groups, long long *cycles_out) {
    extern __shared__ unsigned char raw[];
    const uint32_t pad = (1024u - (smem_u32(raw) & 1023u)) & 1023u;
    unsigned char *smem = raw + pad;
    const uint32_t sA = smem_u32(smem), sB = sA + 32768;
    __shared__ uint64_t full[4], empty[4];
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < (32768 + 65536) / 4; e += 128 + 32 * NI) reinterpret_cast<uint32_t *>(smem)[e] = 0;
    if (tid == 0) {
        for (int k = 0; k < 4; ++k) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[k])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&empty[k])), "r"(128));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = slot;
    if (warp >= 4) {
        const int me = warp - 4;
        const bool leader = elect_one();
        const uint32_t idesc = umma_idesc(128, 128);
        const long long t0 = clock64();
        for (int g = 0; g < groups; ++g) {
            const int a = g % nacc, use = g / nacc;
            if (a % NI != me) continue;
            bar_wait(smem_u32(&empty[a]), (uint32_t)((use & 1) ^ 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t d = tmem + (uint32_t)(a * 128);
            uint64_t ad = umma_desc(sA), bd = umma_desc(sB);
            uint32_t acc = 0;
            for (int kb = 0; kb < 2; ++kb) {
                if (leader) {
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) { mma_ss(d, ad + (uint64_t)(k4 * 2), bd + (uint64_t)(k4 * 2), idesc, acc); acc = 1; }
                }
                acc = 1;
                ad += (128 * 128) >> 4;
                bd += (128 * 128) >> 4;
            }
            if (leader) asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&full[a])) : "memory");
            __syncwarp();
        }
        if (me == 0) {
            for (int a = 0; a < nacc; ++a) {          // drain: every accumulator's last use released
                int last = -1;
                for (int g = groups - 1; g >= 0; --g) if (g % nacc == a) { last = g; break; }
                if (last >= 0) bar_wait(smem_u32(&empty[a]), (uint32_t)((last / nacc) & 1));
            }
            const long long t1 = clock64();
            if (blockIdx.x == 0 && leader) *cycles_out = t1 - t0;
        }
    } else {
        const uint32_t lane_base = ((uint32_t)warp * 32u) << 16;
        float keep = 0.f;
        for (int g = 0; g < groups; ++g) {
            const int a = g % nacc, use = g / nacc;
            bar_wait(smem_u32(&full[a]), (uint32_t)(use & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (LOADS) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t r[64];
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 "
                                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
                                 "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];\n\t"
                                 "tcgen05.wait::ld.sync.aligned;"
                                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
                                   "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
                                   "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
                                   "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]),
                                   "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]),
                                   "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
                                 : "r"(tmem + lane_base + (uint32_t)(a * 128 + h * 64)) : "memory");
                    keep += __uint_as_float(r[0]) + __uint_as_float(r[63]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[a])) : "memory");
        }
        if (keep == 123.456f) cycles_out[1] = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long *d_cyc;
    cudaMalloc(&d_cyc, 16);
    const size_t smem = 32768 + 65536 + 1024;
    cudaFuncSetAttribute(k_umma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_umma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int groups = 4000;
    for (int grid : {1, sms}) {
        for (int ts = 0; ts < 2; ++ts) {
            for (int N : {64, 128, 256}) {
                for (int nacc : {1, 2}) {
                    if (nacc * N + (ts ? 64 : 0) > 512) continue;
                    cudaEvent_t e0, e1;
                    cudaEventCreate(&e0); cudaEventCreate(&e1);
                    for (int rep = 0; rep < 2; ++rep) {
                        cudaEventRecord(e0);
                        if (ts) k_umma<true><<<grid, 128, smem>>>(N, nacc, groups, d_cyc);
                        else k_umma<false><<<grid, 128, smem>>>(N, nacc, groups, d_cyc);
                        cudaEventRecord(e1);
                        cudaEventSynchronize(e1);
                    }
                    float ms; cudaEventElapsedTime(&ms, e0, e1);
                    long long cyc = 0; cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
                    cudaError_t e = cudaGetLastError();
                    const double per = (double)cyc / (groups * 8.0);
                    const double tflops = 2.0 * 128 * N * 16 * groups * 8.0 * grid / (ms * 1e-3) / 1e12;
                    printf("grid %3d  %s  N=%3d  nacc=%d : %7.1f cycles/MMA (floor %3d)  %8.1f TFLOP/s  %s\n", grid, ts ? "TS" : "SS", N,
                           nacc, per, 128 * N / 256, tflops, e == cudaSuccess ? "" : cudaGetErrorString(e));
                }
            }
        }
    }
    {
        auto run = [&](auto kern, const char *name) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            for (int rep = 0; rep < 2; ++rep) { kern<<<sms, 128, smem>>>(128, 2, groups, d_cyc); cudaDeviceSynchronize(); }
            long long cyc = 0; cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
            printf("SS N=128 nacc=2, issuer pause between groups of 8 MMAs %-12s: %7.1f cycles per group (512 = MMA bound)\n", name, (double)cyc / groups);
        };
        run(k_umma<false, 0>, "none");
        run(k_umma<false, -1>, "commit only");
        run(k_umma<false, 64>, "64 cyc");
        run(k_umma<false, 128>, "128 cyc");
        run(k_umma<false, 192>, "192 cyc");
        run(k_umma<false, 256>, "256 cyc");
        run(k_umma<false, 384>, "384 cyc");
        run(k_umma<false, 512>, "512 cyc");
    }
    {
        auto runp = [&](auto kern, int threads, int nacc, const char *name) {
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            for (int rep = 0; rep < 2; ++rep) { kern<<<sms, threads, smem>>>(nacc, 4000, d_cyc); cudaDeviceSynchronize(); }
            long long cyc = 0; cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
            cudaError_t e = cudaGetLastError();
            printf("pipe %-28s nacc=%d : %7.1f cycles per 128x128x128 block (512 = MMA bound) %s\n", name, nacc, (double)cyc / 4000.0,
                   e == cudaSuccess ? "" : cudaGetErrorString(e));
        };
        for (int nacc : {2, 3, 4}) {
            runp(k_pipe<1, false>, 160, nacc, "1 issuer, no TMEM loads");
            runp(k_pipe<1, true>, 160, nacc, "1 issuer, ld x64 x2");
            runp(k_pipe<2, false>, 192, nacc, "2 issuers, no TMEM loads");
            runp(k_pipe<2, true>, 192, nacc, "2 issuers, ld x64 x2");
        }
    }
    for (int shape = 0; shape < 2; ++shape) {
        const int iters = 2000;
        for (int rep = 0; rep < 2; ++rep) {
            if (shape == 0) k_tmem_read<0><<<sms, 128>>>(iters, d_cyc); else k_tmem_read<1><<<sms, 128>>>(iters, d_cyc);
            cudaDeviceSynchronize();
        }
        long long cyc = 0; cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
        cudaError_t e = cudaGetLastError();
        // per iteration the 4 warps read the whole TMEM: 128 lanes x 512 columns x 4 B = 256 KB
        printf("tmem read %s : %.1f bytes/clk/SM  %s\n", shape == 0 ? "32x32b.x64 " : "16x256b.x16", 262144.0 * iters / (double)cyc,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    cudaFuncSetAttribute(k_handshake<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_handshake<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_handshake<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int mode = 0; mode < 3; ++mode) {
        for (int nacc : {1, 2, 3}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k_handshake<0><<<sms, 160, smem>>>(nacc, 2000, d_cyc);
                else if (mode == 1) k_handshake<1><<<sms, 160, smem>>>(nacc, 2000, d_cyc);
                else k_handshake<2><<<sms, 160, smem>>>(nacc, 2000, d_cyc);
                cudaDeviceSynchronize();
            }
            long long cyc = 0; cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
            cudaError_t e = cudaGetLastError();
            printf("handshake %s nacc=%d : %7.1f cycles per 128x128x128 block (MMA floor 512)  %s\n",
                   mode == 0 ? "no loads   " : (mode == 1 ? "ld x64 x2  " : "ld x16 x8  "), nacc, (double)cyc / 2000.0,
                   e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
    return 0;
}
