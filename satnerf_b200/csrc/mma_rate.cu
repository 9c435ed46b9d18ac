// Developer microbenchmark: cycles per tcgen05.mma.kind::f16 instruction issued back to back (operands: whatever is in
// shared memory / TMEM; only timing matters).  mode 0: SS cg1 (A,B from smem)   mode 1: TS cg1 (A from TMEM)
//                                               mode 2: SS cg2 (M=256, leader issues)   mode 3: SS cg1, MN-major operands
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "satnerf_b200_dev.h"

namespace snb {
using namespace ptx;

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}

template <int CG>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int mode_in, int N, int iters, long long* out) {
    const int mode = mode_in & 15; const bool random_data = (mode_in & 16) != 0;
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint64_t bar; __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        // iters < 0: pseudo-random fp16 in +-[0.25, 1) (high toggle rate); otherwise constant 1.0
        reinterpret_cast<uint32_t*>(sm)[i] = random_data ? ((h & 0x83ff83ffu) | 0x34003400u) : 0x3c003c00u;
    }
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    if (warp == 0) { if constexpr (CG == 2) tmem_alloc_2cta(&tptr, 512); else tmem_alloc(&tptr, 512); }
    tc_fence_before(); __syncthreads(); if (CG == 2) cluster_sync_all(); tc_fence_after();
    const uint32_t tmem = tptr;
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 1 && rank == 0) {
        const uint32_t a = smem_u32(sm), b = a + 16384;
        long long t0 = 0, t1 = 0;
        if (lane == 0) {
            const uint32_t idesc = mode == 2 ? umma_idesc_f16_m256((uint32_t)N) : (mode == 3 ? umma_idesc_f16_mn((uint32_t)N) : umma_idesc_f16((uint32_t)N));
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                const uint32_t k = (uint32_t)(i & 3) * 32u;
                if (mode == 0) umma_f16_ss(tmem, umma_desc_k_sw128(a + k), umma_desc_k_sw128(b + k), idesc, 1);
                else if (mode == 1) umma_f16_ts(tmem + 256, tmem, umma_desc_k_sw128(b + k), idesc, 1);
                else if (mode == 2) { if constexpr (CG == 2) umma_f16_ss_2cta(tmem, umma_desc_k_sw128(a + k), umma_desc_k_sw128(b + k), idesc, 1); }
                else umma_f16_ss(tmem, umma_desc_mn_sw128(a + (i & 3) * 2048, 8192, 1024), umma_desc_mn_sw128(b + (i & 3) * 2048, 8192, 1024), idesc, 1);
            }
            if constexpr (CG == 2) umma_commit_2cta(&bar, 1); else umma_commit(&bar);
        }
        __syncwarp();
        mbar_wait(&bar, 0, 31);
        if (lane == 0) { t1 = clock64(); if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = iters; } }
    }
    tc_fence_before(); __syncthreads(); if (CG == 2) cluster_sync_all();
    if (warp == 0) { if constexpr (CG == 2) tmem_dealloc_2cta(tmem, 512); else tmem_dealloc(tmem, 512); }
}
}  // namespace snb

using namespace snb;
extern "C" SNB_API int snb_debug_mma_rate(int mode, int N, int iters, int n_blocks, long long* host_out) {
    long long* d = nullptr;
    SNB_CUDA(cudaMalloc(&d, 16)); SNB_CUDA(cudaMemset(d, 0, 16));
    const size_t smem = 64 * 1024;
    if ((mode & 15) == 2) {
        SNB_CUDA(cudaFuncSetAttribute(mma_rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(n_blocks < 2 ? 2 : (n_blocks & ~1)); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        SNB_CUDA(cudaLaunchKernelEx(&cfg, mma_rate_kernel<2>, mode, N, iters, d));
    } else {
        SNB_CUDA(cudaFuncSetAttribute(mma_rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(n_blocks); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        SNB_CUDA(cudaLaunchKernelEx(&cfg, mma_rate_kernel<1>, mode, N, iters, d));
    }
    SNB_CUDA(cudaDeviceSynchronize());
    SNB_CUDA(cudaMemcpy(host_out, d, 16, cudaMemcpyDeviceToHost));
    cudaFree(d);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Ring microbenchmark: the fused kernels' weight ring in isolation.  One issuer thread runs `groups` stages of
// g = kstage/16 MMAs (M=128, N) each; stage st may be reissued only after the commit of its previous use has
// arrived (flags&2: and a producer thread has refilled it with a bulk copy of N*kstage*2 bytes from an L2-resident
// buffer).  A cycles through 8 K-slabs of a 128 KB tile, D alternates between two accumulators.  Reports cycles / MMA.
//   flags: 1 = pseudo-random operand data, 2 = real bulk copies, 4 = wait for EVERY commit before the next stage,
//          8 = no tcgen05.fence, 16 = no ring wait, 32 = commit every 4th stage only, 64 = fixed A slab / accumulator, 128 = no __syncwarp, 256 = no lane-0 wait block, 512 = accumulate flag 0 on the first MMA of a group
namespace snb {
__global__ void __launch_bounds__(128, 1) mma_ring_kernel(int N, int kstage, int depth, int groups, int flags,
                                                           const unsigned char* __restrict__ src, size_t src_bytes, long long* out) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint64_t full[16], empty[16], done; __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool random_data = flags & 1, copies = flags & 2, serial = flags & 4;
    const uint32_t stage_bytes = (uint32_t)N * (uint32_t)kstage * 2u;   // timing only: for kstage < 64 the SW128 reads run into the next stage
    const uint32_t copy_bytes = (uint32_t)N * (uint32_t)kstage * 2u;
    const uint32_t total = 131072u + (uint32_t)depth * stage_bytes + (uint32_t)N * 128u - stage_bytes;
    for (uint32_t i = threadIdx.x; i < total / 4; i += blockDim.x) {
        uint32_t h = i * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        reinterpret_cast<uint32_t*>(sm)[i] = random_data ? ((h & 0x83ff83ffu) | 0x34003400u) : 0x3c003c00u;
    }
    if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); } mbar_init(&done, 1); fence_barrier_init(); }
    __syncthreads();
    if (warp == 0) tmem_alloc(&tptr, 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = tptr;
    fence_proxy_async_smem();
    __syncthreads();
    const int g = kstage / 16;
    if (warp == 2 && lane == 0 && copies) {
        int st = 0; uint32_t ph = 0; size_t off = (size_t)blockIdx.x * 65536 % src_bytes;
        for (int i = 0; i < groups; ++i) {
            mbar_wait(&empty[st], ph ^ 1, 1);
            mbar_arrive_expect_tx(&full[st], copy_bytes);
            bulk_g2s(sm + 131072 + (size_t)st * stage_bytes, src + off, copy_bytes, &full[st]);
            off += copy_bytes; if (off + copy_bytes > src_bytes) off = 0;
            if (++st == depth) { st = 0; ph ^= 1; }
        }
    }
    if (warp == 1) {
        // same issue pattern as the fused kernels: converged warp, warp-uniform operands, elect for the tcgen05 ops
        const uint32_t a = smem_u32(sm), b = a + 131072u;
        const uint32_t idesc = umma_idesc_f16((uint32_t)N);
        const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
        const uint64_t a_desc0 = umma_desc_k_sw128(a), b_desc0 = umma_desc_k_sw128(b);
        int st = 0; uint32_t ph = 0;
        const long long t0 = clock64();
        for (int i = 0; i < groups; ++i) {
            if (!(flags & 256)) {
                if (lane == 0) {
                    if (copies) mbar_wait(&full[st], ph, 2);
                    else if (i >= depth && !(flags & 16)) mbar_wait(&empty[st], ph ^ 1, 3);
                }
            }
            if (!(flags & 128)) __syncwarp();
            if (!(flags & 8)) tc_fence_after();
            const uint32_t slab = (flags & 64) ? 0u : (uint32_t)((i * g / 4) & 7);
            const uint32_t d_tm = tm + ((flags & 64) ? 0u : (((i * g) >> 5) & 1) * 256u);
            const uint64_t da = a_desc0 + (uint64_t)(slab * (16384u >> 4)) + (uint64_t)(((uint32_t)(i * g) & 3u) * 2u);
            const uint64_t db = b_desc0 + (uint64_t)((uint32_t)st * (stage_bytes >> 4));
            if (elect_one()) {
                if (g == 4) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16_ss(d_tm, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (flags & 512) ? k : 1);
                } else if (g == 2) {
#pragma unroll
                    for (int k = 0; k < 2; ++k) umma_f16_ss(d_tm, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, 1);
                } else umma_f16_ss(d_tm, da, db, idesc, 1);
                if (!(flags & 32) || (i & 3) == 3) umma_commit(&empty[st]);
            }
            if (!(flags & 128)) __syncwarp();
            if (serial) { if (lane == 0) mbar_wait(&empty[st], ph, 4); __syncwarp(); }
            if (++st == depth) { st = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit(&done);
        __syncwarp();
        mbar_wait(&done, 0, 5);
        const long long t1 = clock64();
        if (blockIdx.x == 0 && lane == 0) { out[0] = t1 - t0; out[1] = (long long)groups * g; }
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}
}  // namespace snb

extern "C" SNB_API int snb_debug_mma_ring(int N, int kstage, int depth, int groups, int flags, int n_blocks, long long* host_out) {
    if (depth < 1 || depth > 16 || kstage % 16 || kstage > 64 || N % 16 || N > 256) SNB_FAIL(-2, "snb_debug_mma_ring: bad arguments");
    const size_t smem = 131072 + (size_t)depth * N * kstage * 2 + (size_t)N * (128 - kstage * 2);
    if (smem > 227 * 1024) SNB_FAIL(-2, "snb_debug_mma_ring: ring does not fit");
    long long* d = nullptr; unsigned char* src = nullptr; const size_t src_bytes = 6u << 20;
    SNB_CUDA(cudaMalloc(&d, 16)); SNB_CUDA(cudaMemset(d, 0, 16));
    SNB_CUDA(cudaMalloc(&src, src_bytes)); SNB_CUDA(cudaMemset(src, 0x3a, src_bytes));
    SNB_CUDA(cudaFuncSetAttribute(mma_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mma_ring_kernel<<<n_blocks, 128, smem>>>(N, kstage, depth, groups, flags, src, src_bytes, d);
    SNB_CUDA(cudaDeviceSynchronize());
    SNB_CUDA(cudaMemcpy(host_out, d, 16, cudaMemcpyDeviceToHost));
    cudaFree(d); cudaFree(src);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// CTA-pair ring microbenchmark (cta_group::2): each CTA of the pair streams ITS half of every weight stage (N/2 rows
// x 64 K fp16 = N*64 bytes) into its own ring; the leader issues M=256 MMAs that read both halves.  A relay warp in the
// peer forwards "stage landed" to the leader (peer_full); commits are multicast to both CTAs' empty barriers.
//   flags: 1 = random data, 8 = all 32 lanes poll (no lane-0 block / __syncwarp), 16 = two stages per issue iteration,
//          32 = commits arrive on the leader's barrier only and a relay warp forwards them to the peer, 64 = no copies (producer just arrives), 128 = fixed A slab / accumulator, 256 = leader does not wait for peer_full, 512 = nor for its own full
namespace snb {
__global__ void __launch_bounds__(128, 1) mma_ring2_kernel(int N, int depth, int groups, int flags,
                                                            const unsigned char* __restrict__ src, size_t src_bytes, long long* out) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ uint64_t full[16], empty[16], peer_full[16], done; __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool random_data = flags & 1, all_poll = flags & 8, pairs = flags & 16;
    const uint32_t stage_bytes = (uint32_t)N * 64u;            // this CTA's half of a stage
    const uint32_t total = 131072u + (uint32_t)depth * stage_bytes;
    for (uint32_t i = threadIdx.x; i < total / 4; i += blockDim.x) {
        uint32_t h = i * 2654435761u; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
        reinterpret_cast<uint32_t*>(sm)[i] = random_data ? ((h & 0x83ff83ffu) | 0x34003400u) : 0x3c003c00u;
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < 16; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); mbar_init(&peer_full[i], 1); }
        mbar_init(&done, 1); fence_barrier_init();
    }
    __syncthreads(); cluster_sync_all();
    if (warp == 0) tmem_alloc_2cta(&tptr, 512);
    tc_fence_before(); __syncthreads(); cluster_sync_all(); tc_fence_after();
    const uint32_t tmem = tptr;
    fence_proxy_async_smem();
    __syncthreads();
    if (warp == 2 && lane == 0) {              // producer (both CTAs)
        int st = 0; uint32_t ph = 0; size_t off = (size_t)(blockIdx.x >> 1) * 65536 % src_bytes;
        for (int i = 0; i < groups; ++i) {
            mbar_wait(&empty[st], ph ^ 1, 1);
            if (flags & 64) mbar_arrive(&full[st]);
            else {
                mbar_arrive_expect_tx(&full[st], stage_bytes);
                bulk_g2s(sm + 131072 + (size_t)st * stage_bytes, src + off + rank * stage_bytes, stage_bytes, &full[st]);
            }
            off += 2 * stage_bytes; if (off + 2 * stage_bytes > src_bytes) off = 0;
            if (++st == depth) { st = 0; ph ^= 1; }
        }
    }
    if (warp == 1 && rank == 1) {              // relay: my half of stage st has landed -> tell the leader
        int st = 0; uint32_t ph = 0;
        for (int i = 0; i < groups; ++i) {
            mbar_wait(&full[st], ph, 5);
            if (lane == 0) { if (flags & 1024) mbar_arrive_cluster_relaxed(mapa_u32(smem_u32(&peer_full[st]), 0)); else mbar_arrive_cluster(mapa_u32(smem_u32(&peer_full[st]), 0)); }
            __syncwarp();
            if (++st == depth) { st = 0; ph ^= 1; }
        }
    }
    if (warp == 3 && rank == 0 && (flags & 32)) {      // leader-only commits: forward "stage consumed" to the peer's producer
        int st = 0; uint32_t ph = 0;
        for (int i = 0; i < groups; ++i) {
            mbar_wait(&empty[st], ph, 7);
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&empty[st]), 1));
            __syncwarp();
            if (++st == depth) { st = 0; ph ^= 1; }
        }
    }
    const uint16_t cmask = (flags & 32) ? 1 : 3;
    if (warp == 1 && rank == 0) {
        const uint32_t a = smem_u32(sm), b = a + 131072u;
        const uint32_t idesc = umma_idesc_f16_m256((uint32_t)N);
        const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
        const uint64_t a_desc0 = umma_desc_k_sw128(a), b_desc0 = umma_desc_k_sw128(b);
        int st = 0; uint32_t ph = 0;
        const long long t0 = clock64();
        const int per_it = pairs ? 2 : 1;
        for (int i = 0; i < groups; i += per_it) {
            int st1 = st + 1; uint32_t ph1 = ph; if (st1 == depth) { st1 = 0; ph1 ^= 1; }
            const bool wp = !(flags & 256), wf = !(flags & 512);
            if (all_poll) {
                if (wf) mbar_wait(&full[st], ph, 2); if (wp) mbar_wait(&peer_full[st], ph, 6);
                if (pairs) { if (wf) mbar_wait(&full[st1], ph1, 2); if (wp) mbar_wait(&peer_full[st1], ph1, 6); }
            } else {
                if (lane == 0) {
                    if (wf) mbar_wait(&full[st], ph, 2); if (wp) mbar_wait(&peer_full[st], ph, 6);
                    if (pairs) { if (wf) mbar_wait(&full[st1], ph1, 2); if (wp) mbar_wait(&peer_full[st1], ph1, 6); }
                }
                __syncwarp();
            }
            tc_fence_after();
            const uint32_t slab = (flags & 128) ? 0u : (uint32_t)(i & 7);
            const uint32_t d_tm = tm + ((flags & 128) ? 0u : (uint32_t)((i >> 3) & 1) * 256u);
            const uint64_t da = a_desc0 + (uint64_t)(slab * (16384u >> 4));
            const uint64_t db = b_desc0 + (uint64_t)((uint32_t)st * (stage_bytes >> 4));
            if (elect_one()) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ss_2cta(d_tm, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, 1);
                umma_commit_2cta(&empty[st], cmask);
                if (pairs) {
                    const uint64_t da1 = a_desc0 + (uint64_t)(((slab + 1) & 7) * (16384u >> 4));
                    const uint64_t db1 = b_desc0 + (uint64_t)((uint32_t)st1 * (stage_bytes >> 4));
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_f16_ss_2cta(d_tm, da1 + (uint64_t)(2 * k), db1 + (uint64_t)(2 * k), idesc, 1);
                    umma_commit_2cta(&empty[st1], cmask);
                }
            }
            if (!all_poll) __syncwarp();
            for (int q = 0; q < per_it; ++q) if (++st == depth) { st = 0; ph ^= 1; }
        }
        if (elect_one()) umma_commit_2cta(&done, 1);
        __syncwarp();
        mbar_wait(&done, 0, 5);
        const long long t1 = clock64();
        if (blockIdx.x == 0 && lane == 0) { out[0] = t1 - t0; out[1] = (long long)groups * 4; }
    }
    tc_fence_before(); __syncthreads(); cluster_sync_all();
    if (warp == 0) tmem_dealloc_2cta(tmem, 512);
}
}  // namespace snb

extern "C" SNB_API int snb_debug_mma_ring2(int N, int depth, int groups, int flags, int n_blocks, long long* host_out) {
    if (depth < 2 || depth > 16 || (depth & 1) || N % 32 || N > 256 || (groups & 1)) SNB_FAIL(-2, "snb_debug_mma_ring2: bad arguments");
    const size_t smem = 131072 + (size_t)depth * N * 64;
    if (smem > 227 * 1024 - 1024) SNB_FAIL(-2, "snb_debug_mma_ring2: ring does not fit");
    long long* d = nullptr; unsigned char* src = nullptr; const size_t src_bytes = 6u << 20;
    SNB_CUDA(cudaMalloc(&d, 16)); SNB_CUDA(cudaMemset(d, 0, 16));
    SNB_CUDA(cudaMalloc(&src, src_bytes)); SNB_CUDA(cudaMemset(src, 0x3a, src_bytes));
    SNB_CUDA(cudaFuncSetAttribute(mma_ring2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(n_blocks < 2 ? 2 : (n_blocks & ~1)); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    SNB_CUDA(cudaLaunchKernelEx(&cfg, mma_ring2_kernel, N, depth, groups, flags, (const unsigned char*)src, src_bytes, d));
    SNB_CUDA(cudaDeviceSynchronize());
    SNB_CUDA(cudaMemcpy(host_out, d, 16, cudaMemcpyDeviceToHost));
    cudaFree(d); cudaFree(src);
    return 0;
}
