// Developer microbenchmark: cycles per tcgen05.mma.kind::f16 instruction issued back to back (operands: whatever is in
// shared memory / TMEM; only timing matters).  mode 0: SS cg1 (A,B from smem)   mode 1: TS cg1 (A from TMEM)
//                                               mode 2: SS cg2 (M=256, leader issues)   mode 3: SS cg1, MN-major operands
#include "common.cuh"
#include "sm100_ptx.cuh"

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
