/*
 * Developer entry points of libsatnerf_b200_dev.so — the product library (libsatnerf_b200.so, include/satnerf_b200.h)
 * built with -DSNB_DEV_BUILD plus the microbenchmarks of csrc/mma_rate.cu.  Nothing here is part of the drop-in ABI;
 * profiles/*.py and one unit test (tests/test_gpu_tc_backward_blocks.py) use it.
 */
#ifndef SATNERF_B200_DEV_H
#define SATNERF_B200_DEV_H
#include "satnerf_b200.h"
#ifdef __cplusplus
extern "C" {
#endif
/* copies the fused kernel's phase timestamps (int64 clock64 values; probe builds) to a HOST buffer */
SNB_API int snb_debug_read(void* host_dst, size_t bytes);
/* with env SNB_TC_HANG_MIRROR set (read once at load): 192 uints = 16 x [code, block, thread, parity] of the bounded barrier
 * waits that trapped, by wait code (0xffffffff: none), followed by per-block progress words of SNB_TC_PROGRESS builds */
SNB_API int snb_debug_hang_info(unsigned int* out192);
/* out (Fa x Fb) = Xa^T Xb through the point-atom packing and the tensor-core weight-gradient kernel (csrc/tc_backward.cu);
 * Xa (P x Fa), Xb (P x Fb) fp32 row-major, Fa % 64 == 0, Fb % 64 == 0 */
SNB_API int snb_debug_dw_gemm(const float* xa, const float* xb, int P, int Fa, int Fb, int k_splits, float* out,
                              void* workspace, size_t workspace_bytes, void* stream);
/* cycles for `iters` back-to-back tcgen05.mma (M=128 or 256, K=16) on n_blocks CTAs; mode 0 SS cg1, 1 TS cg1 (A in TMEM),
 * 2 SS cg2, 3 SS cg1 MN-major; host_out[0] = cycles, host_out[1] = iters */
SNB_API int snb_debug_mma_rate(int mode, int N, int iters, int n_blocks, long long* host_out);
SNB_API int snb_debug_mma_ring2(int N, int depth, int groups, int flags, int n_blocks, long long* host_out);
SNB_API int snb_debug_mma_ring(int N, int kstage, int depth, int groups, int flags, int n_blocks, long long* host_out);
#ifdef __cplusplus
}
#endif
#endif
