// Tensor-core backward building blocks (sm_100a).
//
// Point-atom layout ("atoms"): a (P points x F features) fp16 matrix is stored as
//     [tile = p/128][fg = f/64][pg = (p%128)/8][8 points][64 features], 16-byte chunks XOR-swizzled with (p%8)
// i.e. 1024-byte atoms of 8 points x 64 features.  This is byte-for-byte the 128B-swizzled activation tile of the
// fused forward kernel (K-major A operand, K = features) AND an MN-major UMMA operand with K = points, so one
// stored copy serves the input-gradient GEMMs (next layer's A tile) and the weight-gradient GEMMs below, and both
// are streamed with plain 1-D bulk copies.
//
// tc_dw_kernel: dW[m, n] = sum_p Ya[p, m] * Xb[p, n]   (weight gradient: Ya = dY of a layer, Xb = the layer's input)
//   one work item = 128 (m) x <=256 (n) output tile over a range of 128-point tiles; fp32 accumulation in TMEM;
//   split-K partials are written to a workspace and reduced in a fixed order by the caller (deterministic).
#include "tc_backward.cuh"
#include "sm100_ptx.cuh"

namespace snb {

using namespace ptx;

constexpr int kDwThreads = 192;            // warp 0 producer, warp 1 MMA, warps 2-5 epilogue
constexpr int kDwStages = 4;
constexpr int kDwStageBytes = 16384 + 32768;   // 64 points x (128 + 256) features

// row-major fp32 (P x F) -> atoms (zero padded to whole tiles / feature groups)
__global__ void atoms_pack_kernel(const float* __restrict__ src, int P, int F, int ld, __half* __restrict__ dst, int n_tiles, int n_fg) {
    long long total = (long long)n_tiles * n_fg * 128 * 8;     // 16-byte chunks
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < total; c += (long long)gridDim.x * blockDim.x) {
        int chunk = (int)(c & 7); long long r = c >> 3;
        int pl = (int)(r & 127); r >>= 7;
        int fg = (int)(r % n_fg); int tile = (int)(r / n_fg);
        int p = tile * 128 + pl;
        __half v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int f = fg * 64 + chunk * 8 + i;
            v[i] = __float2half_rn((p < P && f < F) ? src[(size_t)p * ld + f] : 0.f);
        }
        size_t off = ((size_t)(tile * n_fg + fg) * 16 + (pl >> 3)) * 1024 + (size_t)(pl & 7) * 128 + (size_t)((chunk ^ (pl & 7)) << 4);
        *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(dst) + off) = *reinterpret_cast<uint4*>(v);
    }
}

__global__ void __launch_bounds__(kDwThreads, 1) tc_dw_kernel(const DwItem* __restrict__ items, int n_items,
                                                               const unsigned char* __restrict__ base, float* __restrict__ partial) {
    extern __shared__ __align__(1024) unsigned char dsm[];
    unsigned char* ring = dsm;                                                  // kDwStages x [M-operand 16 KB | N-operand 32 KB]
    uint64_t* full = reinterpret_cast<uint64_t*>(dsm + kDwStages * kDwStageBytes);
    uint64_t* empty = full + kDwStages;
    uint64_t* acc_full = empty + kDwStages;
    uint64_t* acc_free = acc_full + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_free + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        if ((smem_u32(dsm) & 1023u) != 0) asm volatile("trap;");
        for (int i = 0; i < kDwStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(acc_full, 1); mbar_init(acc_free, 128);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;

    if (warp == 0) {
        int st = 0; uint32_t ph = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            const DwItem w = items[it];
            const uint32_t bytes = 16384u + (uint32_t)w.b_nfg * 8192u;
            for (int t = w.k_tile0; t < w.k_tile0 + w.k_tiles; ++t) {
                for (int hf = 0; hf < 2; ++hf) {                 // two 64-point stages per 128-point tile
                    if (lane == 0) {
                        mbar_wait(&empty[st], ph ^ 1, 11);
                        mbar_arrive_expect_tx(&full[st], bytes);
                        unsigned char* dst = ring + (size_t)st * kDwStageBytes;
                        for (int g = 0; g < 2; ++g)
                            bulk_g2s(dst + g * 8192, base + w.a_off + ((size_t)t * w.a_fgs + w.a_fg0 + g) * 16384 + hf * 8192, 8192, &full[st]);
                        for (int g = 0; g < w.b_nfg; ++g)
                            bulk_g2s(dst + 16384 + g * 8192, base + w.b_off + ((size_t)t * w.b_fgs + w.b_fg0 + g) * 16384 + hf * 8192, 8192, &full[st]);
                    }
                    __syncwarp();
                    if (++st == kDwStages) { st = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        int st = 0; uint32_t ph = 0, free_ph = 0;
        const uint32_t ring_addr = smem_u32(ring);
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            const DwItem w = items[it];
            const uint32_t idesc = umma_idesc_f16_mn((uint32_t)w.b_nfg * 64u);
            mbar_wait(acc_free, free_ph ^ 1, 12); free_ph ^= 1;      // epilogue of the previous item has drained TMEM
            tc_fence_after();
            bool first = true;
            for (int s = 0; s < 2 * w.k_tiles; ++s) {
                mbar_wait(&full[st], ph, 13);
                tc_fence_after();
                if (lane == 0) {
                    const uint32_t a_addr = ring_addr + (uint32_t)st * kDwStageBytes, b_addr = a_addr + 16384u;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {                    // 16 points = two 8-point atoms per step
                        umma_f16_ss(tmem, umma_desc_mn_sw128(a_addr + k * 2048, 8192, 1024), umma_desc_mn_sw128(b_addr + k * 2048, 8192, 1024),
                                    idesc, first ? 0u : 1u);
                        first = false;
                    }
                    umma_commit(&empty[st]);
                    if (s == 2 * w.k_tiles - 1) umma_commit(acc_full);
                }
                first = false;
                __syncwarp();
                if (++st == kDwStages) { st = 0; ph ^= 1; }
            }
        }
    } else {
        const int quad = warp & 3, row = quad * 32 + lane;
        uint32_t acc_ph = 0;
        for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
            const DwItem w = items[it];
            const int N = w.b_nfg * 64;
            mbar_wait(acc_full, acc_ph, 14); acc_ph ^= 1;
            tc_fence_after();
            float* out = partial + w.out_off + (size_t)row * N;
            for (int n0 = 0; n0 < N; n0 += 32) {
                float v[32];
                tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)n0, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(out + n0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
            tc_fence_before();
            mbar_arrive(acc_free);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 256);
}

int launch_atoms_pack(const float* src, int P, int F, int ld, void* dst, cudaStream_t st) {
    int n_tiles = (P + 127) / 128, n_fg = (F + 63) / 64;
    long long chunks = (long long)n_tiles * n_fg * 128 * 8;
    int blocks = (int)((chunks + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
    atoms_pack_kernel<<<blocks, 256, 0, st>>>(src, P, F, ld, (__half*)dst, n_tiles, n_fg);
    SNB_CHECK_LAUNCH();
    return 0;
}

int launch_dw(const DwItem* d_items, int n_items, const void* base, float* partial, cudaStream_t st) {
    static int sm_count = 0;
    if (!sm_count) { int dev = 0; SNB_CUDA(cudaGetDevice(&dev)); SNB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev)); }
    size_t smem = (size_t)kDwStages * kDwStageBytes + 256;
    SNB_CUDA(cudaFuncSetAttribute(tc_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = n_items < sm_count ? n_items : sm_count;
    tc_dw_kernel<<<grid, kDwThreads, smem, st>>>(d_items, n_items, (const unsigned char*)base, partial);
    SNB_CHECK_LAUNCH();
    return 0;
}

}  // namespace snb

#ifdef SNB_DEV_BUILD
#include "satnerf_b200_dev.h"
using namespace snb;

// Developer / test entry: out (Fa x Fb, fp32 row-major) = Xa^T Xb for row-major fp32 Xa (P x Fa), Xb (P x Fb), through the
// atom packing and the tensor-core weight-gradient kernel with `k_splits` split-K partials.  Fa % 128 == 0, Fb % 64 == 0.
extern "C" SNB_API int snb_debug_dw_gemm(const float* xa, const float* xb, int P, int Fa, int Fb, int k_splits, float* out,
                                         void* workspace, size_t workspace_bytes, void* stream) {
    if (!xa || !xb || !out || !workspace) SNB_FAIL(-1, "snb_debug_dw_gemm: null pointer");
    if (Fa % 128 || Fb % 64 || P < 1 || k_splits < 1) SNB_FAIL(-1, "snb_debug_dw_gemm: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const int n_tiles = (P + 127) / 128, fga = Fa / 64, fgb = Fb / 64;
    if (k_splits > n_tiles) k_splits = n_tiles;
    Arena ar(workspace, workspace_bytes);
    unsigned char* a_at = ar.take<unsigned char>((size_t)n_tiles * fga * 16384);
    unsigned char* b_at = ar.take<unsigned char>((size_t)n_tiles * fgb * 16384);
    const int m_tiles = Fa / 128, n_chunks = (fgb + 3) / 4;
    const int n_items = m_tiles * n_chunks * k_splits;
    DwItem* d_items = ar.take<DwItem>(n_items);
    float* partial = ar.take<float>((size_t)n_items * 128 * 256);
    if (ar.overflow) SNB_FAIL(-4, "snb_debug_dw_gemm: workspace too small");
    SNB_TRY(launch_atoms_pack(xa, P, Fa, Fa, a_at, st));
    SNB_TRY(launch_atoms_pack(xb, P, Fb, Fb, b_at, st));
    DwItem* h = new DwItem[n_items]; int n = 0;
    for (int ks = 0; ks < k_splits; ++ks)
        for (int m = 0; m < m_tiles; ++m)
            for (int c = 0; c < n_chunks; ++c) {
                DwItem w; memset(&w, 0, sizeof(w));
                w.a_off = (long long)(a_at - (unsigned char*)workspace); w.b_off = (long long)(b_at - (unsigned char*)workspace);
                w.a_fgs = fga; w.b_fgs = fgb; w.a_fg0 = 2 * m; w.b_fg0 = 4 * c; w.b_nfg = fgb - 4 * c < 4 ? fgb - 4 * c : 4;
                w.k_tile0 = (int)((long long)n_tiles * ks / k_splits); w.k_tiles = (int)((long long)n_tiles * (ks + 1) / k_splits) - w.k_tile0;
                w.out_off = (long long)n * 128 * 256;
                h[n++] = w;
            }
    SNB_CUDA(cudaMemcpyAsync(d_items, h, sizeof(DwItem) * n_items, cudaMemcpyHostToDevice, st));
    SNB_CUDA(cudaStreamSynchronize(st));
    SNB_TRY(launch_dw(d_items, n_items, workspace, partial, st));
    // reduce partials on the host side of this debug entry (tiny): copy back and sum
    float* hp = new float[(size_t)n_items * 128 * 256];
    SNB_CUDA(cudaMemcpyAsync(hp, partial, sizeof(float) * (size_t)n_items * 128 * 256, cudaMemcpyDeviceToHost, st));
    SNB_CUDA(cudaStreamSynchronize(st));
    float* ho = new float[(size_t)Fa * Fb]; memset(ho, 0, sizeof(float) * (size_t)Fa * Fb);
    for (int i = 0; i < n_items; ++i) {
        const DwItem& w = h[i]; const int N = w.b_nfg * 64;
        for (int r = 0; r < 128; ++r)
            for (int c = 0; c < N; ++c)
                ho[(size_t)(w.a_fg0 * 64 + r) * Fb + w.b_fg0 * 64 + c] += hp[w.out_off + (size_t)r * N + c];
    }
    SNB_CUDA(cudaMemcpyAsync(out, ho, sizeof(float) * (size_t)Fa * Fb, cudaMemcpyHostToDevice, st));
    SNB_CUDA(cudaStreamSynchronize(st));
    delete[] h; delete[] hp; delete[] ho;
    return 0;
}
#endif  // SNB_DEV_BUILD
