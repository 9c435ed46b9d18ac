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
//   CTA pairs (cta_group::2): one work item = 256 (m; 128 per CTA) x <= 512 (n) output tile -- all 512 TMEM columns -- over a
//   range of 128-point tiles.  Per 64-point stage each CTA streams 16 KB of its M operand and its half of the N operand
//   (<= 32 KB) for 8 MMAs (M256 x N256 x K16): 48 B per tensor cycle and SM, against 96 for the round-1 128 x 256 tiles --
//   the kernel is bound by the L2 -> shared-memory operand feed.  fp32 accumulation in TMEM; split-K partials go to a
//   workspace and are reduced in a fixed order by the caller (deterministic).  Small-N work (bias / extra-input columns, the
//   N <= 3 heads) runs as multi-accumulator pieces next to the main pieces that read the same arrays (see DwItem).
#include "tc_backward.cuh"
#include "sm100_ptx.cuh"
#include <vector>
#include <cstring>

namespace snb {

using namespace ptx;

constexpr int kDwThreads = 192;            // warp 0 producer, warp 1 MMA issuer (leader) / arrival relay (peer), warps 2-5 epilogue
constexpr int kDwStages = 4;
constexpr int kDwStageBytes = 3 * 16384 + 8192;   // per CTA and 64-point stage: main [M 16K | N chunk 0 16K | N chunk 1 16K]; small [M0 | M1 | M2 | N 8K]

// kind::f16 instruction descriptor, both operands MN-major (fp16), D = fp32, M = 256 (CTA pair)
__device__ __forceinline__ uint32_t umma_idesc_f16_mn_m256(uint32_t n) {
    return (1u << 4) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) | ((256u >> 4) << 24);
}

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

struct DwChunks { int n_chunks, c[2], ce[2]; };       // groups per chunk, and rounded up to even (the pair splits N in halves)
__device__ __forceinline__ DwChunks dw_chunks(int b_nfg) {
    DwChunks k; k.c[0] = b_nfg > 4 ? 4 : b_nfg; k.c[1] = b_nfg - k.c[0]; k.n_chunks = k.c[1] > 0 ? 2 : 1;
    k.ce[0] = (k.c[0] + 1) & ~1; k.ce[1] = (k.c[1] + 1) & ~1;
    return k;
}

__global__ void __launch_bounds__(kDwThreads, 1) tc_dw_kernel(const DwItem* __restrict__ items, int n_items,
                                                               const unsigned char* __restrict__ base, float* __restrict__ partial, int* __restrict__ sync) {
    extern __shared__ __align__(1024) unsigned char dsm[];
    unsigned char* ring = dsm;
    uint64_t* full = reinterpret_cast<uint64_t*>(dsm + kDwStages * kDwStageBytes);
    uint64_t* empty = full + kDwStages;
    uint64_t* acc_full = empty + kDwStages;
    uint64_t* acc_free = acc_full + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_free + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = (int)(blockIdx.x >> 1), n_pairs = (int)(gridDim.x >> 1);

    if (threadIdx.x == 0) {
        if ((smem_u32(dsm) & 1023u) != 0) asm volatile("trap;");
        // leader: a stage is full when its own copies have landed AND the peer has reported its own (relay arrival)
        for (int i = 0; i < kDwStages; ++i) { mbar_init(&full[i], rank == 0 ? 2 : 1); mbar_init(&empty[i], 1); }
        mbar_init(acc_full, 1); mbar_init(acc_free, 2);          // acc_free (leader's): one arrival per CTA once its epilogue has drained TMEM
        fence_barrier_init();
    }
    __syncthreads(); cluster_sync_all();
    if (warp == 1) tmem_alloc_2cta(tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;

    if (warp == 0) {
        // ---- producer: this CTA's 128 rows of the M operand and its half of every N chunk, 64 points per stage ----
        int st = 0; uint32_t ph = 0;
        for (int it = pair; it < n_items; it += n_pairs) {
            const DwItem w = items[it];
            const DwChunks ck = dw_chunks(w.b_nfg);
            const uint32_t bytes = w.n_acc == 1 ? 16384u + 8192u * (uint32_t)(ck.ce[0] / 2 + ck.ce[1] / 2) : 16384u * (uint32_t)w.n_acc + 8192u;
            int s_idx = 0;
            for (int t = w.k_tile0; t < w.k_tile0 + w.k_tiles; ++t) {
                for (int hf = 0; hf < 2; ++hf, ++s_idx) {        // two 64-point stages per 128-point tile
                    if (lane == 0) {
                        if (sync && w.sync_n > 1 && rank == 0 && (s_idx % kDwSyncEvery) == 0) {      // pacing (see kDwSyncEvery); the peer CTA is held by the ring
                            // arrive on this checkpoint, wait for the PREVIOUS one: the slowest member never stalls, the others
                            // lead it by at most two checkpoints, and the atomic's round trip is off the critical path
                            int* c = sync + w.sync_off + s_idx / kDwSyncEvery;
                            atomicAdd(c, 1);
                            if (s_idx) {
                                const long long t0 = clock64();
                                while (*reinterpret_cast<volatile int*>(c - 1) < w.sync_n && clock64() - t0 < 40000) { }
                            }
                        }
                        mbar_wait(&empty[st], ph ^ 1, 11);
                        mbar_arrive_expect_tx(&full[st], bytes);
                        unsigned char* dst = ring + (size_t)st * kDwStageBytes;
                        for (int i = 0; i < w.n_acc; ++i)
                            for (int g = 0; g < 2; ++g) {        // (a group past the end of the block is a copy of the last one: its rows are ignored)
                                int fg = 2 * (int)rank + g; if (fg > w.a_nfg[i] - 1) fg = w.a_nfg[i] - 1;
                                bulk_g2s(dst + i * 16384 + g * 8192, base + w.a_off[i] + ((size_t)t * w.a_fgs[i] + w.a_fg0[i] + fg) * 16384 + hf * 8192, 8192, &full[st]);
                            }
                        if (w.n_acc == 1) {
                            for (int j = 0, g0 = 0; j < ck.n_chunks; g0 += ck.c[j], ++j)
                                for (int g = 0; g < ck.ce[j] / 2; ++g) {
                                    int fg = (int)rank * (ck.ce[j] / 2) + g; if (fg > ck.c[j] - 1) fg = ck.c[j] - 1;
                                    bulk_g2s(dst + 16384 + j * 16384 + g * 8192, base + w.b_off + ((size_t)t * w.b_fgs + w.b_fg0 + g0 + fg) * 16384 + hf * 8192, 8192, &full[st]);
                                }
                        } else      // one N group, the same in both CTAs: the pair's N = 128 columns are [group | group]
                            bulk_g2s(dst + 49152, base + w.b_off + ((size_t)t * w.b_fgs + w.b_fg0) * 16384 + hf * 8192, 8192, &full[st]);
                    }
                    __syncwarp();
                    if (++st == kDwStages) { st = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        int st = 0; uint32_t ph = 0;
        if (rank == 1) {
            // relay: report "my part of the stage has landed" to the leader's stage barrier (relaxed, see tc_pipeline.cuh)
            const uint32_t leader_full = mapa_u32(smem_u32(full), 0);
            for (int it = pair; it < n_items; it += n_pairs) {
                const int n = 2 * items[it].k_tiles;
                for (int s = 0; s < n; ++s) {
                    mbar_wait(&full[st], ph, 15);
                    if (lane == 0) mbar_arrive_cluster_relaxed(leader_full + (uint32_t)st * 8u);
                    __syncwarp();
                    if (++st == kDwStages) { st = 0; ph ^= 1; }
                }
            }
        } else {
            uint32_t free_ph = 0;
            const uint32_t ring_addr = smem_u32(ring);
            for (int it = pair; it < n_items; it += n_pairs) {
                const DwItem w = items[it];
                const DwChunks ck = dw_chunks(w.b_nfg);
                const uint32_t idesc0 = umma_idesc_f16_mn_m256((uint32_t)ck.ce[0] * 64u), idesc1 = umma_idesc_f16_mn_m256((uint32_t)ck.ce[1] * 64u);
                const uint32_t idesc_small = umma_idesc_f16_mn_m256(128u);
                mbar_wait(acc_free, free_ph ^ 1, 12); free_ph ^= 1;      // both CTAs' epilogues of the previous item have drained TMEM
                tc_fence_after();
                const int n = 2 * w.k_tiles;
                for (int s = 0; s < n; ++s) {
                    mbar_wait(&full[st], ph, 13);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a_addr = ring_addr + (uint32_t)st * kDwStageBytes;
                        const uint32_t acc = s ? 1u : 0u;
                        if (w.n_acc == 1) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)                  // 16 points = two 8-point atoms per K-step
                                umma_f16_ss_2cta(tmem, umma_desc_mn_sw128(a_addr + k * 2048, 8192, 1024), umma_desc_mn_sw128(a_addr + 16384 + k * 2048, 8192, 1024),
                                                 idesc0, k ? 1u : acc);
                            if (ck.n_chunks == 2) {
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    umma_f16_ss_2cta(tmem + 256u, umma_desc_mn_sw128(a_addr + k * 2048, 8192, 1024), umma_desc_mn_sw128(a_addr + 32768 + k * 2048, 8192, 1024),
                                                     idesc1, k ? 1u : acc);
                            }
                        } else {
                            for (int i = 0; i < w.n_acc; ++i) {
#pragma unroll
                                for (int k = 0; k < 4; ++k)
                                    umma_f16_ss_2cta(tmem + 128u * (uint32_t)i, umma_desc_mn_sw128(a_addr + i * 16384 + k * 2048, 8192, 1024),
                                                     umma_desc_mn_sw128(a_addr + 49152 + k * 2048, 8192, 1024), idesc_small, k ? 1u : acc);
                            }
                        }
                        umma_commit_2cta(&empty[st], 3);
                        if (s == n - 1) umma_commit_2cta(acc_full, 3);
                    }
                    __syncwarp();
                    if (++st == kDwStages) { st = 0; ph ^= 1; }
                }
            }
        }
    } else {
        // ---- epilogue: this CTA's 128 rows of the [256][N] partial ----
        const int quad = warp & 3, row = quad * 32 + lane;
        const uint32_t free_bar = mapa_u32(smem_u32(acc_free), 0);
        uint32_t acc_ph = 0;
        for (int it = pair; it < n_items; it += n_pairs) {
            const DwItem w = items[it];
            const DwChunks ck = dw_chunks(w.b_nfg);
            const int N = w.b_nfg * 64;
            mbar_wait(acc_full, acc_ph, 14); acc_ph ^= 1;
            tc_fence_after();
            if (w.n_acc == 1) {
                float* out = partial + w.out_off[0] + (size_t)((int)rank * 128 + row) * N;
                for (int j = 0, g0 = 0; j < ck.n_chunks; g0 += ck.c[j], ++j) {
                    for (int n0 = 0; n0 < ck.c[j] * 64; n0 += 32) {
                        float v[32];
                        tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(j * 256 + n0), v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(out + g0 * 64 + n0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    }
                }
            } else {
                for (int a = 0; a < w.n_acc; ++a) {
                    float* out = partial + w.out_off[a] + (size_t)((int)rank * 128 + row) * 64;
                    for (int n0 = 0; n0 < 64; n0 += 32) {
                        float v[32];
                        tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(a * 128 + n0), v);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(out + n0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    }
                }
            }
            tc_fence_before();
            named_bar_sync(1, 128);
            if (threadIdx.x == 64) mbar_arrive_cluster(free_bar);        // (release: orders this CTA's TMEM reads before the leader's next MMAs)
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_2cta(tmem, 512);
}

int launch_atoms_pack(const float* src, int P, int F, int ld, void* dst, cudaStream_t st) {
    int n_tiles = (P + 127) / 128, n_fg = (F + 63) / 64;
    long long chunks = (long long)n_tiles * n_fg * 128 * 8;
    int blocks = (int)((chunks + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
    atoms_pack_kernel<<<blocks, 256, 0, st>>>(src, P, F, ld, (__half*)dst, n_tiles, n_fg);
    SNB_CHECK_LAUNCH();
    return 0;
}

int launch_dw(const DwItem* d_items, int n_items, const void* base, float* partial, int* sync, cudaStream_t st) {
    static int sm_count = 0;
    if (!sm_count) { int dev = 0; SNB_CUDA(cudaGetDevice(&dev)); SNB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev)); }
    int n_pairs = sm_count / 2; if (n_pairs > kDwPairs) n_pairs = kDwPairs; if (n_pairs > n_items) n_pairs = n_items;
    if (n_pairs < 1) return 0;
    size_t smem = (size_t)kDwStages * kDwStageBytes + 256;
    SNB_CUDA(cudaFuncSetAttribute(tc_dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(2 * n_pairs); cfg.blockDim = dim3(kDwThreads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    SNB_CUDA(cudaLaunchKernelEx(&cfg, tc_dw_kernel, d_items, n_items, (const unsigned char*)base, partial, sync));
    ++g_launches;
    return 0;
}

}  // namespace snb

#ifdef SNB_DEV_BUILD
#include "satnerf_b200_dev.h"
using namespace snb;

// Developer / test entry: out (Fa x Fb, fp32 row-major) = Xa^T Xb for row-major fp32 Xa (P x Fa), Xb (P x Fb), through the
// atom packing and the tensor-core weight-gradient kernel with `k_splits` split-K partials.  Fa % 64 == 0, Fb % 64 == 0.
extern "C" SNB_API int snb_debug_dw_gemm(const float* xa, const float* xb, int P, int Fa, int Fb, int k_splits, float* out,
                                         void* workspace, size_t workspace_bytes, void* stream) {
    if (!xa || !xb || !out || !workspace) SNB_FAIL(-1, "snb_debug_dw_gemm: null pointer");
    if (Fa % 64 || Fb % 64 || P < 1 || k_splits < 1) SNB_FAIL(-1, "snb_debug_dw_gemm: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const int n_tiles = (P + 127) / 128, fga = Fa / 64, fgb = Fb / 64;
    if (k_splits > n_tiles) k_splits = n_tiles;
    Arena ar(workspace, workspace_bytes);
    unsigned char* a_at = ar.take<unsigned char>((size_t)n_tiles * fga * 16384);
    unsigned char* b_at = ar.take<unsigned char>((size_t)n_tiles * fgb * 16384);
    const int m_blocks = (fga + 3) / 4, n_blocks = (fgb + 7) / 8;
    const int n_items = m_blocks * n_blocks * k_splits;
    DwItem* d_items = ar.take<DwItem>(n_items);
    float* partial = ar.take<float>((size_t)n_items * 256 * 512);
    if (ar.overflow) SNB_FAIL(-4, "snb_debug_dw_gemm: workspace too small");
    SNB_TRY(launch_atoms_pack(xa, P, Fa, Fa, a_at, st));
    SNB_TRY(launch_atoms_pack(xb, P, Fb, Fb, b_at, st));
    std::vector<DwItem> pieces; int n = 0;
    for (int ks = 0; ks < k_splits; ++ks)
        for (int m = 0; m < m_blocks; ++m)
            for (int c = 0; c < n_blocks; ++c) {
                DwItem w; memset(&w, 0, sizeof(w));
                w.n_acc = 1;
                w.a_off[0] = (long long)(a_at - (unsigned char*)workspace); w.b_off = (long long)(b_at - (unsigned char*)workspace);
                w.a_fgs[0] = fga; w.b_fgs = fgb; w.a_fg0[0] = 4 * m; w.a_nfg[0] = fga - 4 * m < 4 ? fga - 4 * m : 4;
                w.b_fg0 = 8 * c; w.b_nfg = fgb - 8 * c < 8 ? fgb - 8 * c : 8;
                w.k_tile0 = (int)((long long)n_tiles * ks / k_splits); w.k_tiles = (int)((long long)n_tiles * (ks + 1) / k_splits) - w.k_tile0;
                w.out_off[0] = (long long)(n++) * 256 * 512;
                pieces.push_back(w);
            }
    SNB_CUDA(cudaMemcpyAsync(d_items, pieces.data(), sizeof(DwItem) * n_items, cudaMemcpyHostToDevice, st));
    SNB_CUDA(cudaStreamSynchronize(st));
    SNB_TRY(launch_dw(d_items, n_items, workspace, partial, nullptr, st));
    // reduce partials on the host side of this debug entry (tiny): copy back and sum
    std::vector<float> hp((size_t)n_items * 256 * 512);
    SNB_CUDA(cudaMemcpyAsync(hp.data(), partial, sizeof(float) * hp.size(), cudaMemcpyDeviceToHost, st));
    SNB_CUDA(cudaStreamSynchronize(st));
    std::vector<float> ho((size_t)Fa * Fb, 0.f);
    for (const DwItem& w : pieces) {
        const int N = w.b_nfg * 64;
        for (int r = 0; r < w.a_nfg[0] * 64; ++r)
            for (int c = 0; c < N; ++c)
                ho[(size_t)(w.a_fg0[0] * 64 + r) * Fb + w.b_fg0 * 64 + c] += hp[w.out_off[0] + (size_t)r * N + c];
    }
    SNB_CUDA(cudaMemcpyAsync(out, ho.data(), sizeof(float) * ho.size(), cudaMemcpyHostToDevice, st));
    SNB_CUDA(cudaStreamSynchronize(st));
    return 0;
}
#endif  // SNB_DEV_BUILD
