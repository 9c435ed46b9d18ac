// Tensor-core backward of the render pass (SNB_FP16_TC, sat-nerf / s-nerf):
//
//   1. composite_bwd (composite.cu)      d_head (P x C, fp32): gradient w.r.t. the pre-activation head outputs
//   2. tc_bwd_kernel (this file)         fused INPUT-gradient chain per 128-point tile, same warp-specialised tile machinery
//                                        as the forward: A operand = dY tile in shared memory, transposed weight tiles
//                                        streamed by the producer warp, accumulators in TMEM, epilogue multiplies by
//                                        cos(y) (y from the forward's stash) and writes the next dY tile; every dY tile is
//                                        also dumped (bulk S2G) in the point-atom layout for step 3
//   3. tc_dw_kernel (tc_backward.cu)     WEIGHT gradients dW = dY^T [a | x sun t 1] as split-K GEMMs over the points
//   4. finalize kernels (this file)      ordered reduction of the split-K partials, un-scaling, scatter-add into the flat
//                                        gradient buffer; tiny head biases, sky-colour MLP and embedding gradients
//
// Gradients travel in fp16 with one global power-of-two loss scale (taken from max |d_head|) and are accumulated in fp32.
// Everything is deterministic: no float atomics, fixed reduction orders.
#include "tc_common.cuh"
#include "tc_pipeline.cuh"
#include "tc_backward.cuh"
#include "tc_field.cuh"
#include "composite.cuh"
#include <vector>

namespace snb {

enum { BK_S2 = 0, BK_S1, BK_FA, BK_FB, BK_A7, BK_TRUNK };

struct TcBwdStash {                 // dY arrays written by the chain kernel (atoms), byte offsets from the workspace base
    long long dy[kMaxTrunk], df, dr1y, ds1y, ds2y, ds3y, db1y, dhead;
    long long total;
};

struct TcBwdArgs {
    TcProgram prog;                 // backward GEMM list (kind = BK_*), transposed weight tiles
    TcStash fs; const unsigned char* fbase;       // forward stash
    TcBwdStash bs; unsigned char* bbase;          // backward stash (workspace)
    const float* d_head; int C;                   // (P, C) fp32
    const float* absmax;                          // device scalar: max |d_head|
    float* d_t;                                   // (P, tau) fp32, unscaled; or null
    unsigned char* packed;
    int n_layers, R, S, G, n_groups, tiles_per_group;
    int dbg;                                      // developer what-if knobs (dev library only, see tc_common.cuh)
};

__device__ __forceinline__ float loss_scale(float absmax) {      // power of two that brings max |d_head| to ~64
    if (!(absmax > 0.f) || !isfinite(absmax)) return 1.f;
    int e; frexpf(absmax, &e);                                    // absmax = m * 2^e, m in [0.5, 1)
    return ldexpf(1.f, 6 - e);
}

// ---------------------------------------------------------------------------------------------------------------
// packing: transposed weight tiles B[n][k] = W[k][col + n] (sources split along k) + epilogue tables
// ---------------------------------------------------------------------------------------------------------------
__global__ void tc_bwd_pack_kernel(TcProgram P, const float* __restrict__ W, unsigned char* __restrict__ packed) {
    const int gi = blockIdx.y;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    const TcGemm g = P.g[gi];
    long long base = 0;
    for (int i = 0; i < gi; ++i) base += (long long)P.g[i].n_chunks * P.g[i].k_slabs * P.g[i].chunk_n * 128;
    __half* out = reinterpret_cast<__half*>(packed + base);
    // one thread per 16-byte chunk (8 k of one output row n); consecutive threads take consecutive n, so that each of the 8
    // strided reads W[k][col + n] is coalesced across the warp
    const long long total = (long long)g.n_chunks * g.k_slabs * g.chunk_n * 8;
    for (long long e = tid; e < total; e += nthr) {
        const int nl = (int)(e % g.chunk_n); long long r = e / g.chunk_n;
        const int c = (int)(r & 7); r >>= 3;
        const int s = (int)(r % g.k_slabs), j = (int)(r / g.k_slabs);
        const int n = j * g.chunk_n + nl, k0 = s * 64 + c * 8;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k = k0 + i;
            v[i] = k >= g.K ? 0.f : (k < g.rows0 ? W[g.src0 + (long long)k * g.ld0 + g.col0 + n] : W[g.src1 + (long long)(k - g.rows0) * g.ld1 + g.col1 + n]);
        }
        const size_t tile = (size_t)(j * g.k_slabs + s) * g.chunk_n * 64;
        *reinterpret_cast<uint4*>(out + tile + (size_t)nl * 64 + (size_t)((c ^ (nl & 7)) << 3)) =
            make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
    }
}

struct BwdMisc { long long s3_w, r2_w, b2_w, b0_w, sigma_w; int b0_ld, H, H2, tau, has_beta; int t_seed, t_r2, t_beta, t_betav, t_sigma; };

__global__ void tc_bwd_tables_kernel(BwdMisc M, long long tables_base, const float* __restrict__ W, unsigned char* __restrict__ packed) {
    float* T = reinterpret_cast<float*>(packed + tables_base);
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    for (int n = tid; n < M.H2; n += nthr) {
        T[M.t_seed + n] = W[M.s3_w + n];                                        // sun_v_net.6 weight (1 x H2)
        float* r = T + M.t_r2 + n * 4;                                         // rgb_from_xyzdir.2 weight (3 x H2) as [n][4]
        r[0] = W[M.r2_w + n]; r[1] = W[M.r2_w + M.H2 + n]; r[2] = W[M.r2_w + 2 * M.H2 + n]; r[3] = 0.f;
        if (M.has_beta) {
            float* b = T + M.t_beta + n * 4;                                   // [beta2 weight, W_b0[n][H], [H+1], [H+2]]
            b[0] = W[M.b2_w + n];
            for (int c = 0; c < 3; ++c) b[1 + c] = c < M.tau ? W[M.b0_w + (long long)n * M.b0_ld + M.H + c] : 0.f;
            T[M.t_betav + n] = M.tau > 3 ? W[M.b0_w + (long long)n * M.b0_ld + M.H + 3] : 0.f;
        }
    }
    for (int n = tid; n < M.H; n += nthr) {
        T[M.t_sigma + n] = W[M.sigma_w + n];
    }
}

// ---------------------------------------------------------------------------------------------------------------
// the fused input-gradient chain
// ---------------------------------------------------------------------------------------------------------------
struct YBuf { uint4 q[4]; };
__device__ __forceinline__ YBuf yb_load(const unsigned char* arr, int gt, int F, int n0, int row) {      // 32 columns from n0 (n0 % 32 == 0)
    const unsigned char* s = arr + (((size_t)gt * (F >> 3) + (n0 >> 3)) * kTile + row) * 16;                 // [gt][n/8][row][8 fp16]: coalesced over rows
    YBuf b;
#pragma unroll
    for (int c = 0; c < 4; ++c) b.q[c] = __ldg(reinterpret_cast<const uint4*>(s + (size_t)c * kTile * 16));
    return b;
}
__device__ __forceinline__ YBuf yb_zero() { YBuf b; for (int c = 0; c < 4; ++c) b.q[c] = make_uint4(0u, 0u, 0u, 0u); return b; }
// v[i] *= mul * cos(y_i), i < 16: y = columns [16 * hf, 16 * hf + 16) of a 32-column y block
// v[i] *= c_i, i < 16: c = the stashed activation derivatives of columns [16 * hf, 16 * hf + 16) of a 32-column block
__device__ __forceinline__ void mul_cos16(float* v, const YBuf& b, int hf) {
    const __half2* h = reinterpret_cast<const __half2*>(&b.q[2 * hf]);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float2 c = __half22float2(h[i]);
        v[2 * i] *= c.x; v[2 * i + 1] *= c.y;
    }
}
__device__ __forceinline__ void mul_cos32(float* v, const YBuf& b) { mul_cos16(v, b, 0); mul_cos16(v + 16, b, 1); }
// The common case: the 16 fp32 accumulator values go to the next dY tile as fp16(v) * c -- one pack and one packed half multiply
// per pair (two roundings of 2^-11 instead of one; the tile is fp16 anyway), stored straight into the swizzled A tile.
__device__ __forceinline__ void store_mul_cos16(uint32_t a_base, int row, int c0, const float* v, const YBuf& b, int hf) {
    const __half2* h = reinterpret_cast<const __half2*>(&b.q[2 * hf]);
    uint32_t o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const __half2 p = __hmul2(__floats2half2_rn(v[2 * i], v[2 * i + 1]), h[i]);
        o[i] = *reinterpret_cast<const uint32_t*>(&p);
    }
    sts128(a_chunk_addr(a_base, row, c0), o[0], o[1], o[2], o[3]);
    sts128(a_chunk_addr(a_base, row, c0 + 8), o[4], o[5], o[6], o[7]);
}

// The chain kernel follows the forward's tile pipeline (tc_pipeline.cuh): CTA pairs (CG = 2) or single CTAs, 4-deep half-stage
// weight ring, two N-chunks per GEMM with the epilogue of chunk 0 under the MMAs of chunk 1, direct stores into released
// K-slabs, early start of the next GEMM on the K-slabs chunk 0 has rewritten.
//
// Per 128-point tile (sat-nerf; s-nerf drops the beta steps), A = the fp16 dY tile in shared memory, D = TMEM accumulator:
//   seed   A[:, :H2]   = g_sun w_s3 cos(s3y)                                                    -> dump d s3y
//   S2     D = A W_s2 ;  A[:, :H2] = D cos(s2y)                                                 -> dump d s2y
//   S1     D = A W_s1 ;  A[:, :H2] = D cos(s1y),  A[:, H2:] = (g_rgb . W_r2) cos(r1y)           -> dump d s1y, d r1y
//   FA     D = A [W_s0[:, :H]; W_r0]          (d feat, stays in TMEM when a beta head follows)
//          beta: A[:, :H2] = g_beta w_b2 cos(b1y)  (+ d t_emb dot products)                     -> dump d b1y
//   FB     D += A W_b0[:, :H] ;  A = D                                                          -> dump d feat
//   A7     D = A W_feats ;  A = (D + g_sigma w_sigma) cos(y_7)                                  -> dump d y_7
//   TRUNK  l = L-1 .. 1:  D = A W_l[:, skip:] ;  A = D w0_{l-1} cos(y_{l-1})                    -> dump d y_{l-1}
// (cos(y) stands for the stashed activation derivative: cos(y), and 30 cos(30 y) for trunk layer 0.)
// Phase probe (dev library, SNB_TC_DBG bit 4096; profiles/dev/chain_probe.sh): clock64 stamps of block 0's second tile, printed by
// the kernel.  Per GEMM: [section entered, cp.async drained, bulk dump drained] barrier passed, accumulator chunk 0 ready,
// chunk 0 stored, accumulator chunk 1 ready, GEMM done.  Findings: profiles/r2_chain_probe.md.
#ifdef SNB_DEV_BUILD
#define CH_MARK(g, k) do { if (probe) stamps[(g) * 8 + (k)] = clock64(); } while (0)
#else
#define CH_MARK(g, k) do { } while (0)
#endif
template <int CG>
__global__ void __launch_bounds__(64 + 32 * kEpiWarpsTrain, 1) tc_chain_kernel(const __grid_constant__ TcBwdArgs A) {
    constexpr int EW = kEpiWarpsTrain, ES = EW / 4, ET = EW * 32;      // epilogue warps, column-block interleave per quadrant, epilogue threads
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const TcProgram& P = A.prog;
    Smem sm = carve(base, P, CG);
    const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
    const int unit = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int n_units = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int stage_bytes = P.stage_bytes / CG;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* T = reinterpret_cast<const float*>(A.packed + P.tables_base);

    if (threadIdx.x == 0) {
        for (int i = 0; i < P.n_stages; ++i) { mbar_init(&sm.full[i], (CG == 2 && cta_rank == 0) ? 2 : 1); mbar_init(&sm.empty[i], 1); }
        mbar_init(sm.acc_full, 1); mbar_init(sm.acc_full2, 1); mbar_init(sm.a_ready, CG); mbar_init(sm.a_ready2, CG);
        for (int i = 0; i < 4; ++i) mbar_init(&sm.slab_free[i], 1);
        fence_barrier_init();
    }
    if (CG == 2) { __syncthreads(); cluster_sync_all(); }
    if (warp == 1) { if (CG == 2) tmem_alloc_2cta(sm.tmem_ptr, 512); else tmem_alloc(sm.tmem_ptr, 512); }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = *sm.tmem_ptr;
    const int tpg = A.tiles_per_group;
    const int n_work = CG == 2 ? (A.n_groups + 1) / 2 : A.n_groups;

    if (warp == 0) {
        pipe_producer<CG>(P, sm, A.packed, stage_bytes, unit, n_units, n_work, tpg, cta_rank, lane);
    } else if (warp == 1) {
        const uint32_t tm = __shfl_sync(0xffffffffu, *sm.tmem_ptr, 0);
        if (CG == 2 && cta_rank == 1) pipe_relay(P, sm, unit, n_units, n_work, tpg, lane);
        else pipe_issuer<CG>(P, sm, stage_bytes, unit, n_units, n_work, tpg, tm);
    } else {
        const int tid_e = threadIdx.x - 64;
        const int quad = warp & 3, half = (warp - 2) >> 2;
        const int row = quad * 32 + lane;
        const uint32_t a_base = smem_u32(sm.a);
        const uint32_t tm_row = tmem + ((uint32_t)(quad * 32) << 16);
        const uint32_t tF = smem_u32(sm.tblF), tV = smem_u32(sm.tblV);
        const int H = P.H, H2 = P.H2, S = A.S, fgsH = H >> 6, fgs2 = H2 >> 6;
        const float scale = loss_scale(*A.absmax), inv_scale = 1.0f / scale;
        float* scratch = sm.z;                                    // per-point tables of the forward are unused here: (4 x 128 x 4) floats
        const uint32_t ready_bar = CG == 2 ? mapa_u32(smem_u32(sm.a_ready), 0) : 0u;
        const uint32_t ready2_bar = CG == 2 ? mapa_u32(smem_u32(sm.a_ready2), 0) : 0u;
        auto signal_ready = [&](int which) {          // one elected arrival per CTA (see tc_field.cu)
            if (tid_e == 0) {
                if (CG == 2 && cta_rank != 0) mbar_arrive_cluster_relaxed(which ? ready2_bar : ready_bar);
                else mbar_arrive(which ? sm.a_ready2 : sm.a_ready);
            }
        };
        int tile_seq = 0;
        for (int wk = unit; wk < n_work; wk += n_units) {
            const int grp = CG == 2 ? 2 * wk + (int)cta_rank : wk;
            const bool live = grp < A.n_groups;                   // false: this CTA only keeps the pair in lockstep
            const int r0 = grp * A.G;
            const int n_rays = live ? min(A.G, A.R - r0) : 0;
            const int Pg = n_rays * S;
            for (int t = 0; t < tpg; ++t, ++tile_seq) {
#ifdef SNB_DEV_BUILD
                const bool probe = (A.dbg & 4096) && blockIdx.x == 0 && tile_seq == 1 && tid_e == 0;
                long long stamps[kMaxGemms * 8 + 8]; const long long t_tile = clock64();
                if (probe) for (int i = 0; i < kMaxGemms * 8 + 8; ++i) stamps[i] = t_tile;
#endif
                const int gt = grp * tpg + t;
                const int p = t * kTile + row;
                const bool valid = p < Pg;
                float g0 = 0.f, g1 = 0.f, g2 = 0.f, gsig = 0.f, gsun = 0.f, gbeta = 0.f;
                if (valid) {
                    const size_t gp = (size_t)r0 * S + p;
                    const float* dh = A.d_head + gp * A.C;
                    g0 = dh[0] * scale; g1 = dh[1] * scale; g2 = dh[2] * scale; gsig = dh[3] * scale; gsun = dh[4] * scale;
                    if (P.has_beta) gbeta = dh[8] * scale;
                }
                if (half == 0 && live) {                          // head-gradient block for the tiny-N weight gradients
                    unsigned char* da = A.bbase + A.bs.dhead;
                    *reinterpret_cast<uint4*>(atom_chunk(da, gt, 1, row, 0)) = make_uint4(pack_half2(g0, g1), pack_half2(g2, gsig), pack_half2(gsun, gbeta), 0u);
#pragma unroll
                    for (int c = 1; c < 8; ++c) *reinterpret_cast<uint4*>(atom_chunk(da, gt, 1, row, c * 8)) = make_uint4(0u, 0u, 0u, 0u);
                }
                // ---- seed: d s3y = g_sun * w_s3 * cos(s3y)  ->  A[:, 0:H2) ----
                CH_MARK(kMaxGemms, 0);
                table_copy<ET>(sm.tblF, T + P.l0_tbl, H2 * 4, tid_e);
                cp_async_wait_all();
                CH_MARK(kMaxGemms, 1);
                if (tid_e == 0) bulk_wait_read();                 // the previous tile's last dump has left shared memory
                CH_MARK(kMaxGemms, 2);
                named_bar_sync(1, ET);
                CH_MARK(kMaxGemms, 3);
                {
                    const uint32_t tok = fresh_token(0x7fffu);
                    for (int n0 = half * 32; n0 < H2; n0 += 32 * ES) {
                        const YBuf yb = live ? yb_load(A.fbase + A.fs.s3y, gt, H2, n0, row) : yb_zero();
                        float v[32];
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            float4 w = lds128(tF + (uint32_t)(n0 + i) * 4u, tok);
                            v[i] = gsun * w.x; v[i + 1] = gsun * w.y; v[i + 2] = gsun * w.z; v[i + 3] = gsun * w.w;
                        }
                        mul_cos32(v, yb);
                        store_act32(a_base, row, n0, v);
                    }
                }
                fence_proxy_async_smem();
                named_bar_sync(1, ET);
                if (tid_e == 0 && live) { bulk_s2g(A.bbase + A.bs.ds3y + (size_t)gt * fgs2 * kSlabBytes, sm.a, (uint32_t)fgs2 * kSlabBytes); bulk_commit(); }
                {   const TcGemm& gn = P.g[0];                   // tables of the first GEMM (none for S2) -- the seed table is dead
                    (void)gn; }
                signal_ready(0);
                signal_ready(1);
                CH_MARK(kMaxGemms, 4);

                int trunk_l = A.n_layers - 1;                     // layer whose dY the next BK_TRUNK GEMM consumes
                bool prev_split = false;                          // the previous GEMM's dump left as two bulk groups
                for (int gi = 0; gi < P.n_gemms; ++gi) {
                    const TcGemm& g = P.g[gi];
                    const int kind = g.kind, N = g.N, n_chunks = g.n_chunks, chunk_n = g.chunk_n;
                    const bool nodrain = kind == BK_FA && P.has_beta;
                    // epilogue tables of this GEMM + make sure earlier dumps have left shared memory
                    if (kind == BK_S1) table_copy<ET>(sm.tblF, T + g.tbl_off, H2 * 16, tid_e);
                    else if (nodrain) { table_copy<ET>(sm.tblF, T + g.tbl_off, H2 * 16, tid_e); table_copy<ET>(sm.tblV, T + g.vec_off, H2 * 4, tid_e); }
                    else if (kind == BK_A7) table_copy<ET>(sm.tblF, T + g.tbl_off, H * 4, tid_e);
                    CH_MARK(gi, 5);
                    cp_async_wait_all();
                    CH_MARK(gi, 6);
                    // Dumps of a two-chunk GEMM leave in two bulk groups (low K-slabs after chunk 0's epilogue, the rest after
                    // chunk 1's), so each group has half a layer to drain before its slabs are rewritten: chunk 0 of this GEMM
                    // rewrites the low slabs (the previous GEMM's HIGH group may still be reading), chunk 1 the high ones.
                    if (tid_e == 0) { if (prev_split) bulk_wait_read1(); else bulk_wait_read(); }
                    CH_MARK(gi, 7);
                    named_bar_sync(1, ET);
                    CH_MARK(gi, 0);
                    const uint32_t tok = fresh_token((uint32_t)gi);
                    // where cos() comes from
                    const unsigned char* yarr = nullptr; int yF = H;
                    if (kind == BK_S2) { yarr = A.fbase + A.fs.s2y; yF = H2; }
                    else if (kind == BK_S1) { yarr = A.fbase + A.fs.s1y; yF = H2; }
                    else if (kind == BK_A7) yarr = A.fbase + A.fs.y[A.n_layers - 1];
                    else if (kind == BK_TRUNK) yarr = A.fbase + A.fs.y[trunk_l - 1];        // (layer 0's entry already holds 30 cos(30 y))
                    if (!live || (SNB_DEV_DBG(A.dbg) & 512)) yarr = nullptr;
                    const bool stores = !nodrain;
                    const bool next_early = gi + 1 < P.n_gemms && n_chunks > 1 && P.g[gi + 1].k_early > 0;
                    bool early_signaled = false, low_dumped = false;
                    float dt0 = 0.f, dt1 = 0.f, dt2 = 0.f, dt3 = 0.f;
                    // destination of the dY tile this GEMM produces (atoms; the tile's K-slabs are contiguous there as in shared memory)
                    unsigned char* dump_dst = A.bbase;
                    if (kind == BK_FB || kind == BK_FA) dump_dst += A.bs.df + (size_t)gt * fgsH * kSlabBytes;
                    else if (kind == BK_A7) dump_dst += A.bs.dy[A.n_layers - 1] + (size_t)gt * fgsH * kSlabBytes;
                    else if (kind == BK_TRUNK) dump_dst += A.bs.dy[trunk_l - 1] + (size_t)gt * fgsH * kSlabBytes;
                    if (tid_e == 0 && live && gi + 1 < P.n_gemms) {
                        // the next GEMM's pre-activation tile (contiguous) on its way into L2 while this one runs
                        const int kn = P.g[gi + 1].kind;
                        const unsigned char* ya = nullptr; uint32_t yb_bytes = (uint32_t)kTile * (uint32_t)H * 2u;
                        if (kn == BK_S1) { ya = A.fbase + A.fs.s1y; yb_bytes >>= 1; bulk_prefetch_l2(A.fbase + A.fs.r1y + (size_t)gt * yb_bytes, yb_bytes); }
                        else if (kn == BK_FA && P.has_beta) { ya = A.fbase + A.fs.b1y; yb_bytes >>= 1; }
                        else if (kn == BK_A7) ya = A.fbase + A.fs.y[A.n_layers - 1];
                        else if (kn == BK_TRUNK) { const int ln = (kind == BK_TRUNK ? trunk_l - 1 : trunk_l) - 1; ya = A.fbase + A.fs.y[ln]; }
                        if (ya) bulk_prefetch_l2(ya + (size_t)gt * yb_bytes, yb_bytes);
                    }
                    // first y block of chunk 0 in flight while the MMAs run
                    YBuf ynext = yb_zero();
                    if (yarr && half * 32 < min(chunk_n, N)) ynext = yb_load(yarr, gt, yF, half * 32, row);
                    for (int ch = 0; ch < n_chunks; ++ch) {
                        if (tid_e == 0) {
                            if (ch == 0) mbar_wait(sm.acc_full, (uint32_t)(tile_seq * P.n_gemms + gi) & 1u, 24);
                            else {
                                mbar_wait(sm.acc_full2, (uint32_t)(tile_seq * P.n_two + g.two_idx) & 1u, 24);
                                if (low_dumped) bulk_wait_read1(); else bulk_wait_read();      // the high slabs' previous dump has left
                            }
                        }
                        named_bar_sync(2, ET);
                        tc_fence_after();
                        CH_MARK(gi, ch == 0 ? 1 : 3);
                        const bool final_chunk = ch == n_chunks - 1;
                        if (!nodrain) {
                            const int n_end = min((ch + 1) * chunk_n, N);
                            int n0 = ch * chunk_n + half * 32;
                            bool have = n0 < n_end;
                            uint64_t* const slab_bar = (stores && !final_chunk) ? sm.slab_free : nullptr;
                            const uint32_t slab_par = (uint32_t)(tile_seq * P.n_store2 + g.store2_idx) & 1u;
                            uint32_t va[16], vb[16];
                            if (ch > 0 && yarr && have) ynext = yb_load(yarr, gt, yF, n0, row);
                            if (have) tmem_ld16(tm_row + (uint32_t)n0, va);
                            while (have) {
                                const YBuf ycur = ynext;
                                const int n1 = n0 + 32 * ES;
                                const bool more = n1 < n_end;
                                if (yarr && more) ynext = yb_load(yarr, gt, yF, n1, row);
                                tmem_ld_wait16(va);
                                tmem_ld16(tm_row + (uint32_t)(n0 + 16), vb);
#pragma unroll
                                for (int hf = 0; hf < 2; ++hf) {
                                    float* v = reinterpret_cast<float*>(hf ? vb : va);
                                    const int c0 = n0 + 16 * hf;
                                    if (hf == 1) { tmem_ld_wait16(vb); if (more) tmem_ld16(tm_row + (uint32_t)n1, va); }
                                    if (kind == BK_A7) {
#pragma unroll
                                        for (int i = 0; i < 16; i += 4) {
                                            float4 w = lds128(tF + (uint32_t)(c0 + i) * 4u, tok);
                                            v[i] = fmaf(gsig, w.x, v[i]); v[i + 1] = fmaf(gsig, w.y, v[i + 1]); v[i + 2] = fmaf(gsig, w.z, v[i + 2]); v[i + 3] = fmaf(gsig, w.w, v[i + 3]);
                                        }
                                    }
                                    if (slab_bar) mbar_wait(slab_bar + (c0 >> 6), slab_par, 28);
                                    if (yarr) { store_mul_cos16(a_base, row, c0, v, ycur, hf); continue; }
                                    if (kind != BK_FA && kind != BK_FB) {
#pragma unroll
                                        for (int i = 0; i < 16; ++i) v[i] = 0.f;       // idle half of a pair: keep the tile finite
                                    }
                                    store_act_cols<16>(a_base, row, c0, v);
                                }
                                if (kind == BK_S1) {
                                    // d r1y = (sum_c g_c W_r2[c][m]) cos(r1y)  ->  A[:, H2 + m)
                                    const YBuf yr = live ? yb_load(A.fbase + A.fs.r1y, gt, H2, n0, row) : yb_zero();
                                    float u[32];
#pragma unroll
                                    for (int i = 0; i < 32; ++i) {
                                        float4 w = lds128(tF + (uint32_t)(n0 + i) * 16u, tok);
                                        u[i] = fmaf(g2, w.z, fmaf(g1, w.y, g0 * w.x));
                                    }
                                    mul_cos32(u, yr);
                                    store_act32(a_base, row, H2 + n0, u);
                                }
                                n0 = n1; have = more;
                            }
                        } else if (final_chunk) {
                            // d feat (so far) stays in TMEM; build d b1y = g_beta * w_b2 * cos(b1y) as the operand of FB
                            for (int n0 = half * 32; n0 < H2; n0 += 32 * ES) {
                                const YBuf yb = live ? yb_load(A.fbase + A.fs.b1y, gt, H2, n0, row) : yb_zero();
                                float v[32];
#pragma unroll
                                for (int i = 0; i < 32; ++i) v[i] = gbeta * lds128(tF + (uint32_t)(n0 + i) * 16u, tok).x;
                                mul_cos32(v, yb);
#pragma unroll
                                for (int i = 0; i < 32; ++i) {
                                    float4 w = lds128(tF + (uint32_t)(n0 + i) * 16u, tok);
                                    dt0 = fmaf(w.y, v[i], dt0); dt1 = fmaf(w.z, v[i], dt1); dt2 = fmaf(w.w, v[i], dt2);
                                }
#pragma unroll
                                for (int i = 0; i < 32; i += 4) {
                                    float4 w = lds128(tV + (uint32_t)(n0 + i) * 4u, tok);
                                    dt3 = fmaf(w.x, v[i], dt3); dt3 = fmaf(w.y, v[i + 1], dt3); dt3 = fmaf(w.z, v[i + 2], dt3); dt3 = fmaf(w.w, v[i + 3], dt3);
                                }
                                store_act32(a_base, row, n0, v);
                            }
                        }
                        if (!final_chunk) {
                            tc_fence_before();
                            CH_MARK(gi, 2);
                            if (next_early) {
                                fence_proxy_async_smem();
                                named_bar_sync(1, ET);
                                signal_ready(0);
                                early_signaled = true;
                                if (stores && chunk_n % 64 == 0) {                 // low K-slabs are final: first bulk group of this GEMM's dump
                                    if (tid_e == 0 && live && !(SNB_DEV_DBG(A.dbg) & 1024)) { bulk_s2g(dump_dst, sm.a, (uint32_t)(chunk_n >> 6) * kSlabBytes); bulk_commit(); }
                                    low_dumped = true;
                                }
                            }
                        }
                    }
                    tc_fence_before();
                    fence_proxy_async_smem();
                    if (nodrain && A.d_t) {
                        float* sc = scratch + (size_t)(half * kTile + row) * 4;
                        sc[0] = dt0; sc[1] = dt1; sc[2] = dt2; sc[3] = dt3;
                    }
                    named_bar_sync(1, ET);              // all TMEM reads / A writes / table reads of this GEMM done
                    CH_MARK(gi, 4);
                    if (nodrain && A.d_t && half == 0 && valid) {
                        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                        for (int h2 = 0; h2 < ES; ++h2) { const float* sc = scratch + (size_t)(h2 * kTile + row) * 4; s0 += sc[0]; s1 += sc[1]; s2 += sc[2]; s3 += sc[3]; }
                        float* o = A.d_t + ((size_t)r0 * S + p) * P.tau;
                        o[0] = s0 * inv_scale; if (P.tau > 1) o[1] = s1 * inv_scale; if (P.tau > 2) o[2] = s2 * inv_scale; if (P.tau > 3) o[3] = s3 * inv_scale;
                    }
                    if (tid_e == 0 && live && !(SNB_DEV_DBG(A.dbg) & 1024)) {
                        // dump the dY tile(s) this step produced (A-tile image = atoms)
                        unsigned char* bb = A.bbase;
                        if (kind == BK_S2) bulk_s2g(bb + A.bs.ds2y + (size_t)gt * fgs2 * kSlabBytes, sm.a, (uint32_t)fgs2 * kSlabBytes);
                        else if (kind == BK_S1) {
                            bulk_s2g(bb + A.bs.ds1y + (size_t)gt * fgs2 * kSlabBytes, sm.a, (uint32_t)fgs2 * kSlabBytes);
                            bulk_s2g(bb + A.bs.dr1y + (size_t)gt * fgs2 * kSlabBytes, sm.a + (size_t)fgs2 * kSlabBytes, (uint32_t)fgs2 * kSlabBytes);
                        } else if (nodrain) bulk_s2g(bb + A.bs.db1y + (size_t)gt * fgs2 * kSlabBytes, sm.a, (uint32_t)fgs2 * kSlabBytes);
                        else {
                            const uint32_t lo = low_dumped ? (uint32_t)(chunk_n >> 6) * kSlabBytes : 0u;
                            bulk_s2g(dump_dst + lo, sm.a + lo, (uint32_t)fgsH * kSlabBytes - lo);
                        }
                        bulk_commit();
                    }
                    prev_split = low_dumped;
                    if (gi + 1 < P.n_gemms) {
                        if (!early_signaled) signal_ready(0);
                        signal_ready(1);
                    }
                    if (kind == BK_TRUNK) --trunk_l;
                }
#ifdef SNB_DEV_BUILD
                if (probe) {
                    printf("chain tile probe (cycles from tile start; kind: [in cp.async-drained dump-drained] top acc0 st0 acc1 done)\n");
                    for (int gi = 0; gi < P.n_gemms; ++gi)
                        printf("  g%02d k%d: [in %7lld cpw %7lld blk %7lld] %7lld %7lld %7lld %7lld %7lld\n", gi, P.g[gi].kind, stamps[gi * 8 + 5] - t_tile,
                               stamps[gi * 8 + 6] - t_tile, stamps[gi * 8 + 7] - t_tile, stamps[gi * 8] - t_tile, stamps[gi * 8 + 1] - t_tile,
                               stamps[gi * 8 + 2] - t_tile, stamps[gi * 8 + 3] - t_tile, stamps[gi * 8 + 4] - t_tile);
                    printf("  seed: entered %lld tables %lld dump-drained %lld barrier %lld signalled %lld\n", stamps[kMaxGemms * 8] - t_tile, stamps[kMaxGemms * 8 + 1] - t_tile,
                           stamps[kMaxGemms * 8 + 2] - t_tile, stamps[kMaxGemms * 8 + 3] - t_tile, stamps[kMaxGemms * 8 + 4] - t_tile);
                }
#endif
            }
        }
        if (tid_e == 0) bulk_wait_read();
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();          // the leader's MMAs read the peer's shared memory: leave together
    if (warp == 1) { if (CG == 2) tmem_dealloc_2cta(tmem, 512); else tmem_dealloc(tmem, 512); }
}

// ---------------------------------------------------------------------------------------------------------------
// finalize: ordered split-K reduction + scatter-add into the flat gradient buffer
// ---------------------------------------------------------------------------------------------------------------
struct DwOut {
    long long part_off, part_stride; int ks, N;      // partial tiles: [ks][256][N] floats
    int kind;                                         // 0 weight block, 1 extra-input block [x sun t 1], 2 tiny head (transposed)
    long long w_off, b_off; int ld, m0, M, col_off, n0, ncols;    // kind 0: G[w_off + (m0+r)*ld + col_off + n0 + c], c < ncols
    int xcol, suncol, tcol, tau;                      // kind 1: destination columns of x / sun / t (-1: none); bias always
    int hc0, nhc;                                     // kind 2: head columns [hc0, hc0+nhc) -> rows of the tiny weight (ld = its n_in)
};

__global__ void dw_finalize_kernel(const DwOut* __restrict__ outs, const float* __restrict__ partial, const float* __restrict__ absmax, float* __restrict__ G) {
    const DwOut o = outs[blockIdx.x];
    const float inv = 1.0f / loss_scale(*absmax);
    // four consecutive columns per thread: one 16-byte load per split-K partial (independent loads in flight), ordered sum
    const int n4 = 64 * o.N;                                       // 256 rows x N / 4
    for (int q = blockIdx.y * blockDim.x + threadIdx.x; q < n4; q += gridDim.y * blockDim.x) {
        const int e = q * 4, r = e / o.N, c0 = e - r * o.N;
        if (o.m0 + r >= o.M) continue;
        if (o.kind == 0 ? c0 >= o.ncols : (o.kind == 1 ? c0 >= 12 : (c0 >= o.hc0 + o.nhc || c0 + 4 <= o.hc0))) continue;
        const float4* src = reinterpret_cast<const float4*>(partial + o.part_off + e);
        const long long stride4 = o.part_stride >> 2;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        int sp = 0;
        for (; sp + 2 <= o.ks; sp += 2) {
            const float4 a0 = src[(long long)sp * stride4], a1 = src[(long long)(sp + 1) * stride4];
            acc.x = (acc.x + a0.x) + a1.x; acc.y = (acc.y + a0.y) + a1.y; acc.z = (acc.z + a0.z) + a1.z; acc.w = (acc.w + a0.w) + a1.w;
        }
        if (sp < o.ks) { const float4 a0 = src[(long long)sp * stride4]; acc.x += a0.x; acc.y += a0.y; acc.z += a0.z; acc.w += a0.w; }
        const float v[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = c0 + i;
            long long dst = -1;
            if (o.kind == 0) { if (c < o.ncols) dst = o.w_off + (long long)(o.m0 + r) * o.ld + o.col_off + o.n0 + c; }
            else if (o.kind == 1) {
                if (c < 3) { if (o.xcol >= 0) dst = o.w_off + (long long)(o.m0 + r) * o.ld + o.xcol + c; }
                else if (c < 6) { if (o.suncol >= 0) dst = o.w_off + (long long)(o.m0 + r) * o.ld + o.suncol + (c - 3); }
                else if (c < 10) { if (o.tcol >= 0 && c - 6 < o.tau) dst = o.w_off + (long long)(o.m0 + r) * o.ld + o.tcol + (c - 6); }
                else if (c == 10) dst = o.b_off + o.m0 + r;
            } else { if (c >= o.hc0 && c < o.hc0 + o.nhc) dst = o.w_off + (long long)(c - o.hc0) * o.ld + o.m0 + r; }
            if (dst >= 0) G[dst] += v[i] * inv;
        }
    }
}

// biases of the N<=3 heads: sums over the rays of the per-ray channel sums (fixed-order tree, double accumulation).
// block c handles channel c of ray_sums (R,16); dst[c] < 0: channel not wanted.
struct BiasDst { long long off[9]; };
__device__ __forceinline__ void head_bias_block(double* sh, int c, const float* __restrict__ ray_sums, int R, const BiasDst& d, float* __restrict__ G) {
    if (d.off[c] < 0) return;
    double acc = 0.0;
    for (int r = threadIdx.x; r < R; r += blockDim.x) acc += (double)ray_sums[(size_t)r * 16 + c];
    sh[threadIdx.x] = acc; __syncthreads();
    for (int s = 128; s; s >>= 1) { if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s]; __syncthreads(); }
    if (threadIdx.x == 0) G[d.off[c]] += (float)sh[0];
}

// sky_color MLP (per ray, satnerf.py:138-143): gradients of sky0 (H2 x 3) and sky2 (3 x H2) from the per-ray sums of the
// sky channels.  One block per hidden unit n; its threads stride over the rays and a fixed-order tree combines them
// (deterministic).
__device__ __forceinline__ void sky_bwd_block(float (*sh)[256], int n, const float* __restrict__ ray_sums, const float* __restrict__ rays, int ray_cols,
                                              const float* __restrict__ aux, int R, int H2, const float* __restrict__ W, long long w0, long long b0, long long w2,
                                              float* __restrict__ G) {
    const float w0x = W[w0 + n * 3], w0y = W[w0 + n * 3 + 1], w0z = W[w0 + n * 3 + 2], bb = W[b0 + n];
    const float v0 = W[w2 + n], v1 = W[w2 + H2 + n], v2 = W[w2 + 2 * H2 + n];
    float acc[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};      // gw2[0..2], gw0[x,y,z], gb0
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        const float d0 = ray_sums[(size_t)r * 16 + 5], d1 = ray_sums[(size_t)r * 16 + 6], d2 = ray_sums[(size_t)r * 16 + 7];
        const float* sd = aux ? aux + (size_t)r * 3 : rays + (size_t)r * ray_cols + 8;
        float pre = fmaf(w0z, sd[2], fmaf(w0y, sd[1], fmaf(w0x, sd[0], bb)));
        float h = fmaxf(pre, 0.f);
        acc[0] = fmaf(d0, h, acc[0]); acc[1] = fmaf(d1, h, acc[1]); acc[2] = fmaf(d2, h, acc[2]);
        float dh_ = pre > 0.f ? fmaf(d2, v2, fmaf(d1, v1, d0 * v0)) : 0.f;
        acc[3] = fmaf(dh_, sd[0], acc[3]); acc[4] = fmaf(dh_, sd[1], acc[4]); acc[5] = fmaf(dh_, sd[2], acc[5]); acc[6] += dh_;
    }
#pragma unroll
    for (int q = 0; q < 7; ++q) sh[q][threadIdx.x] = acc[q];
    __syncthreads();
    for (int st = 128; st; st >>= 1) {
        if ((int)threadIdx.x < st) {
#pragma unroll
            for (int q = 0; q < 7; ++q) sh[q][threadIdx.x] += sh[q][threadIdx.x + st];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        G[w2 + n] += sh[0][0]; G[w2 + H2 + n] += sh[1][0]; G[w2 + 2 * H2 + n] += sh[2][0];
        G[w0 + n * 3] += sh[3][0]; G[w0 + n * 3 + 1] += sh[4][0]; G[w0 + n * 3 + 2] += sh[5][0]; G[b0 + n] += sh[6][0];
    }
}

__device__ __forceinline__ void ray_sum_t_block(int blk, const float* __restrict__ per_point, float* __restrict__ per_ray, int n_rays, int S, int D) {
    int idx = blk * blockDim.x + threadIdx.x;
    if (idx >= n_rays * D) return;
    int r = idx / D, d = idx - r * D;
    float acc = 0.f;
    for (int i = 0; i < S; ++i) acc += per_point[((size_t)r * S + i) * D + d];
    per_ray[idx] = acc;
}

// The small per-ray remainders of the backward in ONE launch (256 threads per block): blocks [0, 9) the biases of the N <= 3 heads,
// [9, 9 + H2) the hidden units of the sky-colour MLP, the rest the per-ray sums of the embedding gradient.
struct TailArgs {
    const float *ray_sums, *rays, *aux, *W, *d_t; float *G, *g_t_emb;
    int R, S, H2, ray_cols, t_dims; long long sky0_w, sky0_b, sky2_w; BiasDst bd;
};
__global__ void __launch_bounds__(256) bwd_tail_kernel(const __grid_constant__ TailArgs A) {
    __shared__ double shd[256];
    __shared__ float shf[7][256];
    const int b = blockIdx.x;
    if (b < 9) head_bias_block(shd, b, A.ray_sums, A.R, A.bd, A.G);
    else if (b < 9 + A.H2) sky_bwd_block(shf, b - 9, A.ray_sums, A.rays, A.ray_cols, A.aux, A.R, A.H2, A.W, A.sky0_w, A.sky0_b, A.sky2_w, A.G);
    else if (A.d_t) ray_sum_t_block(b - 9 - A.H2, A.d_t, A.g_t_emb, A.R, A.S, A.t_dims);
}

// ---------------------------------------------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------------------------------------------
// Small host -> device transfers without staging memory: the bytes ride in the kernel's parameter block.
struct ParamBlob { uint4 q[1984]; };                      // 31 744 bytes (the limit is 32 764 per launch)
__global__ void param_upload_kernel(const __grid_constant__ ParamBlob blob, uint4* __restrict__ dst, int n16) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) dst[i] = blob.q[i];
}
static int upload_by_param(void* dst, const void* src, size_t bytes, cudaStream_t st) {
    static_assert(sizeof(ParamBlob) <= 32000, "parameter block too large");
    for (size_t off = 0; off < bytes; off += sizeof(ParamBlob)) {
        ParamBlob blob;
        const size_t n = bytes - off < sizeof(ParamBlob) ? bytes - off : sizeof(ParamBlob);
        memcpy(&blob, (const char*)src + off, n);
        const int n16 = (int)((n + 15) / 16);            // (destination arrays are padded to 256 bytes by the arena)
        param_upload_kernel<<<(n16 + 255) / 256, 256, 0, st>>>(blob, (uint4*)((char*)dst + off), n16);
        SNB_CHECK_LAUNCH();
    }
    return 0;
}

bool tc_bwd_supported(const FieldLayout& L, const snb_pass_desc* p);
static bool bwd_supported(const FieldLayout& L, const snb_pass_desc* p) {
    if (L.variant == SNB_NERF) return false;
    if (L.width % 128 != 0 || L.width > 512) return false;
    if (L.n_layers + 4 > kMaxGemms || L.n_layers < 2) return false;
    if (p->n_samples > kMaxGroupPts) return false;
    if (L.t_dims > 4) return false;
    return true;
}

bool tc_bwd_supported(const FieldLayout& L, const snb_pass_desc* p) { return bwd_supported(L, p); }

static int group_for(int S) {
    int best = 1; double best_u = 0.0;
    for (int G = 1; G <= kMaxGroupRays; ++G) {
        int pts = G * S; if (pts > kMaxGroupPts) break;
        double u = (double)pts / (double)(((pts + kTile - 1) / kTile) * kTile);
        if (u > best_u + 1e-9) { best_u = u; best = G; }
    }
    return best;
}

static int build_bwd_program(const FieldLayout& L, TcProgram* P, BwdMisc* M) {
    memset(P, 0, sizeof(*P)); memset(M, 0, sizeof(*M));
    const int H = L.width, H2 = H / 2;
    P->H = H; P->H2 = H2; P->tau = L.t_dims; P->has_beta = L.variant == SNB_SATNERF; P->a_slabs = H / 64;
    int ng = 0, tbl = 0;
    auto add = [&](int kind, int N, int K) -> TcGemm& {
        TcGemm& g = P->g[ng++]; memset(&g, 0, sizeof(g));
        g.kind = kind; g.N = N; g.K = K; g.n_chunks = (N + 255) / 256; g.chunk_n = N / g.n_chunks; g.k_slabs = (K + 63) / 64;
        return g;
    };
    P->l0_tbl = tbl; M->t_seed = tbl; tbl += H2;                                      // sun_v_net.6 weight
    { TcGemm& g = add(BK_S2, H2, H2); g.src0 = L.sun[2].w; g.ld0 = H2; g.rows0 = H2; }
    { TcGemm& g = add(BK_S1, H2, H2); g.src0 = L.sun[1].w; g.ld0 = H2; g.rows0 = H2; g.tbl_off = tbl; M->t_r2 = tbl; tbl += 4 * H2; }
    { TcGemm& g = add(BK_FA, H, H); g.src0 = L.sun[0].w; g.ld0 = L.sun[0].n_in; g.rows0 = H2; g.src1 = L.rgb0.w; g.ld1 = L.rgb0.n_in;
      g.tbl_off = tbl; M->t_beta = tbl; tbl += 4 * H2; g.vec_off = tbl; M->t_betav = tbl; tbl += H2; }
    if (P->has_beta) { TcGemm& g = add(BK_FB, H, H2); g.src0 = L.beta0.w; g.ld0 = L.beta0.n_in; g.rows0 = H2; g.accumulate = 1; }
    { TcGemm& g = add(BK_A7, H, H); g.src0 = L.feats.w; g.ld0 = H; g.rows0 = H; g.tbl_off = tbl; M->t_sigma = tbl; tbl += H; }
    for (int l = L.n_layers - 1; l >= 1; --l) {
        TcGemm& g = add(BK_TRUNK, H, H); g.src0 = L.trunk[l].w; g.ld0 = L.trunk[l].n_in; g.col0 = l == L.skip ? L.in_xyz : 0; g.rows0 = H;
    }
    P->n_gemms = ng;
    // pipeline fields (tc_pipeline.cuh): every GEMM rewrites the tile from its accumulator except FA when a beta head follows
    // (d feat stays in TMEM and FB adds onto it); the seed publishes the whole first input tile at once (no early start of S2)
    const bool has_beta = P->has_beta != 0;
    pipe_finish_program(P, [has_beta](const TcGemm& g) { return !(g.kind == BK_FA && has_beta); }, 0);
    if (dev_knobs().no_early) for (int i = 0; i < ng; ++i) P->g[i].k_early = 0;
    long long wbytes = 0; int max_stage = 0;
    for (int i = 0; i < ng; ++i) {
        wbytes += (long long)P->g[i].n_chunks * P->g[i].k_slabs * P->g[i].chunk_n * 128;
        if (P->g[i].chunk_n * 128 > max_stage) max_stage = P->g[i].chunk_n * 128;
    }
    P->stage_bytes = max_stage; P->tables_base = (wbytes + 255) & ~255LL;
    M->s3_w = L.sun[3].w; M->r2_w = L.rgb2.w; M->b2_w = L.beta2.w; M->b0_w = L.beta0.w; M->b0_ld = L.beta0.n_in; M->sigma_w = L.sigma.w;
    M->H = H; M->H2 = H2; M->tau = L.t_dims; M->has_beta = P->has_beta;
    return tbl;
}

static void fwd_stash_layout(const FieldLayout& L, int n_tiles, int tpg, TcStash* S) {      // must mirror tc_field.cu::stash_layout
    memset(S, 0, sizeof(*S));
    const int H = L.width, H2 = H / 2;
    const long long tH = (long long)(H / 64) * kSlabBytes, tH2 = (long long)((H2 + 63) / 64) * kSlabBytes;
    const long long yH = (long long)kTile * H * 2, yH2 = (long long)kTile * H2 * 2;
    long long off = 0;
    auto take = [&](long long per_tile) { long long o = off; off += per_tile * n_tiles; off = (off + 1023) & ~1023LL; return o; };
    for (int l = 0; l < L.n_layers; ++l) { S->a[l] = take(tH); S->y[l] = take(yH); }
    S->feat = take(tH);
    S->r1 = take(tH2); S->s1 = take(tH2); S->s2 = take(tH2); S->s3 = take(tH2); S->b1 = take(tH2);
    S->r1y = take(yH2); S->s1y = take(yH2); S->s2y = take(yH2); S->s3y = take(yH2); S->b1y = take(yH2);
    S->e = take(kSlabBytes);
    S->total = off; S->n_tiles = n_tiles; S->tiles_per_group = tpg;
}

struct BwdPlan {
    int G, groups, tpg, n_tiles, ks, n_outs, n_items;
    TcBwdStash bs;
    size_t off_dhead, off_raysums, off_absmax, off_dt, off_packed, packed_bytes, off_bstash, off_items, off_outs, off_partial, partial_floats, off_sync, sync_ints, total;
};

// Work of the weight-gradient kernel for one pass.  Outputs (DwOut): one per 256-row block of dY features and per N operand
// (the layer's input, <= 512 features; the extra block [x sun t 1]; the head-gradient block for the N <= 3 heads, computed
// transposed with the features as M).  Pieces (DwItem): a "main" piece per input-block output, and the small-N outputs of the
// same arrays grouped three at a time into multi-accumulator pieces, so that every piece of a bundle streams a similar number
// of bytes per stage; all pieces cover the same `ks` point ranges and the pieces of one (bundle, range) are adjacent in the
// list -- piece i runs on pair i % kDwPairs, so they run at the same time, sweep the points in step, and every stash array
// comes from HBM once (the other readers hit L2).  fbase / bbase: forward and backward stash (null in the sizing pass: only
// counts and sizes are used then).
static void dw_work(const FieldLayout& L, const BwdPlan& B, int n_tiles, const unsigned char* fbase, const unsigned char* bbase,
                    std::vector<DwOut>* outs, std::vector<DwItem>* items, size_t* partial_floats, size_t* sync_ints, const TcStash* fs_in = nullptr) {
    const int H = L.width, H2 = H / 2, fgH = H / 64, fg2 = H2 / 64;
    TcStash fs_local; const TcStash* fs = fs_in;
    if (!fs) { fwd_stash_layout(L, n_tiles, B.tpg, &fs_local); fs = &fs_local; }
    struct Opnd { const unsigned char* arr; int fgs, fg0, nfg; DwOut o; };
    struct Piece { int n_acc; Opnd a[kDwMaxAcc]; const unsigned char* b_arr; int b_fgs, b_nfg; };
    std::vector<std::vector<Piece>> bundles;
    const unsigned char* e_arr = fbase + fs->e; const unsigned char* dh_arr = bbase + B.bs.dhead;
    auto blocks = [](int fgs) { return (fgs + 3) / 4; };
    auto block_nfg = [](int fgs, int m) { return fgs - 4 * m < 4 ? fgs - 4 * m : 4; };
    // main pieces of one linear layer: dY (atoms dy_arr, dy_fgs groups) x IN (in_fgs groups, Nin features)
    auto add_main = [&](std::vector<Piece>& bd, const Lin& l, const unsigned char* dy_arr, int dy_fgs, const unsigned char* in_arr, int in_fgs, int Nin, int col_off) {
        for (int m = 0; m < blocks(dy_fgs); ++m) {
            Piece pc; memset(&pc, 0, sizeof(pc)); pc.n_acc = 1; pc.b_arr = in_arr; pc.b_fgs = in_fgs; pc.b_nfg = in_fgs;
            Opnd& a = pc.a[0]; a.arr = dy_arr; a.fgs = dy_fgs; a.fg0 = 4 * m; a.nfg = block_nfg(dy_fgs, m);
            a.o.kind = 0; a.o.w_off = l.w; a.o.ld = l.n_in; a.o.m0 = m * 256; a.o.M = l.n_out; a.o.col_off = col_off; a.o.n0 = 0; a.o.ncols = Nin; a.o.N = in_fgs * 64;
            bd.push_back(pc);
        }
    };
    // small outputs: collected per bundle, then packed three accumulators to a piece
    struct Small { Opnd a; const unsigned char* b_arr; };
    auto add_extra = [&](std::vector<Small>& sm, const Lin& l, const unsigned char* dy_arr, int dy_fgs, int xcol, int suncol, int tcol) {   // dY x [x sun t 1]
        for (int m = 0; m < blocks(dy_fgs); ++m) {
            Small q; memset(&q, 0, sizeof(q)); q.b_arr = e_arr;
            q.a.arr = dy_arr; q.a.fgs = dy_fgs; q.a.fg0 = 4 * m; q.a.nfg = block_nfg(dy_fgs, m);
            q.a.o.kind = 1; q.a.o.w_off = l.w; q.a.o.b_off = l.b; q.a.o.ld = l.n_in; q.a.o.m0 = m * 256; q.a.o.M = l.n_out;
            q.a.o.xcol = xcol; q.a.o.suncol = suncol; q.a.o.tcol = tcol; q.a.o.tau = L.t_dims; q.a.o.N = 64;
            sm.push_back(q);
        }
    };
    auto add_tiny = [&](std::vector<Small>& sm, const Lin& l, const unsigned char* in_arr, int in_fgs, int hc0, int nhc) {     // W (nhc x n_in): rows = head columns
        for (int m = 0; m < blocks(in_fgs); ++m) {
            Small q; memset(&q, 0, sizeof(q)); q.b_arr = dh_arr;
            q.a.arr = in_arr; q.a.fgs = in_fgs; q.a.fg0 = 4 * m; q.a.nfg = block_nfg(in_fgs, m);
            q.a.o.kind = 2; q.a.o.w_off = l.w; q.a.o.ld = l.n_in; q.a.o.m0 = m * 256; q.a.o.M = l.n_in; q.a.o.hc0 = hc0; q.a.o.nhc = nhc; q.a.o.N = 64;
            sm.push_back(q);
        }
    };
    auto pack_small = [&](std::vector<Piece>& bd, const std::vector<Small>& sm) {
        for (size_t i = 0; i < sm.size();) {
            Piece pc; memset(&pc, 0, sizeof(pc)); pc.b_arr = sm[i].b_arr; pc.b_fgs = 1; pc.b_nfg = 1;
            while (i < sm.size() && pc.n_acc < kDwMaxAcc && sm[i].b_arr == pc.b_arr) pc.a[pc.n_acc++] = sm[i++].a;
            bd.push_back(pc);
        }
    };
    const unsigned char* fb = fbase; const unsigned char* bb = bbase;
    for (int l = L.n_layers - 1; l >= 1; --l) {
        std::vector<Piece> bd; std::vector<Small> sm;
        add_main(bd, L.trunk[l], bb + B.bs.dy[l], fgH, fb + fs->a[l - 1], fgH, H, l == L.skip ? L.in_xyz : 0);
        add_extra(sm, L.trunk[l], bb + B.bs.dy[l], fgH, l == L.skip ? 0 : -1, -1, -1);
        pack_small(bd, sm); bundles.push_back(bd);
    }
    {   std::vector<Piece> bd; std::vector<Small> sm;                                   // feats + sigma head (both read a_{L-1}) + trunk layer 0 (no input block)
        add_main(bd, L.feats, bb + B.bs.df, fgH, fb + fs->a[L.n_layers - 1], fgH, H, 0);
        add_extra(sm, L.feats, bb + B.bs.df, fgH, -1, -1, -1);
        pack_small(bd, sm); sm.clear();
        add_tiny(sm, L.sigma, fb + fs->a[L.n_layers - 1], fgH, 3, 1);
        pack_small(bd, sm); sm.clear();
        add_extra(sm, L.trunk[0], bb + B.bs.dy[0], fgH, 0, -1, -1);
        pack_small(bd, sm); bundles.push_back(bd); }
    {   std::vector<Piece> bd; std::vector<Small> sm;                                   // first head layers: all read feat
        add_main(bd, L.rgb0, bb + B.bs.dr1y, fg2, fb + fs->feat, fgH, H, 0);
        add_main(bd, L.sun[0], bb + B.bs.ds1y, fg2, fb + fs->feat, fgH, H, 0);
        if (L.variant == SNB_SATNERF) add_main(bd, L.beta0, bb + B.bs.db1y, fg2, fb + fs->feat, fgH, H, 0);
        add_extra(sm, L.rgb0, bb + B.bs.dr1y, fg2, -1, -1, -1);
        add_extra(sm, L.sun[0], bb + B.bs.ds1y, fg2, -1, H, -1);
        if (L.variant == SNB_SATNERF) add_extra(sm, L.beta0, bb + B.bs.db1y, fg2, -1, -1, H);
        pack_small(bd, sm); bundles.push_back(bd); }
    {   std::vector<Piece> bd; std::vector<Small> sm;                                   // sun_v_net.2 / .4 and the N <= 3 heads of the H/2-wide activations
        add_main(bd, L.sun[1], bb + B.bs.ds2y, fg2, fb + fs->s1, fg2, H2, 0);
        add_main(bd, L.sun[2], bb + B.bs.ds3y, fg2, fb + fs->s2, fg2, H2, 0);
        add_extra(sm, L.sun[1], bb + B.bs.ds2y, fg2, -1, -1, -1);
        add_extra(sm, L.sun[2], bb + B.bs.ds3y, fg2, -1, -1, -1);
        pack_small(bd, sm); sm.clear();
        add_tiny(sm, L.rgb2, fb + fs->r1, fg2, 0, 3);
        add_tiny(sm, L.sun[3], fb + fs->s3, fg2, 4, 1);
        if (L.variant == SNB_SATNERF) add_tiny(sm, L.beta2, fb + fs->b1, fg2, 5, 1);
        pack_small(bd, sm); bundles.push_back(bd); }
    // split-K: the same ranges for every piece; ks = the count (<= 8) that fills whole waves of kDwPairs pieces best
    int per_range = 0;
    for (const auto& bd : bundles) per_range += (int)bd.size();
    int ks = 1; double best = -1.0;
    for (int k = 1; k <= 8 && k <= n_tiles; ++k) {
        const int n = per_range * k, waves = (n + kDwPairs - 1) / kDwPairs;
        double eff = (double)n / ((double)waves * kDwPairs);
        if (k == 1 && n_tiles >= 16) eff *= 0.5;                  // a single range leaves most pairs idle on a real batch
        if (eff > best + 1e-9) { best = eff; ks = k; }
    }
    outs->clear(); items->clear();
    long long off = 0;
    // outputs first (their partial offsets), then the pieces range by range, bundle by bundle
    std::vector<std::vector<std::vector<int>>> out_idx(bundles.size());
    for (size_t bi = 0; bi < bundles.size(); ++bi) {
        out_idx[bi].resize(bundles[bi].size());
        for (size_t pi = 0; pi < bundles[bi].size(); ++pi)
            for (int a = 0; a < bundles[bi][pi].n_acc; ++a) {
                DwOut o = bundles[bi][pi].a[a].o; o.ks = ks; o.part_off = off; o.part_stride = 256LL * o.N;
                off += (long long)ks * o.part_stride;
                out_idx[bi][pi].push_back((int)outs->size());
                outs->push_back(o);
            }
    }
    const int max_stages = 2 * ((n_tiles + ks - 1) / ks), sync_per_group = (max_stages + kDwSyncEvery - 1) / kDwSyncEvery;
    *sync_ints = (size_t)ks * bundles.size() * sync_per_group;
    for (int sp = 0; sp < ks; ++sp)
        for (size_t bi = 0; bi < bundles.size(); ++bi)
            for (size_t pi = 0; pi < bundles[bi].size(); ++pi) {
                const Piece& pc = bundles[bi][pi];
                DwItem w; memset(&w, 0, sizeof(w));
                w.n_acc = pc.n_acc; w.b_off = (long long)(uintptr_t)pc.b_arr; w.b_fgs = pc.b_fgs; w.b_fg0 = 0; w.b_nfg = pc.b_nfg;
                w.k_tile0 = (int)((long long)n_tiles * sp / ks); w.k_tiles = (int)((long long)n_tiles * (sp + 1) / ks) - w.k_tile0;
                w.sync_off = (int)((sp * bundles.size() + bi) * sync_per_group); w.sync_n = (int)bundles[bi].size();
                for (int a = 0; a < pc.n_acc; ++a) {
                    w.a_off[a] = (long long)(uintptr_t)pc.a[a].arr; w.a_fgs[a] = pc.a[a].fgs; w.a_fg0[a] = pc.a[a].fg0; w.a_nfg[a] = pc.a[a].nfg;
                    const DwOut& o = (*outs)[out_idx[bi][pi][a]];
                    w.out_off[a] = o.part_off + (long long)sp * o.part_stride;
                }
                items->push_back(w);
            }
    *partial_floats = (size_t)off;
}

static void plan_bwd(const FieldLayout& L, const snb_pass_desc* p, const TcProgram& P, int n_tbl_floats, BwdPlan* B) {
    const int H = L.width, H2 = H / 2, S = p->n_samples;
    B->G = group_for(S); B->groups = (p->n_rays + B->G - 1) / B->G; B->tpg = (B->G * S + kTile - 1) / kTile; B->n_tiles = B->groups * B->tpg;
    const long long tH = (long long)(H / 64) * kSlabBytes, tH2 = (long long)(H2 / 64) * kSlabBytes;
    long long off = 0;
    auto take = [&](long long per_tile) { long long o = off; off += per_tile * B->n_tiles; off = (off + 1023) & ~1023LL; return o; };
    for (int l = 0; l < L.n_layers; ++l) B->bs.dy[l] = take(tH);
    B->bs.df = take(tH); B->bs.dr1y = take(tH2); B->bs.ds1y = take(tH2); B->bs.ds2y = take(tH2); B->bs.ds3y = take(tH2); B->bs.db1y = take(tH2);
    B->bs.dhead = take(kSlabBytes);
    B->bs.total = off;
    // weight-gradient GEMMs: outputs (256-row blocks), split-K pieces and the partial buffer
    std::vector<DwOut> outs; std::vector<DwItem> items;
    dw_work(L, *B, B->n_tiles, nullptr, nullptr, &outs, &items, &B->partial_floats, &B->sync_ints);
    B->n_outs = (int)outs.size(); B->n_items = (int)items.size(); B->ks = 0;
    Arena ar(nullptr, 0);
    B->off_dhead = ar.off; ar.take<float>((size_t)p->n_rays * S * L.n_channels);
    B->off_raysums = ar.off; ar.take<float>((size_t)p->n_rays * 16);
    B->off_absmax = ar.off; ar.take<float>(64);
    B->off_dt = ar.off; ar.take<float>((size_t)p->n_rays * S * (L.t_dims > 0 ? L.t_dims : 1));
    B->packed_bytes = (size_t)P.tables_base + (size_t)n_tbl_floats * 4 + 256;
    B->off_packed = ar.off; ar.take<unsigned char>(B->packed_bytes);
    B->off_bstash = ar.off; ar.take<unsigned char>((size_t)B->bs.total + 1024);
    B->off_items = ar.off; ar.take<DwItem>(B->n_items);
    B->off_outs = ar.off; ar.take<DwOut>(B->n_outs);
    B->off_partial = ar.off; ar.take<float>(B->partial_floats);
    B->off_sync = ar.off; ar.take<int>(B->sync_ints);
    B->total = ar.off;
}

int tc_bwd_workspace(const FieldLayout& L, const snb_pass_desc* p, size_t* bytes) {
    *bytes = 0;
    if (!bwd_supported(L, p)) return 0;
    TcProgram P; BwdMisc M; int nt = build_bwd_program(L, &P, &M);
    BwdPlan B; memset(&B, 0, sizeof(B)); plan_bwd(L, p, P, nt, &B);
    *bytes = B.total + 4096;
    return 0;
}

int tc_render_backward(const FieldLayout& L, const snb_pass_desc* p, const snb_render_io* io, const snb_render_grads* g,
                       void* workspace, size_t workspace_bytes, cudaStream_t st) {
    if (!bwd_supported(L, p) || !io->stash) return 1;
    if (dev_knobs().bwd_off) return 1;
    static int sm_count = 0, max_smem = 0;
    if (!sm_count) {
        int dev = 0; SNB_CUDA(cudaGetDevice(&dev));
        SNB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        SNB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    const int H = L.width, H2 = H / 2, S = p->n_samples, C = L.n_channels, R = p->n_rays;
    const long long Ptot = (long long)R * S;
    TcBwdArgs A; memset(&A, 0, sizeof(A));
    BwdMisc M; int nt = build_bwd_program(L, &A.prog, &M);
    TcProgram& P = A.prog;
    BwdPlan B; memset(&B, 0, sizeof(B)); plan_bwd(L, p, P, nt, &B);
    if (B.total > workspace_bytes) SNB_FAIL(-4, "tensor-core backward: workspace too small (%zu < %zu)", workspace_bytes, B.total);
    unsigned char* ws = (unsigned char*)workspace;
    float* d_head = (float*)(ws + B.off_dhead);
    float* absmax = (float*)(ws + B.off_absmax);
    float* d_t = (float*)(ws + B.off_dt);

    // 1. compositing backward -> d_head
    CompositeBwdArgs b{};
    b.R = R; b.S = S; b.C = C; b.z = io->z_vals; b.noise = p->noise_std != 0.f ? io->noise : nullptr; b.noise_std = p->noise_std;
    b.weights = io->weights; b.transparency = io->transparency; b.sigma = io->sigma; b.albedo = io->albedo; b.sun = io->sun;
    b.sky = io->sky; b.beta = io->beta; b.nerf_rgb = io->nerf_rgb;
    b.g_rgb = g->g_rgb; b.g_depth = g->g_depth; b.g_weights = g->g_weights; b.g_transparency = g->g_transparency;
    b.g_albedo = g->g_albedo; b.g_sun = g->g_sun; b.g_sky = g->g_sky; b.g_beta = g->g_beta; b.d_head = d_head;
    SNB_TRY(fill_loss(b, g->loss));
    float* ray_sums = (float*)(ws + B.off_raysums);
    SNB_CUDA(cudaMemsetAsync(absmax, 0, 256, st));
    b.absmax = absmax;                                   // max |d_head| is taken while d_head is written
    SNB_TRY(launch_composite_bwd_warp(b, ray_sums, st));
    (void)Ptot;

    // 2. packed transposed weights + tables, then the input-gradient chain
    A.packed = ws + B.off_packed;
    tc_bwd_pack_kernel<<<dim3(64, P.n_gemms), 256, 0, st>>>(P, io->params, A.packed);
    SNB_CHECK_LAUNCH();
    tc_bwd_tables_kernel<<<4, 256, 0, st>>>(M, P.tables_base, io->params, A.packed);
    SNB_CHECK_LAUNCH();
    // CTA pairs as in the forward (each CTA streams and buffers half of every transposed weight tile: 4 x 16 KB ring)
    int cg = B.groups >= 2 ? 2 : 1;
    if (p->flags & SNB_PASS_SINGLE_CTA) cg = 1;
    if (dev_knobs().cg == 1 || dev_knobs().cg == 2) cg = dev_knobs().cg;
    size_t fixed = (size_t)P.a_slabs * kSlabBytes + (kTblF + kTblV + 8 * kMaxGroupPts * 4 + 2 * kMaxGroupRays * 256 * 4 + 768) + 1024;
    int ns = (int)(((size_t)max_smem - fixed) / (P.stage_bytes / cg)); if (ns > 8) ns = 8;
    if (ns < 2) SNB_FAIL(-6, "tensor-core backward: not enough shared memory for the weight ring");
    P.n_stages = ns;
    size_t smem = fixed + (size_t)ns * (P.stage_bytes / cg);
    fwd_stash_layout(L, B.n_tiles, B.tpg, &A.fs); A.fbase = (const unsigned char*)io->stash;
    A.bs = B.bs; A.bbase = ws + B.off_bstash;
    A.d_head = d_head; A.C = C; A.absmax = absmax; A.d_t = (L.t_dims && g->g_t_emb) ? d_t : nullptr;
    A.n_layers = L.n_layers; A.R = R; A.S = S; A.G = B.G; A.n_groups = B.groups; A.tiles_per_group = B.tpg;
    A.dbg = dev_knobs().dbg;
    if (cg == 2) {
        SNB_CUDA(cudaFuncSetAttribute(tc_chain_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int n_pairs = (B.groups + 1) / 2, max_pairs = sm_count / 2;
        cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(2 * (n_pairs < max_pairs ? n_pairs : max_pairs)); cfg.blockDim = dim3(64 + 32 * kEpiWarpsTrain);
        cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        SNB_CUDA(cudaLaunchKernelEx(&cfg, tc_chain_kernel<2>, A));
        ++g_launches;
    } else {
        SNB_CUDA(cudaFuncSetAttribute(tc_chain_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        tc_chain_kernel<1><<<B.groups < sm_count ? B.groups : sm_count, 64 + 32 * kEpiWarpsTrain, smem, st>>>(A);
        SNB_CHECK_LAUNCH();
    }

    // 3. weight-gradient GEMMs: balanced per-pair work lists
    std::vector<DwOut> outs; std::vector<DwItem> items; size_t partial_floats = 0;
    size_t sync_ints = 0;
    dw_work(L, B, B.n_tiles, (const unsigned char*)io->stash, ws + B.off_bstash, &outs, &items, &partial_floats, &sync_ints, &A.fs);
    if ((int)items.size() != B.n_items || (int)outs.size() != B.n_outs || partial_floats != B.partial_floats)
        SNB_FAIL(-3, "internal: weight-gradient work list mismatch (%zu/%d, %zu/%d)", items.size(), B.n_items, outs.size(), B.n_outs);
    DwItem* d_items = (DwItem*)(ws + B.off_items); DwOut* d_outs = (DwOut*)(ws + B.off_outs);
    // The work lists (~25 KB) travel to the device as KERNEL PARAMETERS (up to 32 KB per launch since CUDA 12.1): no pinned
    // staging buffer, no event, no allocation -- the library keeps no state between calls.
    SNB_TRY(upload_by_param(d_items, items.data(), sizeof(DwItem) * items.size(), st));
    SNB_TRY(upload_by_param(d_outs, outs.data(), sizeof(DwOut) * outs.size(), st));
    float* partial = (float*)(ws + B.off_partial);
    int* d_sync = (int*)(ws + B.off_sync);
    SNB_CUDA(cudaMemsetAsync(d_sync, 0, B.sync_ints * sizeof(int), st));
    SNB_TRY(launch_dw(d_items, (int)items.size(), nullptr, partial, d_sync, st));

    // 4. reductions / scatter into the flat gradient
    dw_finalize_kernel<<<dim3((unsigned)outs.size(), 32), 256, 0, st>>>(d_outs, partial, absmax, g->g_params);
    SNB_CHECK_LAUNCH();
    BiasDst bd; for (int c = 0; c < 9; ++c) bd.off[c] = -1;
    bd.off[0] = L.rgb2.b; bd.off[1] = L.rgb2.b + 1; bd.off[2] = L.rgb2.b + 2; bd.off[3] = L.sigma.b; bd.off[4] = L.sun[3].b;
    bd.off[5] = L.sky2.b; bd.off[6] = L.sky2.b + 1; bd.off[7] = L.sky2.b + 2;
    if (L.variant == SNB_SATNERF) bd.off[8] = L.beta2.b;
    TailArgs ta; memset(&ta, 0, sizeof(ta));
    ta.ray_sums = ray_sums; ta.rays = io->rays; ta.aux = io->aux_dir; ta.W = io->params; ta.G = g->g_params; ta.R = R; ta.S = S; ta.H2 = H2;
    ta.ray_cols = p->ray_cols; ta.sky0_w = L.sky0.w; ta.sky0_b = L.sky0.b; ta.sky2_w = L.sky2.w; ta.bd = bd;
    ta.d_t = A.d_t; ta.g_t_emb = g->g_t_emb; ta.t_dims = L.t_dims;
    const int t_blocks = A.d_t ? (R * L.t_dims + 255) / 256 : 0;
    bwd_tail_kernel<<<9 + H2 + t_blocks, 256, 0, st>>>(ta);
    SNB_CHECK_LAUNCH();
    return 0;
}

}  // namespace snb
