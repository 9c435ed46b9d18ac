// Fused tensor-core render pass for sm_100a (SNB_FP16_TC): one persistent, warp-specialised kernel
// evaluates the whole field (SatNeRF / ShadowNeRF forward, models/satnerf.py:156-208) on 128-point
// tiles and alpha-composites the rays of the tile group (satnerf.py:43-78) without activations ever
// leaving the SM.  CTAs run as pairs (cta_group::2; template CG = 1: single CTAs): each CTA owns one
// tile, the pair's leader issues M = 256 MMAs, each CTA streams and buffers half of every weight tile.
//
//   warp 0      weight producer: streams this CTA's half of the pre-swizzled fp16 weight tiles
//               (L2-resident, ~5 MB) into a 4-deep shared-memory ring with 1-D bulk async copies
//               (TMA engine) + mbarrier transactions
//   warp 1      leader: MMA issuer -- tcgen05.mma (M = 2 x 128 points, N <= 256 features, K = 16), the
//               activation tile as the K-major A operand in shared memory, accumulators in TMEM (all 512
//               columns), two weight stages (8 MMAs) per elected region;  peer: relay that reports
//               "my half of the stage has landed" to the leader's stage barrier
//   warps 2-17  epilogue: sample positions and the K=3 first layer in fp32 on CUDA cores, then per
//               layer TMEM -> registers (16-column halves, next load in flight), sin, fp16 pack into the
//               swizzled A tile of the next layer; biases and the skip layer's xyz term arrive inside
//               the accumulator through one extra K = 16 step on 32B-swizzled aux operand tiles; the
//               tiny output heads (sigma, rgb, sun, beta) are dot products folded into the epilogue of
//               the layer that feeds them; finally a warp-scan transmittance product composites each ray.
//
// Layer pipeline: N in two 256-column chunks with one accumulator barrier each; chunk 0's
// epilogue runs under chunk 1's MMAs and stores into the K-slabs chunk 1 has released (slab_free); the
// next layer's MMAs on those slabs are queued as soon as chunk 0 is stored (first ready signal), the
// rest after chunk 1's epilogue (second ready signal).
//
// Arithmetic: fp16 operands / fp32 accumulation for the h x h contractions; first layer, per-ray sun /
// embedding terms, heads and compositing in fp32; bias / skip terms as fp16 hi+lo pairs (exact to 2^-22).
#include "tc_field.cuh"
#include "tc_common.cuh"

namespace snb {

// --------------------------------------------------------------------------------------------------
// host: build the per-tile GEMM program
// --------------------------------------------------------------------------------------------------
static bool tc_supported(const FieldLayout& L, const snb_pass_desc* p) {
    if (L.variant == SNB_NERF && (L.in_xyz > 64 || L.in_xyz % 6 != 0 || L.in_dir > 48 || L.in_dir % 6 != 0 || L.skip < 1)) return false;   // (Mapping off / > 10 frequencies: fp32 path)
    if (L.width % 64 != 0 || L.width > 512 || L.width < 64) return false;
    if (L.n_layers + 4 > kMaxGemms) return false;
    if (p->n_samples > kMaxGroupPts) return false;
    if (L.t_dims > 32) return false;
    return true;
}

static int build_program(const FieldLayout& L, TcProgram* P, bool no_beta = false, bool sigma_only = false) {
    memset(P, 0, sizeof(*P));
    const int H = L.width, H2 = H / 2;
    P->H = H; P->H2 = H2; P->tau = L.t_dims; P->has_beta = L.variant == SNB_SATNERF && !no_beta;
    P->a_slabs = H / 64;
    int ng = 0; int tbl = 0;
    auto add = [&](int kind, int N, int K) -> TcGemm& {
        TcGemm& g = P->g[ng++]; memset(&g, 0, sizeof(g));
        g.kind = kind; g.N = N; g.K = K; g.n_chunks = (N + 255) / 256; g.chunk_n = N / g.n_chunks; g.k_slabs = (K + 63) / 64;   // 128-wide chunks measured slower
        return g;
    };
    auto tables = [&](TcGemm& g, int fmt, int vec) {
        g.fmt = fmt; g.has_vec = vec; g.tbl_off = tbl; tbl += g.N * (fmt == TF_F4 ? 4 : (fmt == TF_F1 ? 1 : 0));
        g.vec_off = tbl; if (vec) tbl += g.N;
    };
    P->nerf = L.variant == SNB_NERF;
    if (P->nerf) {
        // models/nerf.py:184-227: every trunk layer is a GEMM (the encoded position is 60 wide: one extra K-slab of the A tile that
        // layer 0 and the skip layer read), ReLU epilogues, biases on the aux K-step; sigma head folded into the last trunk layer;
        // feats; rgb_from_xyzdir.0 with the encoded view direction as a per-ray bias
        P->pe_xyz = L.in_xyz / 6; P->pe_dir = L.in_dir / 6; P->pe_slab = H / 64; P->a_slabs = H / 64 + 1;
        P->l0_tbl = tbl;
        for (int i = 0; i < L.n_layers; ++i) {
            const bool skip = i == L.skip;
            TcGemm& g = add(GK_TRUNK, H, i == 0 ? 64 : (skip ? H + 64 : H));
            g.relu = 1; g.aux = 1; g.last = i == L.n_layers - 1; g.skip = 0;
            g.src0 = L.trunk[i].w; g.ld0 = L.trunk[i].n_in; g.rows0 = H;
            if (i == 0) { g.a_slab0 = P->pe_slab; g.col0 = 0; g.kvalid = L.in_xyz; }
            else if (skip) { g.col0 = L.in_xyz; g.kvalid = H; g.k2_start = H; g.k2_cols = L.in_xyz; g.col_k2 = 0; }
            else g.kvalid = H;
            tables(g, TF_NONE, g.last);
        }
        { TcGemm& g = add(GK_FEAT, H, H); g.src0 = L.feats.w; g.ld0 = H; g.rows0 = H; g.aux = 1; g.kvalid = H; tables(g, TF_NONE, 0); }
        { TcGemm& g = add(GK_HEADN, H2, H); g.src0 = L.rgb0.w; g.ld0 = L.rgb0.n_in; g.rows0 = H2; g.kvalid = H; tables(g, TF_F4, 0); }
    } else {
    P->l0_tbl = tbl; tbl += H * 4;
    P->l0_w = L.trunk[0].w; P->l0_b = L.trunk[0].b;
    for (int i = 1; i < L.n_layers; ++i) {
        TcGemm& g = add(GK_TRUNK, H, H);
        g.skip = i == L.skip; g.last = i == L.n_layers - 1;
        g.src0 = L.trunk[i].w; g.ld0 = L.trunk[i].n_in; g.col0 = g.skip ? L.in_xyz : 0; g.rows0 = H;
        g.aux = g.skip ? 2 : 1;                  // bias (+ xyz term of the skip layer) ride on the aux K-step: no epilogue table
        tables(g, TF_NONE, g.last);
    }
    if (!sigma_only) {
    { TcGemm& g = add(GK_FEAT, H, H); g.src0 = L.feats.w; g.ld0 = H; g.rows0 = H; g.aux = 1; tables(g, TF_NONE, 0); }
    { int nb = P->has_beta ? H2 : 0;
      TcGemm& g = add(GK_HEADA, nb + H2, H);
      g.src0 = P->has_beta ? L.beta0.w : L.rgb0.w; g.ld0 = P->has_beta ? L.beta0.n_in : L.rgb0.n_in; g.rows0 = P->has_beta ? nb : H2;
      g.src1 = L.rgb0.w; g.ld1 = L.rgb0.n_in;
      tables(g, TF_F4, P->has_beta ? 1 : 0); }       // extra vector: beta_from_xyz.2 weights (one float per column of the beta half)
    { TcGemm& g = add(GK_SUN1, H2, H); g.src0 = L.sun[0].w; g.ld0 = L.sun[0].n_in; g.rows0 = H2; tables(g, TF_NONE, 0); }
    { TcGemm& g = add(GK_SUN2, H2, H2); g.src0 = L.sun[1].w; g.ld0 = H2; g.rows0 = H2; tables(g, TF_F1, 0); }
    { TcGemm& g = add(GK_SUN3, H2, H2); g.src0 = L.sun[2].w; g.ld0 = H2; g.rows0 = H2; tables(g, TF_F1, 1); }
    }      // (sigma_only: the program ends with the last trunk layer, whose epilogue carries the density head)
    }
    for (int i = 0; i < ng; ++i) if (!P->g[i].kvalid) P->g[i].kvalid = P->g[i].K;
    P->n_gemms = ng;
    P->n_two = 0; P->n_store2 = 0;
    for (int i = 0; i < ng; ++i) {
        TcGemm& g = P->g[i];
        g.two_idx = P->n_two; if (g.n_chunks == 2) ++P->n_two;
        // chunk 0 of a two-chunk GEMM that writes the tile stores its results straight into the K-slabs the MMAs of chunk 1
        // have finished with (released slab by slab through slab_free[])
        const bool st = g.kind == GK_TRUNK || g.kind == GK_FEAT;
        g.store2_idx = P->n_store2; g.free_slabs = 0;
        // (a GEMM that does not read the low K-slabs -- nerf layer 0 -- has nothing to wait for: no slab_free phases)
        if (st && g.n_chunks == 2 && g.a_slab0 == 0) { g.free_slabs = (g.chunk_n + 63) / 64; ++P->n_store2; }
    }
    // Early start of a GEMM's first K-slabs (see the kernel): after layer 0 / a two-chunk producer the low half of the
    // input tile is published before the high half; HEADA leaves the tile untouched, so SUN1 may start on all of it.
    P->g[0].k_early = (!P->nerf && H % 128 == 0 && H >= 256) ? H / 128 : 0;
    for (int i = 1; i < ng; ++i) {
        const TcGemm& pr = P->g[i - 1];
        const bool pr_stores = pr.kind == GK_TRUNK || pr.kind == GK_FEAT || pr.kind == GK_SUN1 || pr.kind == GK_SUN2;
        if (pr.n_chunks == 2 && pr_stores && pr.chunk_n % 64 == 0) P->g[i].k_early = pr.chunk_n / 64;
        else if (pr.n_chunks == 2 && pr.kind == GK_HEADA) P->g[i].k_early = P->g[i].k_slabs - 1;
        else P->g[i].k_early = 0;
        // At least one stage of every GEMM waits for the second signal of BOTH CTAs of a pair: otherwise a fast CTA could
        // send the second signal of the following GEMM before its partner has sent this one's (barrier phase aliasing).
#ifdef SNB_V_SUN1_ALL
        if (pr.n_chunks == 2 && pr.kind == GK_HEADA) P->g[i].k_early = P->g[i].k_slabs; else
#endif
        if (P->g[i].k_early > P->g[i].k_slabs - 1) P->g[i].k_early = P->g[i].k_slabs - 1;
    }
    if (dev_knobs().no_early) for (int i = 0; i < ng; ++i) P->g[i].k_early = 0;
    P->consts = tbl; tbl += 8;
    P->sunw = tbl; tbl += (P->nerf ? L.in_dir + 1 : 4) * H2;      // [3][H2] weights + [H2] bias (nerf: [in_dir][H2] view-direction columns of rgb_from_xyzdir.0 + bias)
    P->betaw = tbl; tbl += (L.t_dims + 1) * H2;   // [tau][H2] weights + [H2] bias
    P->sky = tbl; tbl += 4 * H2 + 3 * H2 + 4;     // sky0 [H2][3]+b[H2] as [H2][4]; sky2 [3][H2]; b2[3]
    long long wbytes = 0; int max_stage = 0;
    for (int i = 0; i < ng; ++i) {
        wbytes += gemm_stream_bytes(P->g[i]);
        if (P->g[i].chunk_n * 128 > max_stage) max_stage = P->g[i].chunk_n * 128;
    }
    P->stage_bytes = max_stage;
    P->tables_base = (wbytes + 255) & ~255LL;
    return tbl;      // number of floats in the table area
}

// Stash layout for n_tiles tiles (see TcStash).
static void stash_layout(const FieldLayout& L, int n_tiles, int tiles_per_group, TcStash* S) {
    memset(S, 0, sizeof(*S));
    const int H = L.width, H2 = H / 2;
    const long long tH = (long long)(H / 64) * kSlabBytes, tH2 = (long long)((H2 + 63) / 64) * kSlabBytes;   // atoms bytes per tile
    const long long yH = (long long)kTile * H * 2, yH2 = (long long)kTile * H2 * 2;                           // yb bytes per tile
    long long off = 0;
    auto take = [&](long long per_tile) { long long o = off; off += per_tile * n_tiles; off = (off + 1023) & ~1023LL; return o; };
    for (int l = 0; l < L.n_layers; ++l) { S->a[l] = take(tH); S->y[l] = take(yH); }
    S->feat = take(tH);
    S->r1 = take(tH2); S->s1 = take(tH2); S->s2 = take(tH2); S->s3 = take(tH2); S->b1 = take(tH2);
    S->r1y = take(yH2); S->s1y = take(yH2); S->s2y = take(yH2); S->s3y = take(yH2); S->b1y = take(yH2);
    S->e = take(kSlabBytes);
    S->total = off; S->n_tiles = n_tiles; S->tiles_per_group = tiles_per_group;
}

static size_t smem_fixed_bytes() {
    return kTblF + kTblV + 8 * kMaxGroupPts * 4 /*raw SoA incl. z, w, T*/ + 2 * kMaxGroupRays * 256 * 4 + 768;
}

// --------------------------------------------------------------------------------------------------
// weight / table packing (runs every call: parameters change every optimiser step; ~5 MB, a few us)
// --------------------------------------------------------------------------------------------------
__global__ void tc_pack_kernel(TcProgram P, const float* __restrict__ W, unsigned char* __restrict__ packed) {
    const int gi = blockIdx.y;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    float* T = reinterpret_cast<float*>(packed + P.tables_base);
    const int H = P.H;
    if (gi < P.n_gemms) {
        const TcGemm g = P.g[gi];
        long long base = 0;
        for (int i = 0; i < gi; ++i) base += gemm_stream_bytes(P.g[i]);
        __half* out = reinterpret_cast<__half*>(packed + base);
        const size_t chunk_halves = (size_t)(chunk_stream_bytes(g) / 2);
        // B tiles: for chunk j, slab s: [chunk_n rows][64 k] fp16, 128-byte swizzle.  One thread per 16-byte chunk (8 k of one row):
        // the 8 threads of a row read 256 contiguous bytes of the fp32 weights and write one 128-byte line.
        const long long total = (long long)g.n_chunks * g.k_slabs * g.chunk_n * 8;
        for (long long e = tid; e < total; e += nthr) {
            const int c = (int)(e & 7); long long r = e >> 3;
            const int nl = (int)(r % g.chunk_n); const long long t = r / g.chunk_n;
            const int s = (int)(t % g.k_slabs), j = (int)(t / g.k_slabs);
            const int n = j * g.chunk_n + nl, k0 = s * 64 + c * 8;
            const float* row = n < g.rows0 ? W + g.src0 + (long long)n * g.ld0 : W + g.src1 + (long long)(n - g.rows0) * g.ld1;
            const int c0 = n < g.rows0 ? g.col0 : g.col1;
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int k = k0 + i;
                v[i] = k < g.kvalid ? row[c0 + k] : ((k >= g.k2_start && k < g.k2_start + g.k2_cols) ? row[g.col_k2 + k - g.k2_start] : 0.f);
            }
            const size_t tile = (size_t)j * chunk_halves + (size_t)s * g.chunk_n * 64;
            *reinterpret_cast<uint4*>(out + tile + (size_t)nl * 64 + (size_t)((c ^ (nl & 7)) << 3)) =
                make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
        }
        // aux tile of every chunk: [chunk_n rows][16 k] fp16, 32-byte swizzle.  k = 0,1: bias hi, lo (fp16 pair: exact to 2^-22);
        // skip layer: k = 2..4 W_xyz hi (x A's xyz hi), 5..7 W_xyz hi (x xyz lo), 8..10 W_xyz lo (x xyz hi); the rest 0
        if (g.aux) {
            const long long boff = g.src0 + (long long)g.N * g.ld0;   // bias follows the weight block (N x ld0)
            for (int e = tid; e < g.N * 16; e += nthr) {
                const int n = e >> 4, k = e & 15, j = n / g.chunk_n, nl = n - j * g.chunk_n;
                float v = 0.f;
                if (k < 2) { float b = W[boff + n]; float hi = __half2float(__float2half_rn(b)); v = k == 0 ? hi : b - hi; }
                else if (g.aux == 2 && k < 11) {
                    float w = W[g.src0 + (long long)n * g.ld0 + (k - 2) % 3]; float hi = __half2float(__float2half_rn(w));
                    v = k < 8 ? hi : w - hi;
                }
                size_t off = (size_t)j * chunk_halves + (size_t)g.k_slabs * g.chunk_n * 64
                           + (size_t)(nl >> 3) * 128 + (size_t)(nl & 7) * 16 + (size_t)((((k >> 3) ^ ((nl >> 2) & 1)) << 3) | (k & 7));
                out[off] = __float2half_rn(v);
            }
        }
        // epilogue tables
        for (int n = tid; n < g.N; n += nthr) {
            float* t4 = T + g.tbl_off;
            switch (g.kind) {
                case GK_SUN2: case GK_SUN3: t4[n] = W[g.src0 + (long long)g.N * g.ld0 + n]; break;
                default: break;      // HEADA / per-ray tables are written by tc_pack_misc_kernel
            }
        }
    }
}

struct MiscOffsets { long long sigma_w, sigma_b, rgb0_b, rgb2_w, rgb2_b, sun0_w, sun0_b, sun3_w, sun3_b, sky0_w, sky0_b, sky2_w, sky2_b,
                               beta0_w, beta0_b, beta2_w, beta2_b, rgb0_w; int sun0_ld, beta0_ld, rgb0_ld; };

__global__ void tc_pack_misc_kernel(TcProgram P, MiscOffsets M, const float* __restrict__ W, unsigned char* __restrict__ packed) {
    float* T = reinterpret_cast<float*>(packed + P.tables_base);
    const int H = P.H, H2 = P.H2;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    if (P.nerf) {
        // sigma vector, rgb head table [0, w_r, w_g, w_b] per hidden unit, constants, view-direction columns + bias of rgb_from_xyzdir.0
        const int nd = 6 * P.pe_dir;
        for (int gi = 0; gi < P.n_gemms; ++gi) {
            const TcGemm& g = P.g[gi];
            if (g.kind == GK_TRUNK && g.last) for (int n = tid; n < H; n += nthr) T[g.vec_off + n] = W[M.sigma_w + n];
            if (g.kind == GK_HEADN) for (int n = tid; n < H2; n += nthr) {
                float* t = T + g.tbl_off + n * 4;
                t[0] = 0.f; t[1] = W[M.rgb2_w + n]; t[2] = W[M.rgb2_w + H2 + n]; t[3] = W[M.rgb2_w + 2 * H2 + n];
            }
        }
        if (tid == 0) {
            float* c = T + P.consts;
            c[0] = W[M.sigma_b]; c[1] = W[M.rgb2_b]; c[2] = W[M.rgb2_b + 1]; c[3] = W[M.rgb2_b + 2]; c[4] = c[5] = c[6] = c[7] = 0.f;
        }
        for (int n = tid; n < H2; n += nthr) {
            for (int c = 0; c < nd; ++c) T[P.sunw + c * H2 + n] = W[M.rgb0_w + (long long)n * M.rgb0_ld + H + c];
            T[P.sunw + nd * H2 + n] = W[M.rgb0_b + n];
        }
        return;
    }
    // layer-0 table [H][4] = (wx, wy, wz, b)
    for (int n = tid; n < H; n += nthr) {
        float* t = T + P.l0_tbl + n * 4;
        t[0] = W[P.l0_w + n * 3]; t[1] = W[P.l0_w + n * 3 + 1]; t[2] = W[P.l0_w + n * 3 + 2]; t[3] = W[P.l0_b + n];
    }
    // sigma vector on the last trunk GEMM, sun_v_net.6 vector on SUN3, HEADA table
    for (int gi = 0; gi < P.n_gemms; ++gi) {
        const TcGemm& g = P.g[gi];
        if (g.kind == GK_TRUNK && g.last) for (int n = tid; n < H; n += nthr) T[g.vec_off + n] = W[M.sigma_w + n];
        if (g.kind == GK_SUN3) for (int n = tid; n < H2; n += nthr) T[g.vec_off + n] = W[M.sun3_w + n];
        if (g.kind == GK_HEADA) {
            int nb = P.has_beta ? H2 : 0;
            for (int n = tid; n < g.N; n += nthr) {
                float* t = T + g.tbl_off + n * 4;
                if (n < nb) { t[0] = 0.f; t[1] = W[M.beta2_w + n]; t[2] = 0.f; t[3] = 0.f; T[g.vec_off + n] = W[M.beta2_w + n]; }
                else { int m = n - nb; t[0] = W[M.rgb0_b + m]; t[1] = W[M.rgb2_w + m]; t[2] = W[M.rgb2_w + H2 + m]; t[3] = W[M.rgb2_w + 2 * H2 + m]; }
            }
        }
    }
    if (tid == 0) {
        float* c = T + P.consts;
        c[0] = W[M.sigma_b]; c[1] = W[M.rgb2_b]; c[2] = W[M.rgb2_b + 1]; c[3] = W[M.rgb2_b + 2];
        c[4] = W[M.sun3_b]; c[5] = P.has_beta ? W[M.beta2_b] : 0.f; c[6] = 0.f; c[7] = 0.f;
    }
    for (int n = tid; n < H2; n += nthr) {
        for (int c = 0; c < 3; ++c) T[P.sunw + c * H2 + n] = W[M.sun0_w + (long long)n * M.sun0_ld + H + c];
        T[P.sunw + 3 * H2 + n] = W[M.sun0_b + n];
        if (P.has_beta) {
            for (int c = 0; c < P.tau; ++c) T[P.betaw + c * H2 + n] = W[M.beta0_w + (long long)n * M.beta0_ld + H + c];
            T[P.betaw + P.tau * H2 + n] = W[M.beta0_b + n];
        }
        float* s = T + P.sky + n * 4;
        s[0] = W[M.sky0_w + n * 3]; s[1] = W[M.sky0_w + n * 3 + 1]; s[2] = W[M.sky0_w + n * 3 + 2]; s[3] = W[M.sky0_b + n];
        for (int c = 0; c < 3; ++c) T[P.sky + 4 * H2 + c * H2 + n] = W[M.sky2_w + c * H2 + n];
    }
    if (tid < 3) T[P.sky + 7 * H2 + tid] = W[M.sky2_b + tid];
}

// --------------------------------------------------------------------------------------------------
// the fused kernel
// --------------------------------------------------------------------------------------------------
// Phase timestamps (clock64) of block 0's second tile, read back through snb_debug_read: per GEMM
// [wait-for-accumulator start, accumulator ready, epilogue done]; slot 60.. = layer-0 / compositing marks.
__device__ long long g_tc_dbg[64 * 4];
// The marks and the SNB_TC_DBG knobs exist only in probe builds (-DSNB_TC_PROBE, profiles/dev/build_variant.sh): the
// production kernel is register-bound (112 per thread) and every extra live value in the epilogue spills.
#ifdef SNB_TC_PROBE
#define TC_MARK(slot, k) do { if (dbg_on && tid_e == 0) g_tc_dbg[(slot) * 4 + (k)] = clock64(); } while (0)
#define TC_DBG(x) (x)
#else
#define TC_MARK(slot, k) do { (void)dbg_on; } while (0)
#define TC_DBG(x) 0
#endif

// Epilogue of NC (16) consecutive columns of one row: v = fp32 accumulators of columns n0..n0+NC-1.
// Tables live in shared memory (32-bit addresses): tF = [N] floats (F1) or [N][4] (F4), tV = [N] extra vector.
// Training mode (ys != nullptr): the pre-activations go to the yb stash and activations that are consumed in
// registers (first layers of the rgb / beta heads, last sun layer) go to their atoms stash.
struct EpiStash { unsigned char *y0, *y1, *act0, *act1; int gt; };     // 0: first column half of HEADA (beta) / every other kind

#define SIN_(x) ((TC_DBG(dbg) & 1) ? (x) : __sinf(x))
template <int NC>
__device__ __forceinline__ void sin_cols(int dbg, float* v, unsigned char* yarr, int gt, int F, int n0, int row) {
    if (yarr) {                      // training: the derivative cos(y) goes to the stash (same range reduction as the sine)
        float c[NC];
#pragma unroll
        for (int i = 0; i < NC; ++i) c[i] = __cosf(v[i]);
        yb_store_cols<NC>(yarr, gt, F, n0, row, c);
    }
#pragma unroll
    for (int i = 0; i < NC; ++i) v[i] = SIN_(v[i]);
}
#define LDS_T(addr) ((TC_DBG(dbg) & 64) ? make_float4(0.f, 0.f, 0.f, 0.f) : lds128(addr, tok))
template <int NC>
__device__ __forceinline__ void relu_cols(float* v) {
#pragma unroll
    for (int i = 0; i < NC; ++i) v[i] = fmaxf(v[i], 0.f);
}
template <int NC, bool NERF = false>
__device__ __forceinline__ void epi_cols(uint32_t tok, int dbg, int kind, bool skip, bool last, int has_beta, int n0, int H2, float* v,
                                          uint32_t a_base, int row, uint32_t tF, uint32_t tV, uint32_t sunb_row, uint32_t betab_row,
                                          float px, float py, float pz, const EpiStash& es, uint64_t* slab_bar, uint32_t slab_par,
                                          float& sig_dot, float& beta_dot, float& rgb0, float& rgb1, float& rgb2, float& sun_dot) {
    const int H = 2 * H2;
    if (NERF && kind == GK_HEADN) {  // rgb_from_xyzdir.0 (nerf.py:172): + per-ray view-direction term and bias, ReLU, dot with the 3 x H2 output layer
#pragma unroll
        for (int i = 0; i < NC; i += 4) {
            float4 b = LDS_T(sunb_row + (uint32_t)(n0 + i) * 4u);
            v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
        }
        relu_cols<NC>(v);
#pragma unroll
        for (int i = 0; i < NC; ++i) {
            float4 w = LDS_T(tF + (uint32_t)(n0 + i) * 16u);
            rgb0 = fmaf(w.y, v[i], rgb0); rgb1 = fmaf(w.z, v[i], rgb1); rgb2 = fmaf(w.w, v[i], rgb2);
        }
    } else if (kind == GK_TRUNK) {          // bias (and the skip layer's xyz term) are already in the accumulator (aux K-step)
        if (NERF) relu_cols<NC>(v); else
        sin_cols<NC>(dbg, v, es.y0, es.gt, H, n0, row);
        if (last) {
#pragma unroll
            for (int i = 0; i < NC; i += 4) {
                float4 w = LDS_T(tV + (uint32_t)(n0 + i) * 4u);
                sig_dot = fmaf(w.x, v[i], sig_dot); sig_dot = fmaf(w.y, v[i + 1], sig_dot);
                sig_dot = fmaf(w.z, v[i + 2], sig_dot); sig_dot = fmaf(w.w, v[i + 3], sig_dot);
            }
        }
        if (!(TC_DBG(dbg) & 4)) { if (slab_bar) mbar_wait(slab_bar + (n0 >> 6), slab_par, 8); store_act_cols<NC>(a_base, row, n0, v); }
    } else if (kind == GK_FEAT) {
        if (!(TC_DBG(dbg) & 4)) { if (slab_bar) mbar_wait(slab_bar + (n0 >> 6), slab_par, 8); store_act_cols<NC>(a_base, row, n0, v); }
    } else if (kind == GK_HEADA) {
        if (has_beta && n0 < H2) {
#pragma unroll
            for (int i = 0; i < NC; i += 4) {
                float4 b = LDS_T(betab_row + (uint32_t)(n0 + i) * 4u);
                v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
            }
            sin_cols<NC>(dbg, v, es.y0, es.gt, H2, n0, row);
            if (es.act0) atom_store_cols<NC>(es.act0, es.gt, (H2 + 63) >> 6, row, n0, v);
#pragma unroll
            for (int i = 0; i < NC; i += 4) {
                float4 w = LDS_T(tV + (uint32_t)(n0 + i) * 4u);
                beta_dot = fmaf(w.x, v[i], beta_dot); beta_dot = fmaf(w.y, v[i + 1], beta_dot);
                beta_dot = fmaf(w.z, v[i + 2], beta_dot); beta_dot = fmaf(w.w, v[i + 3], beta_dot);
            }
        } else {
            const int m0 = has_beta ? n0 - H2 : n0;
#pragma unroll
            for (int i = 0; i < NC; ++i) v[i] += LDS_T(tF + (uint32_t)(n0 + i) * 16u).x;
            sin_cols<NC>(dbg, v, es.y1, es.gt, H2, m0, row);
            if (es.act1) atom_store_cols<NC>(es.act1, es.gt, (H2 + 63) >> 6, row, m0, v);
#pragma unroll
            for (int i = 0; i < NC; ++i) {
                float4 w = LDS_T(tF + (uint32_t)(n0 + i) * 16u);
                rgb0 = fmaf(w.y, v[i], rgb0); rgb1 = fmaf(w.z, v[i], rgb1); rgb2 = fmaf(w.w, v[i], rgb2);
            }
        }
    } else if (kind == GK_SUN1) {
#pragma unroll
        for (int i = 0; i < NC; i += 4) {
            float4 b = LDS_T(sunb_row + (uint32_t)(n0 + i) * 4u);
            v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
        }
        sin_cols<NC>(dbg, v, es.y0, es.gt, H2, n0, row);
        if (!(TC_DBG(dbg) & 4)) store_act_cols<NC>(a_base, row, n0, v);
    } else {   // GK_SUN2 / GK_SUN3
#pragma unroll
        for (int i = 0; i < NC; i += 4) {
            float4 b = LDS_T(tF + (uint32_t)(n0 + i) * 4u);
            v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
        }
        sin_cols<NC>(dbg, v, es.y0, es.gt, H2, n0, row);
        if (kind == GK_SUN2) { if (!(TC_DBG(dbg) & 4)) store_act_cols<NC>(a_base, row, n0, v); }
        else {
            if (es.act0) atom_store_cols<NC>(es.act0, es.gt, (H2 + 63) >> 6, row, n0, v);
#pragma unroll
            for (int i = 0; i < NC; i += 4) {
                float4 w = LDS_T(tV + (uint32_t)(n0 + i) * 4u);
                sun_dot = fmaf(w.x, v[i], sun_dot); sun_dot = fmaf(w.y, v[i + 1], sun_dot);
                sun_dot = fmaf(w.z, v[i + 2], sun_dot); sun_dot = fmaf(w.w, v[i + 3], sun_dot);
            }
        }
    }
}
#undef LDS_T
#undef SIN_
// CG = 1: one CTA per 128-point tile.  CG = 2: CTA pair (cta_group::2): two SMs run two tiles in lockstep, the
// leader issues M=256 MMAs that read each CTA's own activation tile and HALF of every weight tile from each
// CTA's shared memory, so each SM streams / buffers only half of the weights.
template <int CG, bool TRAIN, bool NERF>
__global__ void __launch_bounds__(64 + 32 * (TRAIN ? kEpiWarpsFwdTrain : (NERF ? kEpiWarpsTrain : kEpiWarps)), 1) tc_render_kernel(const __grid_constant__ TcArgs A) {
    constexpr int EW = TRAIN ? kEpiWarpsFwdTrain : (NERF ? kEpiWarpsTrain : kEpiWarps), ES = EW / 4, ET = EW * 32;      // epilogue warps, column-block interleave per quadrant, epilogue threads
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const TcProgram& P = A.prog;
    Smem sm = carve(base, P, CG);
    const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0u;
    const int unit = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;      // work unit index (cluster or CTA)
    const int n_units = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int stage_bytes = P.stage_bytes / CG;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* T = reinterpret_cast<const float*>(A.packed + P.tables_base);

    if (threadIdx.x == 0) {
        // leader of a pair: a stage is full when its own copy has landed AND the peer has reported its half (relay arrival)
        for (int i = 0; i < P.n_stages; ++i) { mbar_init(&sm.full[i], (CG == 2 && cta_rank == 0) ? 2 : 1); mbar_init(&sm.empty[i], 1); }
        mbar_init(sm.acc_full, 1); mbar_init(sm.acc_full2, 1); mbar_init(sm.a_ready, CG); mbar_init(sm.a_ready2, CG);
        for (int i = 0; i < 4; ++i) mbar_init(&sm.slab_free[i], 1);     // one elected arrival per CTA of the pair
        fence_barrier_init();
    }
    if (CG == 2) { __syncthreads(); cluster_sync_all(); }       // both CTAs of the pair are running and their barriers are initialised
    if (warp == 1) { if (CG == 2) tmem_alloc_2cta(sm.tmem_ptr, 512); else tmem_alloc(sm.tmem_ptr, 512); }
    if (threadIdx.x >= 64 && threadIdx.x < 72) sm.consts[threadIdx.x - 64] = T[P.consts + threadIdx.x - 64];
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();          // peer barriers are initialised before anyone signals them
    tc_fence_after();
    const uint32_t tmem = *sm.tmem_ptr;
    // Work: groups of G rays.  CG = 2: the pair takes groups (2q, 2q+1); a missing / short group runs as masked rows.
    const int tiles_per_group = (A.G * A.S + kTile - 1) / kTile;
    const int n_work = CG == 2 ? (A.n_groups + 1) / 2 : A.n_groups;

    if (warp == 0) {
        // ================= weight producer (lane 0 issues; the warp stays converged) =================
        int st = 0; uint32_t ph = 0;
        for (int wk = unit; wk < n_work; wk += n_units) {
            for (int t = 0; t < tiles_per_group; ++t) {
                const unsigned char* src = A.packed;
                for (int gi = 0; gi < P.n_gemms; ++gi) {
                    const int per_chunk = P.g[gi].k_slabs + (P.g[gi].aux ? 1 : 0);       // the aux tile (one K-step wide) closes every chunk
                    const int n = P.g[gi].n_chunks * per_chunk;
                    for (int i = 0, sc = 0; i < n; ++i) {
                        const uint32_t tile_bytes = (uint32_t)P.g[gi].chunk_n * (sc < P.g[gi].k_slabs ? 128u : 32u), bytes = tile_bytes / CG;
                        if (lane == 0) {
                            mbar_wait(&sm.empty[st], ph ^ 1, 1);
                            if (TC_DBG(A.dbg) & 16) mbar_arrive(&sm.full[st]);        // knob: no copy (tensor pipe alone)
                            else {
                                mbar_arrive_expect_tx(&sm.full[st], bytes);
                                bulk_g2s(sm.b + (size_t)st * stage_bytes, src + cta_rank * bytes, bytes, &sm.full[st]);   // this CTA's rows of the tile
                            }
                        }
                        __syncwarp();
                        src += tile_bytes;
                        if (++sc == per_chunk) sc = 0;
                        if (++st == P.n_stages) { st = 0; ph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (leader CTA) / weight-arrival relay (peer CTA) =================
        int st = 0; uint32_t ph = 0, ready_ph = 0;
        const uint32_t a_base = smem_u32(sm.a), b_base = smem_u32(sm.b);
        const uint32_t tmem = __shfl_sync(0xffffffffu, *sm.tmem_ptr, 0);      // provably warp-uniform copy of the TMEM base
        if (CG == 2 && cta_rank == 1) {
            // The leader's MMA reads this CTA's half of every weight tile: when a stage has landed here, arrive on the LEADER's
            // full barrier of that stage (count 2 there: its own copy + this arrival).  Relaxed: a cluster-scope release costs
            // ~1000 cycles per arrival and the data were delivered by the async proxy before this CTA's barrier completed.
            const uint32_t leader_full = mapa_u32(smem_u32(sm.full), 0);
            for (int wk = unit; wk < n_work; wk += n_units)
                for (int t = 0; t < tiles_per_group; ++t)
                    for (int gi = 0; gi < P.n_gemms; ++gi) {
                        const int n = P.g[gi].n_chunks * (P.g[gi].k_slabs + (P.g[gi].aux ? 1 : 0));
                        for (int i = 0; i < n; ++i) {
                            mbar_wait(&sm.full[st], ph, 5);
                            if (lane == 0) mbar_arrive_cluster_relaxed(leader_full + (uint32_t)st * 8u);
                            __syncwarp();
                            if (++st == P.n_stages) { st = 0; ph ^= 1; }
                        }
                    }
        } else {
            // Single-thread issue.  The loop is latency-critical: one elected region issues up to two weight stages (8 MMAs,
            // 1024 tensor cycles) and must cost less than that -- every lane polls the barriers (no lane-0 block + __syncwarp),
            // the GEMM's fields live in registers, descriptors are 64-bit adds on precomputed bases.
            const uint64_t a_desc0 = umma_desc_k_sw128(a_base), b_desc0 = umma_desc_k_sw128(b_base);
            const uint32_t stage_desc = (uint32_t)(stage_bytes >> 4);
            const uint64_t aux_a_desc = umma_desc_k_sw32(smem_u32(sm.a_aux)), aux_b_desc0 = umma_desc_k_sw32(b_base);
            const bool deep = P.n_stages >= 4;
#ifdef SNB_TC_PROBE
            int itile = 0;
#endif
            for (int wk = unit; wk < n_work; wk += n_units) {
                for (int t = 0; t < tiles_per_group; ++t) {
#ifdef SNB_TC_PROBE
                    const bool trace_tile = blockIdx.x == 0 && itile++ == 1; int trace_n = 0;
#endif
                    for (int gi = 0; gi < P.n_gemms; ++gi) {
                        const int k_slabs = P.g[gi].k_slabs, n_chunks = P.g[gi].n_chunks, chunk_n = P.g[gi].chunk_n, K = P.g[gi].K;
                        const int k_early = P.g[gi].k_early, free_slabs = P.g[gi].free_slabs, aux = P.g[gi].aux, a_slab0 = P.g[gi].a_slab0;
                        const uint32_t idesc = CG == 2 ? umma_idesc_f16_m256((uint32_t)chunk_n) : umma_idesc_f16((uint32_t)chunk_n);
#ifdef SNB_TC_PROGRESS
                        if (g_hang_host && blockIdx.x < 16 && lane == 0) ((volatile unsigned int*)g_hang_host)[64 + blockIdx.x * 8 + 3] = (unsigned)(wk << 16 | t << 8 | gi);
#endif
                        mbar_wait(sm.a_ready, ready_ph, 2);          // (both ready barriers flip once per GEMM: one parity)
                        for (int j = 0; j < n_chunks; ++j) {
                            const uint32_t d_tm = tmem + (uint32_t)(j * chunk_n);
                            for (int s = 0; s < k_slabs;) {
                                if (j == 0 && s == k_early) mbar_wait(sm.a_ready2, ready_ph, 7);     // the rest of the input tile / accumulator columns of the later chunks
                                const bool pair = deep && s + 1 < k_slabs && !(j == 0 && s + 1 == k_early);
                                int st1 = st + 1; uint32_t ph1 = ph; if (st1 == P.n_stages) { st1 = 0; ph1 ^= 1; }
                                const uint64_t da = a_desc0 + (uint64_t)((uint32_t)(a_slab0 + s) * (kSlabBytes >> 4));
                                const uint64_t db = b_desc0 + (uint64_t)((uint32_t)st * stage_desc);
                                const uint64_t db1 = b_desc0 + (uint64_t)((uint32_t)st1 * stage_desc);
                                int ksteps = K - s * 64; ksteps = (ksteps > 64 ? 64 : ksteps) >> 4;
                                int ksteps1 = K - (s + 1) * 64; ksteps1 = (ksteps1 > 64 ? 64 : ksteps1) >> 4;
#ifdef SNB_TC_PROBE
                                const bool tr = trace_tile && (gi == 1 || gi == 2) && trace_n < 40 && lane == 0;
                                if (tr) g_tc_dbg[(20 + trace_n) * 4 + 0] = clock64();
#endif
                                mbar_wait(&sm.full[st], ph, 3);           // (pair mode: this CTA's copy and the peer's relay arrival)
                                if (pair) mbar_wait(&sm.full[st1], ph1, 3);
                                tc_fence_after();
#ifdef SNB_TC_PROBE
                                if (tr) g_tc_dbg[(20 + trace_n) * 4 + 1] = clock64();
#endif
                                if (elect_one()) {        // warp-uniform operands + elect: UTCHMMA takes uniform registers directly
                                    if (ksteps == 4) {
#pragma unroll
                                        for (int k = 0; k < 4; ++k) {
                                            if (CG == 2) umma_f16_ss_2cta(d_tm, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (s | k) != 0);
                                            else umma_f16_ss(d_tm, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (s | k) != 0);
                                        }
                                    } else {
                                        for (int k = 0; k < ksteps; ++k) {
                                            if (CG == 2) umma_f16_ss_2cta(d_tm, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (s | k) != 0);
                                            else umma_f16_ss(d_tm, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (s | k) != 0);
                                        }
                                    }
                                    if (CG == 2) umma_commit_2cta(&sm.empty[st], 3); else umma_commit(&sm.empty[st]);
                                    if (j >= 1 && s < free_slabs) { if (CG == 2) umma_commit_2cta(&sm.slab_free[s], 3); else umma_commit(&sm.slab_free[s]); }
                                    if (pair) {
                                        const uint64_t da1 = da + (uint64_t)(kSlabBytes >> 4);
                                        if (ksteps1 == 4) {
#pragma unroll
                                            for (int k = 0; k < 4; ++k) {
                                                if (CG == 2) umma_f16_ss_2cta(d_tm, da1 + (uint64_t)(2 * k), db1 + (uint64_t)(2 * k), idesc, 1);
                                                else umma_f16_ss(d_tm, da1 + (uint64_t)(2 * k), db1 + (uint64_t)(2 * k), idesc, 1);
                                            }
                                        } else {
                                            for (int k = 0; k < ksteps1; ++k) {
                                                if (CG == 2) umma_f16_ss_2cta(d_tm, da1 + (uint64_t)(2 * k), db1 + (uint64_t)(2 * k), idesc, 1);
                                                else umma_f16_ss(d_tm, da1 + (uint64_t)(2 * k), db1 + (uint64_t)(2 * k), idesc, 1);
                                            }
                                        }
                                        if (CG == 2) umma_commit_2cta(&sm.empty[st1], 3); else umma_commit(&sm.empty[st1]);
                                        if (j >= 1 && s + 1 < free_slabs) { if (CG == 2) umma_commit_2cta(&sm.slab_free[s + 1], 3); else umma_commit(&sm.slab_free[s + 1]); }
                                    }
                                    // one barrier per N-chunk index: two commits of one GEMM on a single barrier could both land before the
                                    // epilogue looks at it (it may still be draining its stash copy), and the parity wait would miss a phase
                                    if (!aux && s + (pair ? 2 : 1) == k_slabs) {
                                        if (j == 0) { if (CG == 2) umma_commit_2cta(sm.acc_full, 3); else umma_commit(sm.acc_full); }
                                        else { if (CG == 2) umma_commit_2cta(sm.acc_full2, 3); else umma_commit(sm.acc_full2); }
                                    }
                                }
                                __syncwarp();
#ifdef SNB_TC_PROBE
                                if (tr) { g_tc_dbg[(20 + trace_n) * 4 + 2] = clock64(); g_tc_dbg[(20 + trace_n) * 4 + 3] = (long long)(gi << 16 | (j * k_slabs + s) << 8 | (pair ? 1 : 0)); }
                                if (trace_tile && (gi == 1 || gi == 2)) ++trace_n;
#endif
                                if (pair) { s += 2; st = st1 + 1; ph = ph1; if (st == P.n_stages) { st = 0; ph ^= 1; } }
                                else { s += 1; st = st1; ph = ph1; }
                            }
                            if (aux) {
                                // the chunk's last K-step: [1 1 | xyz..] x [bias hi lo | W_xyz..] adds the bias (and the skip layer's
                                // xyz term) inside the accumulator, so the epilogue needs no per-column table
                                mbar_wait(&sm.full[st], ph, 3);
                                tc_fence_after();
                                if (elect_one()) {
                                    const uint64_t dbx = aux_b_desc0 + (uint64_t)((uint32_t)st * stage_desc);
                                    if (CG == 2) umma_f16_ss_2cta(d_tm, aux_a_desc, dbx, idesc, 1); else umma_f16_ss(d_tm, aux_a_desc, dbx, idesc, 1);
                                    if (CG == 2) umma_commit_2cta(&sm.empty[st], 3); else umma_commit(&sm.empty[st]);
                                    if (j == 0) { if (CG == 2) umma_commit_2cta(sm.acc_full, 3); else umma_commit(sm.acc_full); }
                                    else { if (CG == 2) umma_commit_2cta(sm.acc_full2, 3); else umma_commit(sm.acc_full2); }
                                }
                                __syncwarp();
                                if (++st == P.n_stages) { st = 0; ph ^= 1; }
                            }
                        }
                        ready_ph ^= 1;
                    }
                }
            }
        }
    } else {
        // ================= epilogue warps (EW warps; ES threads per point) =================
        // Warp w may touch TMEM lanes [32*(w%4), +32).  The ES warps of a quadrant interleave the 32-column
        // blocks of every layer; several warps per scheduler let MUFU work of one overlap pack/store work of another.
        const int tid_e = threadIdx.x - 64;
        const int quad = warp & 3, half = (warp - 2) >> 2;      // half = sub-warp index 0..ES-1 within the quadrant
        const int row = quad * 32 + lane;                // point (row of the tile) owned by this thread
        const uint32_t a_base = smem_u32(sm.a);
        const uint32_t tm_row = tmem + ((uint32_t)(quad * 32) << 16);
        const uint32_t tF = smem_u32(sm.tblF), tV = smem_u32(sm.tblV);
        const int H = P.H, H2 = P.H2, S = A.S;
        const int aux_col = NERF ? 3 : 8;                // per-ray auxiliary direction in the ray row: sun (8:11), nerf: view direction (3:6)
        int tile_counter = 0;
        const uint32_t ready_bar = CG == 2 ? mapa_u32(smem_u32(sm.a_ready), 0) : 0u;     // the leader's a_ready barriers
        const uint32_t ready2_bar = CG == 2 ? mapa_u32(smem_u32(sm.a_ready2), 0) : 0u;
#ifdef SNB_TC_PROGRESS      // deadlock diagnosis: per-block progress words in the pinned hang mirror (SNB_TC_HANG_MIRROR=1)
        unsigned int n_sig[2] = {0u, 0u};
#endif
        auto signal_ready = [&](int which) {          // one elected arrival per CTA
            if (tid_e == 0) {
#ifdef SNB_TC_PROGRESS
                if (g_hang_host && blockIdx.x < 16) { ++n_sig[which]; ((volatile unsigned int*)g_hang_host)[64 + blockIdx.x * 8 + which] = n_sig[which]; }
#endif
                // (the writes this publishes were fenced into the async proxy and ordered by the named barrier before this point;
                //  the tensor core that reads them is this CTA's own, so the remote arrival needs no cluster-scope release)
                if (CG == 2 && cta_rank != 0) mbar_arrive_cluster_relaxed(which ? ready2_bar : ready_bar);
                else mbar_arrive(which ? sm.a_ready2 : sm.a_ready);
            }
        };
        for (int wk = unit; wk < n_work; wk += n_units) {
            const int grp = CG == 2 ? 2 * wk + (int)cta_rank : wk;
            const int r0 = grp * A.G;
            const int n_rays = grp < A.n_groups ? min(A.G, A.R - r0) : 0;        // 0: this CTA only keeps the pair in lockstep
            const int Pg = n_rays * S;
            const int n_tiles = tiles_per_group;
            if (NERF) {
                // ---- nerf: encoded view direction (Mapping, nerf.py:60-66: per frequency sin(f d) then cos(f d), f = 2^j) and its
                //      contribution to rgb_from_xyzdir.0 as a per-ray bias (the direction is constant along the ray) ----
                const int nd = 6 * P.pe_dir;
                for (int idx = tid_e; idx < n_rays * nd; idx += ET) {
                    const int gr = idx / nd, k = idx - gr * nd, j = k / 6, f = k - 6 * j;
                    const float* sd = A.aux ? A.aux + (size_t)(r0 + gr) * 3 : A.rays + (size_t)(r0 + gr) * A.ray_cols + aux_col;
                    const float arg = __fmul_rn((float)(1 << j), sd[f % 3]);
                    sm.betab[gr * 64 + k] = f < 3 ? sinf(arg) : cosf(arg);
                }
                named_bar_sync(1, ET);
                for (int idx = tid_e; idx < n_rays * H2; idx += ET) {
                    const int gr = idx / H2, n = idx - gr * H2;
                    float v = T[P.sunw + nd * H2 + n];
                    for (int k = 0; k < nd; ++k) v = fmaf(T[P.sunw + k * H2 + n], sm.betab[gr * 64 + k], v);
                    sm.sunb[gr * H2 + n] = v;
                }
            }
            // ---- per-ray tables: sun / embedding terms of the first head layers, sky colour ----
            for (int idx = tid_e; !NERF && idx < n_rays * H2; idx += ET) {
                int gr = idx / H2, n = idx - gr * H2;
                const float* sd = A.aux ? A.aux + (size_t)(r0 + gr) * 3 : A.rays + (size_t)(r0 + gr) * A.ray_cols + aux_col;
                float v = T[P.sunw + 3 * H2 + n];
                v = fmaf(T[P.sunw + n], sd[0], v); v = fmaf(T[P.sunw + H2 + n], sd[1], v); v = fmaf(T[P.sunw + 2 * H2 + n], sd[2], v);
                sm.sunb[gr * H2 + n] = v;
                if (P.has_beta) {
                    float b = T[P.betaw + P.tau * H2 + n];
                    const float* te = A.t_emb + (size_t)(r0 + gr) * P.tau;
                    for (int c = 0; c < P.tau; ++c) b = fmaf(T[P.betaw + c * H2 + n], te[c], b);
                    sm.betab[gr * H2 + n] = b;
                }
            }
            for (int gr = warp - 2; !NERF && gr < n_rays; gr += EW) {          // sky_color(sun_d): per ray (satnerf.py:201)
                const float* sd = A.aux ? A.aux + (size_t)(r0 + gr) * 3 : A.rays + (size_t)(r0 + gr) * A.ray_cols + aux_col;
                float o0 = 0.f, o1 = 0.f, o2 = 0.f;
                for (int n = lane; n < H2; n += 32) {
                    const float* s4 = T + P.sky + n * 4;
                    float hdn = fmaxf(fmaf(s4[2], sd[2], fmaf(s4[1], sd[1], fmaf(s4[0], sd[0], s4[3]))), 0.f);
                    o0 = fmaf(T[P.sky + 4 * H2 + n], hdn, o0); o1 = fmaf(T[P.sky + 5 * H2 + n], hdn, o1); o2 = fmaf(T[P.sky + 6 * H2 + n], hdn, o2);
                }
                for (int off = 16; off; off >>= 1) { o0 += __shfl_xor_sync(~0u, o0, off); o1 += __shfl_xor_sync(~0u, o1, off); o2 += __shfl_xor_sync(~0u, o2, off); }
                if (lane == 0) {
                    sm.skyc[gr * 4] = sigmoid_f(o0 + T[P.sky + 7 * H2]); sm.skyc[gr * 4 + 1] = sigmoid_f(o1 + T[P.sky + 7 * H2 + 1]);
                    sm.skyc[gr * 4 + 2] = sigmoid_f(o2 + T[P.sky + 7 * H2 + 2]);
                }
            }
            for (int t = 0; t < n_tiles; ++t) {
                const bool dbg_on = TC_DBG(blockIdx.x == 0 && tile_counter == 1); ++tile_counter;
                TC_MARK(60, 0);
                // ---- sample position of this thread's point ----
                const int p = t * kTile + row;
                const bool valid = p < Pg;
                const int rl = valid ? p / S : 0;                 // ray within the group
                float px = 0.f, py = 0.f, pz = 0.f;
                if (valid) {
                    const size_t gp = (size_t)r0 * S + p;
                    float zz = A.z[gp];
                    if (half == 0) sm.z[p] = zz;
                    if (A.xyz) { px = A.xyz[gp * 3]; py = A.xyz[gp * 3 + 1]; pz = A.xyz[gp * 3 + 2]; }
                    else {
                        const float* ray = A.rays + (size_t)(r0 + rl) * A.ray_cols;       // rendering.py:81 / :104
                        px = __fadd_rn(ray[0], __fmul_rn(ray[A.dir_col], zz));
                        py = __fadd_rn(ray[1], __fmul_rn(ray[A.dir_col + 1], zz));
                        pz = __fadd_rn(ray[2], __fmul_rn(ray[A.dir_col + 2], zz));
                    }
                }
                const int gt = grp * tiles_per_group + t;         // global tile id (stash index)
                unsigned char* const sb = (TRAIN && grp < A.n_groups) ? A.stash_base : nullptr;     // compile-time null in the inference instantiation; the idle half of an odd pair stashes nothing
                if (sb && half == 0) {
                    // extra input block of the weight-gradient GEMMs: [x y z | sun_d | t_emb | 1 | 0 ...] (64 fp16 per point)
                    float e[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) e[i] = 0.f;
                    if (valid) {
                        const float* sd = A.aux ? A.aux + (size_t)(r0 + rl) * 3 : A.rays + (size_t)(r0 + rl) * A.ray_cols + aux_col;
                        e[0] = px; e[1] = py; e[2] = pz; e[3] = sd[0]; e[4] = sd[1]; e[5] = sd[2]; e[10] = 1.f;
                        if (P.has_beta) {
                            const float* te = A.t_emb + (size_t)(r0 + rl) * P.tau;
                            e[6] = te[0]; if (P.tau > 1) e[7] = te[1]; if (P.tau > 2) e[8] = te[2]; if (P.tau > 3) e[9] = te[3];
                        }
                    }
                    unsigned char* ea = sb + A.stash.e;
                    *reinterpret_cast<uint4*>(atom_chunk(ea, gt, 1, row, 0)) = make_uint4(pack_half2(e[0], e[1]), pack_half2(e[2], e[3]), pack_half2(e[4], e[5]), pack_half2(e[6], e[7]));
                    *reinterpret_cast<uint4*>(atom_chunk(ea, gt, 1, row, 8)) = make_uint4(pack_half2(e[8], e[9]), pack_half2(e[10], e[11]), 0u, 0u);
#pragma unroll
                    for (int c = 2; c < 8; ++c) *reinterpret_cast<uint4*>(atom_chunk(ea, gt, 1, row, c * 8)) = make_uint4(0u, 0u, 0u, 0u);
                }
                const uint32_t sunb_row = smem_u32(sm.sunb) + (uint32_t)(rl * H2) * 4u;
                const uint32_t betab_row = smem_u32(sm.betab) + (uint32_t)(rl * H2) * 4u;
                const int l0_split = NERF ? 0 : P.g[0].k_early * 64;         // columns published with the first ready signal (0: none)
                if (NERF) {
                    // ---- nerf: the encoded position (Mapping, nerf.py:60-66; 6 * pe_xyz <= 64 values, the rest 0) as fp16 into the
                    //      extra K-slab of the A tile; trunk layer 0 is the first GEMM ----
                    named_bar_sync(1, ET);                        // (the previous tile's scratch reads of the A tile are done)
                    const int nx = 6 * P.pe_xyz;
                    const float pc[3] = {px, py, pz};
                    for (int c = half; c < 8; c += ES) {
                        float v[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int k = c * 8 + i, j = k / 6, f = k - 6 * j;
                            const float arg = __fmul_rn((float)(1 << j), f % 3 == 0 ? pc[0] : (f % 3 == 1 ? pc[1] : pc[2]));
                            v[i] = k < nx ? (f < 3 ? sinf(arg) : cosf(arg)) : 0.f;
                        }
                        sts128(a_chunk_addr(a_base, row, P.pe_slab * 64 + c * 8), pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
                    }
                } else {
                // ---- trunk layer 0 on CUDA cores: sin(30 (W0 x + b0)), K = 3 (satnerf.py:105-106) ----
                if (half == 0) { sm.wt[row] = px; sm.wt[kTile + row] = py; sm.wt[2 * kTile + row] = pz; }     // positions of the tile's points
                table_copy<ET>(sm.tblF, T + P.l0_tbl, H * 16, tid_e);
                cp_async_wait_all();
                named_bar_sync(1, ET);
                const uint32_t tok0 = fresh_token(0xffffu);
                // Thread mapping of this layer: 4 rows (lane + 32k) x 8 columns per step, so that one 16-byte table read
                // ([wx wy wz b] of a column: a 512-byte register fill per warp) feeds 4 outputs instead of 1 -- with one row per
                // thread the table reads alone took ~8 000 cycles of the shared-memory pipe per tile.
                float rx[4], ry[4], rz[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) { rx[k] = sm.wt[lane + 32 * k]; ry[k] = sm.wt[kTile + lane + 32 * k]; rz[k] = sm.wt[2 * kTile + lane + 32 * k]; }
                for (int pass = 0; pass < 2; ++pass) {
                    const int c_lo = pass ? l0_split : 0, c_hi = pass ? H : l0_split;
                    for (int n0 = c_lo + (warp - 2) * 8; n0 < c_hi; n0 += 8 * EW) {
                        float4 w[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) w[i] = lds128(tF + (uint32_t)(n0 + i) * 16u, tok0);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const int r = lane + 32 * k;
                            float v[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                float y = fmaf(w[i].z, rz[k], fmaf(w[i].y, ry[k], fmaf(w[i].x, rx[k], w[i].w)));
                                v[i] = __fmul_rn(30.0f, y);
                            }
                            if (sb) {                       // d sin(30 y) / dy = 30 cos(30 y)
                                float c[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) c[i] = 30.0f * __cosf(v[i]);
                                yb_store_cols<8>(sb + A.stash.y[0], gt, H, n0, r, c);
                            }
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] = __sinf(v[i]);
                            sts128(a_chunk_addr(a_base, r, n0), pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
                        }
                    }
                    if (pass == 0 && l0_split > 0) {              // low K-slabs of the first GEMM's input are in place
                        fence_proxy_async_smem();
                        named_bar_sync(1, ET);
                        signal_ready(0);
                        // training: activation tiles leave for the stash in two bulk groups (low K-slabs now, the rest when the tile is
                        // complete), so each group has half a layer to drain before its slabs are rewritten
                        if (sb && tid_e == 0) { bulk_s2g(sb + A.stash.a[0] + (size_t)gt * P.a_slabs * kSlabBytes, sm.a, (uint32_t)(l0_split >> 6) * kSlabBytes); bulk_commit(); }
                    }
                }
                }
                fence_proxy_async_smem();
                named_bar_sync(1, ET);                  // the whole input tile of the first GEMM is in place; layer-0 table dead
                if (half == 0) {                                 // aux operand row of this point: [1 1 | xyz hi | xyz lo | xyz hi | 0..] (fp16 hi/lo pairs)
                    const __half xh = __float2half_rn(px), yh = __float2half_rn(py), zh = __float2half_rn(pz);
                    const uint32_t one = 0x3c003c00u;
                    const uint32_t w1 = pack_half2(__half2float(xh), __half2float(yh));
                    const uint32_t w2 = pack_half2(__half2float(zh), px - __half2float(xh));
                    const uint32_t w3 = pack_half2(py - __half2float(yh), pz - __half2float(zh));
                    const uint32_t w4 = pack_half2(__half2float(xh), __half2float(yh));
                    const uint32_t w5 = pack_half2(__half2float(zh), 0.f);
                    const uint32_t base_aux = smem_u32(sm.a_aux) + (uint32_t)(row >> 3) * 256u + (uint32_t)(row & 7) * 32u;
                    const uint32_t sw = ((uint32_t)(row >> 2) & 1u) << 4;
                    sts128(base_aux + (0u ^ sw), one, w1, w2, w3);          // k 0..7 : 1 1 | x y z hi | x y z lo
                    sts128(base_aux + (16u ^ sw), w4, w5, 0u, 0u);          // k 8..15: x y z hi | 0
                }
                fence_proxy_async_smem();
                named_bar_sync(1, ET);                  // aux rows in place before the second ready signal
                if (sb && tid_e == 0) {
                    const uint32_t lo = (uint32_t)(l0_split >> 6) * kSlabBytes;
                    bulk_s2g(sb + A.stash.a[0] + (size_t)gt * P.a_slabs * kSlabBytes + lo, sm.a + lo, (uint32_t)P.a_slabs * kSlabBytes - lo); bulk_commit();
                }
                bool prev_split = l0_split > 0;                  // the previous activation tile left as two bulk groups
                {   const TcGemm& g0 = P.g[0];
                    if (g0.fmt != TF_NONE) table_copy<ET>(sm.tblF, T + g0.tbl_off, g0.N * g0.fmt * 4, tid_e);
                    if (g0.has_vec) table_copy<ET>(sm.tblV, T + g0.vec_off, g0.N * 4, tid_e); }
                if (l0_split == 0) signal_ready(0);
                signal_ready(1);
                TC_MARK(60, 1);

                float sig_dot = 0.f, beta_dot = 0.f, rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f, sun_dot = 0.f;
                for (int gi = 0; gi < P.n_gemms; ++gi) {
                    const TcGemm& g = P.g[gi];
#ifdef SNB_TC_PROGRESS
                    if (g_hang_host && blockIdx.x < 16 && tid_e == 0) ((volatile unsigned int*)g_hang_host)[64 + blockIdx.x * 8 + 2] = (unsigned)(wk << 16 | t << 8 | gi);
#endif
                    cp_async_wait_all();
                    // the stash copy of the slabs chunk 0 is about to rewrite (low K-slabs) has left shared memory; a HIGH group of
                    // the previous tile may still be draining
                    if (sb && tid_e == 0) { if (prev_split) bulk_wait_read1(); else bulk_wait_read(); }
                    named_bar_sync(1, ET);              // tables of this GEMM are in shared memory
                    TC_MARK(gi, 0);
                    // N-chunks (<=256 columns) complete one after the other: the epilogue of chunk 0 runs while the tensor core
                    // works on chunk 1.  Chunk 1's MMAs walk the K-slabs in order and release each one (slab_free[]) as soon as
                    // they are done with it, so chunk 0's fp16 results go straight into the low K-slabs of the tile.
                    const int kind = g.kind, N = g.N, n_chunks = g.n_chunks, chunk_n = g.chunk_n;
                    const bool skip = g.skip != 0, last = g.last != 0;
                    const bool stores = kind == GK_TRUNK || kind == GK_FEAT || kind == GK_SUN1 || kind == GK_SUN2;
                    EpiStash es; es.gt = gt; es.y0 = es.y1 = es.act0 = es.act1 = nullptr;
                    if (sb && !(SNB_DEV_DBG(A.dbg) & 128)) {
                        if (kind == GK_TRUNK) es.y0 = sb + A.stash.y[gi + 1];
                        else if (kind == GK_HEADA) { es.y0 = sb + A.stash.b1y; es.act0 = sb + A.stash.b1; es.y1 = sb + A.stash.r1y; es.act1 = sb + A.stash.r1; }
                        else if (kind == GK_SUN1) es.y0 = sb + A.stash.s1y;
                        else if (kind == GK_SUN2) es.y0 = sb + A.stash.s2y;
                        else if (kind == GK_SUN3) { es.y0 = sb + A.stash.s3y; es.act0 = sb + A.stash.s3; }
                    }
                    uint32_t tok = 0;
                    const bool next_early = gi + 1 < P.n_gemms && n_chunks > 1 && P.g[gi + 1].k_early > 0;
                    bool early_signaled = false, low_dumped = false;
                    unsigned char* dump_dst = nullptr;           // stash array of the activation tile this GEMM produces
                    if (sb && !(SNB_DEV_DBG(A.dbg) & 256)) {
                        const int fgs2 = (H2 + 63) >> 6;
                        if (kind == GK_TRUNK) dump_dst = sb + A.stash.a[gi + 1] + (size_t)gt * P.a_slabs * kSlabBytes;
                        else if (kind == GK_FEAT) dump_dst = sb + A.stash.feat + (size_t)gt * P.a_slabs * kSlabBytes;
                        else if (kind == GK_SUN1) dump_dst = sb + A.stash.s1 + (size_t)gt * fgs2 * kSlabBytes;
                        else if (kind == GK_SUN2) dump_dst = sb + A.stash.s2 + (size_t)gt * fgs2 * kSlabBytes;
                    }
                    for (int ch = 0; ch < n_chunks; ++ch) {
                        // one thread polls the mbarrier; the others park in a hardware barrier instead of spinning on shared memory
#ifdef SNB_V_ONE_ACC
#error "SNB_V_ONE_ACC is no longer supported"
#else
                        // barrier phases follow from the tile / GEMM counters (no per-thread parity registers):
                        // chunk 0 flips acc_full once per GEMM, chunk 1 flips acc_full2 once per two-chunk GEMM
                        if (tid_e == 0) {
                            const int tile_seq = tile_counter - 1;
                            if (ch == 0) mbar_wait(sm.acc_full, (uint32_t)(tile_seq * P.n_gemms + gi) & 1u, 4);
                            else {
                                mbar_wait(sm.acc_full2, (uint32_t)(tile_seq * P.n_two + g.two_idx) & 1u, 4);
                                if (sb) { if (low_dumped) bulk_wait_read1(); else bulk_wait_read(); }      // the high slabs' previous stash copy has left
                            }
                        }
#endif
#ifdef SNB_TC_PROBE
                        if (gi == 1 && ch == 1) TC_MARK(53, 0);
#endif
                        named_bar_sync(2, ET);
#ifdef SNB_TC_PROBE
                        if (gi == 1 && ch == 1) TC_MARK(53, 1);
#endif
                        tc_fence_after();
                        if (ch == 0) { tok = fresh_token((uint32_t)gi); TC_MARK(gi, 1); } else { TC_MARK(gi, 3); }
                        const bool final_chunk = ch == n_chunks - 1;
                        // 32-column blocks in two 16-column halves: the TMEM load of one half is in flight while the other half goes
                        // through bias / sin / pack / store (a block-wide load + wait cost ~200 exposed cycles per block)
                        {
                            const int n_end = min((ch + 1) * chunk_n, N);
                            int n0 = ch * chunk_n + half * 32;
                            bool have = n0 < n_end;
                            uint32_t va[16], vb[16];
                            uint64_t* const slab_bar = (stores && !final_chunk && g.free_slabs > 0) ? sm.slab_free : nullptr;      // (free_slabs = 0: the GEMM does not read the slabs chunk 0 rewrites)
                            const uint32_t slab_par = (uint32_t)((tile_counter - 1) * P.n_store2 + g.store2_idx) & 1u;
                            // (the plain trunk layers -- 6 of the 12 GEMMs -- take a copy of the loop with the layer kind as a compile-time
                            //  constant: the epilogue is bound by instruction issue, and the per-call kind dispatch is ~5 % of its instructions)
#define SNB_BLOCK_LOOP(KIND, LAST, NERF_)                                                                                                             \
                            if (have) tmem_ld16(tm_row + (uint32_t)n0, va);                                                                           \
                            while (have) {                                                                                                            \
                                tmem_ld_wait16(va);                                                                                                   \
                                tmem_ld16(tm_row + (uint32_t)(n0 + 16), vb);                                                                          \
                                epi_cols<16, NERF_>(tok, A.dbg, KIND, skip, LAST, P.has_beta, n0, H2, reinterpret_cast<float*>(va), a_base, row, tF, tV, sunb_row, betab_row, \
                                             px, py, pz, es, slab_bar, slab_par, sig_dot, beta_dot, rgb0, rgb1, rgb2, sun_dot);                       \
                                tmem_ld_wait16(vb);                                                                                                   \
                                const int n1 = n0 + 32 * ES;                                                                                     \
                                const bool more = n1 < n_end;                                                                                         \
                                if (more) tmem_ld16(tm_row + (uint32_t)n1, va);                                                                       \
                                epi_cols<16, NERF_>(tok, A.dbg, KIND, skip, LAST, P.has_beta, n0 + 16, H2, reinterpret_cast<float*>(vb), a_base, row, tF, tV, sunb_row, betab_row, \
                                             px, py, pz, es, slab_bar, slab_par, sig_dot, beta_dot, rgb0, rgb1, rgb2, sun_dot);                       \
                                n0 = n1; have = more;                                                                                                 \
                            }
                            if (NERF) { SNB_BLOCK_LOOP(kind, last, true) }
                            else if (kind == GK_TRUNK && !last) { SNB_BLOCK_LOOP(GK_TRUNK, false, false) }
                            else { SNB_BLOCK_LOOP(kind, last, false) }
#undef SNB_BLOCK_LOOP
                        }
                        if (!final_chunk) {
                            tc_fence_before();
                            if (next_early) {
                                // Chunk 0 is drained and its results sit (fenced) in the low K-slabs: the next GEMM's first MMAs may be
                                // queued right behind this GEMM's chunk 1 -- they touch neither its accumulator columns nor the high
                                // K-slabs the last chunk's epilogue is about to write -- so the tensor pipe never drains between layers.
                                if (!(TC_DBG(A.dbg) & 32)) fence_proxy_async_smem();
                                named_bar_sync(1, ET);
                                signal_ready(0);
                                early_signaled = true;
                                if (dump_dst && stores && chunk_n % 64 == 0) {          // low K-slabs are final: first bulk group of the stash copy
                                    if (tid_e == 0) { bulk_s2g(dump_dst, sm.a, (uint32_t)(chunk_n >> 6) * kSlabBytes); bulk_commit(); }
                                    low_dumped = true;
                                }
                            }
                        }
                    }
#ifdef SNB_TC_PROBE
                    if (gi == 1) TC_MARK(52, 0);
#endif
                    tc_fence_before();
#ifdef SNB_TC_PROBE
                    if (gi == 1) TC_MARK(52, 1);
#endif
                    if (!(TC_DBG(A.dbg) & 32)) fence_proxy_async_smem();
#ifdef SNB_TC_PROBE
                    if (gi == 1) TC_MARK(52, 2);
#endif
                    named_bar_sync(1, ET);              // all TMEM reads / A writes / table reads of this GEMM done
                    TC_MARK(gi, 2);
                    if (dump_dst && tid_e == 0) {                // dump (the rest of) the activation tile this GEMM produced (A-tile image = atoms)
                        const uint32_t lo = low_dumped ? (uint32_t)(chunk_n >> 6) * kSlabBytes : 0u;
                        const uint32_t all = (uint32_t)((kind == GK_TRUNK || kind == GK_FEAT) ? P.a_slabs : ((H2 + 63) >> 6)) * kSlabBytes;
                        bulk_s2g(dump_dst + lo, sm.a + lo, all - lo); bulk_commit();
                    }
                    prev_split = low_dumped;
                    if (gi + 1 < P.n_gemms) {
                        const TcGemm& gn = P.g[gi + 1];
                        if (gn.fmt != TF_NONE) table_copy<ET>(sm.tblF, T + gn.tbl_off, gn.N * gn.fmt * 4, tid_e);
                        if (gn.has_vec) table_copy<ET>(sm.tblV, T + gn.vec_off, gn.N * 4, tid_e);
                        if (!early_signaled) signal_ready(0);
                        signal_ready(1);
                    }
                }
                // ---- head outputs of this point: combine the partial dot products of the ES column interleaves.
                //      The activation tile is dead here (every MMA that read it has completed), so it is the scratch. ----
                if (sb && tid_e == 0) bulk_wait_read();
                named_bar_sync(1, ET);
                float* dots = reinterpret_cast<float*>(sm.a) + (size_t)((half * kTile + row) * 8);
                if (half != 0) { dots[0] = sig_dot; dots[1] = beta_dot; dots[2] = rgb0; dots[3] = rgb1; dots[4] = rgb2; dots[5] = sun_dot; }
                named_bar_sync(1, ET);
                if (half == 0 && valid) {
#pragma unroll
                    for (int h2 = 1; h2 < ES; ++h2) {
                        const float* d = reinterpret_cast<const float*>(sm.a) + (size_t)((h2 * kTile + row) * 8);
                        sig_dot += d[0]; beta_dot += d[1]; rgb0 += d[2]; rgb1 += d[3]; rgb2 += d[4]; sun_dot += d[5];
                    }
                    if (NERF) { sun_dot = 0.f; beta_dot = 0.f; }
                    sm.sg[p] = softplus_f(sig_dot + sm.consts[0]);                                    // satnerf.py:183
                    sm.al0[p] = __fsub_rn(__fmul_rn(sigmoid_f(rgb0 + sm.consts[1]), 1.002f), 0.001f);  // :193-195
                    sm.al1[p] = __fsub_rn(__fmul_rn(sigmoid_f(rgb1 + sm.consts[2]), 1.002f), 0.001f);
                    sm.al2[p] = __fsub_rn(__fmul_rn(sigmoid_f(rgb2 + sm.consts[3]), 1.002f), 0.001f);
                    sm.sn[p] = sigmoid_f(sun_dot + sm.consts[4]);                                      // :200
                    sm.bt[p] = P.has_beta ? softplus_f(beta_dot + sm.consts[5]) : 0.f;                 // :205
                    // per-sample outputs that do not depend on the transmittance scan leave from here (128 threads) instead of
                    // from the one warp per ray that composites
                    const size_t gp = (size_t)r0 * S + p;
                    if (A.sigma) A.sigma[gp] = sm.sg[p];
                    if (NERF && A.nerf_rgb) { A.nerf_rgb[gp * 3] = sm.al0[p]; A.nerf_rgb[gp * 3 + 1] = sm.al1[p]; A.nerf_rgb[gp * 3 + 2] = sm.al2[p]; }
                    if (A.sun) A.sun[gp] = sm.sn[p];
                    if (A.beta && P.has_beta) A.beta[gp] = sm.bt[p];
                    if (A.albedo) { A.albedo[gp * 3] = sm.al0[p]; A.albedo[gp * 3 + 1] = sm.al1[p]; A.albedo[gp * 3 + 2] = sm.al2[p]; }
                    if (A.sky) { A.sky[gp * 3] = sm.skyc[rl * 4]; A.sky[gp * 3 + 1] = sm.skyc[rl * 4 + 1]; A.sky[gp * 3 + 2] = sm.skyc[rl * 4 + 2]; }
                }
                named_bar_sync(1, ET);                  // scratch reads done before the next tile's layer 0 overwrites A
            }
            named_bar_sync(1, ET);
            { const bool dbg_on = TC_DBG(blockIdx.x == 0 && tile_counter == 2); TC_MARK(61, 0); }
            // ---- alpha compositing: one warp per ray, transmittance by warp scan (satnerf.py:52-70) ----
            for (int gr = warp - 2; gr < n_rays; gr += EW) {
                const int ray = r0 + gr;
                float carry = 1.f, depth = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
                float xs = 0.f, xa0 = 0.f, xa1 = 0.f, xa2 = 0.f, xb = 0.f, xw = 0.f;      // sum w*{sun, albedo, beta, 1} (aux_sums)
                const float k0 = NERF ? 0.f : sm.skyc[gr * 4], k1 = NERF ? 0.f : sm.skyc[gr * 4 + 1], k2 = NERF ? 0.f : sm.skyc[gr * 4 + 2];
                for (int b0 = 0; b0 < S; b0 += 32) {
                    const int i = b0 + lane, p = gr * S + i;
                    const bool ok = i < S;
                    if (A.t_min > 0.f && carry < A.t_min) {          // early termination (snb_pass_desc.t_min): the tail's weights sum to < t_min
                        if (ok) { const size_t gp = (size_t)ray * S + i; if (A.weights) A.weights[gp] = 0.f; if (A.transparency) A.transparency[gp] = 0.f; }
                        continue;
                    }
                    float alpha = 0.f, q = 1.f, zi = 0.f;
                    if (ok) {
                        zi = sm.z[p];
                        float delta = i < S - 1 ? __fsub_rn(sm.z[p + 1], zi) : 1e10f;
                        float nz = A.noise ? A.noise[(size_t)ray * S + i] * A.noise_std : 0.f;
                        alpha = 1.0f - expf(-delta * fmaxf(sm.sg[p] + nz, 0.f));
                        q = (1.0f - alpha) + 1e-10f;
                    }
                    float incl = q;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) { float o = __shfl_up_sync(~0u, incl, off); if (lane >= off) incl *= o; }
                    float excl = __shfl_up_sync(~0u, incl, 1); if (lane == 0) excl = 1.f;
                    const float Tr = carry * excl, w = alpha * Tr;
                    carry *= __shfl_sync(~0u, incl, 31);
                    if (ok) {
                        const size_t gp = (size_t)ray * S + i;
                        if (A.weights) A.weights[gp] = w;
                        if (A.transparency) A.transparency[gp] = Tr;
                        const float s = NERF ? 1.f : sm.sn[p];                  // nerf: rgb = sum w * rgb_i (nerf.py:128): irradiance factor 1
                        depth = fmaf(w, zi, depth);
                        c0 += w * sm.al0[p] * (s + (1.f - s) * k0);
                        c1 += w * sm.al1[p] * (s + (1.f - s) * k1);
                        c2 += w * sm.al2[p] * (s + (1.f - s) * k2);
                        if (A.aux_sums) { xs = fmaf(w, s, xs); xa0 = fmaf(w, sm.al0[p], xa0); xa1 = fmaf(w, sm.al1[p], xa1); xa2 = fmaf(w, sm.al2[p], xa2);
                                          xb = fmaf(w, sm.bt[p], xb); xw += w; }
                    }
                }
#pragma unroll
                for (int off = 16; off; off >>= 1) {
                    depth += __shfl_xor_sync(~0u, depth, off); c0 += __shfl_xor_sync(~0u, c0, off);
                    c1 += __shfl_xor_sync(~0u, c1, off); c2 += __shfl_xor_sync(~0u, c2, off);
                }
                if (A.aux_sums) {                        // eval_satnerf.py:125-146: sum over samples of w * {sun, albedo, beta, sky}
#pragma unroll
                    for (int off = 16; off; off >>= 1) {
                        xs += __shfl_xor_sync(~0u, xs, off); xa0 += __shfl_xor_sync(~0u, xa0, off); xa1 += __shfl_xor_sync(~0u, xa1, off);
                        xa2 += __shfl_xor_sync(~0u, xa2, off); xb += __shfl_xor_sync(~0u, xb, off); xw += __shfl_xor_sync(~0u, xw, off);
                    }
                    if (lane == 0) {
                        float4* o = reinterpret_cast<float4*>(A.aux_sums + (size_t)ray * 8);      // two 16-byte stores per ray
                        o[0] = make_float4(xs, xa0, xa1, xa2); o[1] = make_float4(xb, xw * k0, xw * k1, xw * k2);
                    }
                }
                if (lane == 0) {
                    if (A.depth) A.depth[ray] = depth;
                    if (A.rgb && NERF) { A.rgb[ray * 3] = c0; A.rgb[ray * 3 + 1] = c1; A.rgb[ray * 3 + 2] = c2; }      // (no clamp in nerf's inference)
                    else if (A.rgb) { A.rgb[ray * 3] = fminf(fmaxf(c0, 0.f), 1.f); A.rgb[ray * 3 + 1] = fminf(fmaxf(c1, 0.f), 1.f); A.rgb[ray * 3 + 2] = fminf(fmaxf(c2, 0.f), 1.f); }
                }
            }
            named_bar_sync(1, ET);                      // group tables are reused by the next group
            { const bool dbg_on = TC_DBG(blockIdx.x == 0 && tile_counter == 2); TC_MARK(61, 1); }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (CG == 2) cluster_sync_all();          // the leader's MMAs read the peer's shared memory: leave together
    if (warp == 1) { if (CG == 2) tmem_dealloc_2cta(tmem, 512); else tmem_dealloc(tmem, 512); }
}

// --------------------------------------------------------------------------------------------------
// host entry points
// --------------------------------------------------------------------------------------------------
static int choose_group(int S) {
    int best = 1; double best_u = 0.0;
    for (int G = 1; G <= kMaxGroupRays; ++G) {
        int pts = G * S; if (pts > kMaxGroupPts) break;
        double u = (double)pts / (double)(((pts + kTile - 1) / kTile) * kTile);
        if (u > best_u + 1e-9) { best_u = u; best = G; }
    }
    return best;
}

int tc_workspace(const FieldLayout& L, const snb_pass_desc* p, bool backward, size_t* bytes) {
    *bytes = 0;
    if (!tc_supported(L, p)) return 0;
    if (backward) return tc_bwd_workspace(L, p, bytes);
    TcProgram P; int nfl = build_program(L, &P);          // (the NO_BETA program is never larger)
    *bytes = (size_t)P.tables_base + (size_t)nfl * 4 + 1024;
    return 0;
}

#ifdef SNB_DEV_BUILD
const DevKnobs& dev_knobs() {
    static const DevKnobs k = [] {
        auto num = [](const char* n) { const char* e = getenv(n); return e ? atoi(e) : 0; };
        DevKnobs d; d.no_early = num("SNB_TC_NO_EARLY"); d.cg = num("SNB_TC_CG"); d.dbg = num("SNB_TC_DBG");
        { const char* e = getenv("SNB_TC_BWD"); d.bwd_off = (e && atoi(e) == 0) ? 1 : 0; }
        d.hang_mirror = getenv("SNB_TC_HANG_MIRROR") ? 1 : 0;
        return d;
    }();
    return k;
}
#endif

static unsigned int* g_hang_pinned = nullptr;
static int hang_mirror_init() {
    if (g_hang_pinned) return 0;
    SNB_CUDA(cudaHostAlloc((void**)&g_hang_pinned, 1024, cudaHostAllocMapped));
    memset(g_hang_pinned, 0xff, 1024);
    unsigned int* dptr = nullptr;
    SNB_CUDA(cudaHostGetDevicePointer((void**)&dptr, g_hang_pinned, 0));
    SNB_CUDA(cudaMemcpyToSymbol(g_hang_host, &dptr, sizeof(dptr)));
    return 0;
}
int tc_debug_hang_info(unsigned int* out) {
    for (int i = 0; i < 192; ++i) out[i] = g_hang_pinned ? g_hang_pinned[i] : 0xffffffffu;
    return 0;
}

int tc_render_forward(const FieldLayout& L, const snb_pass_desc* p, const snb_render_io* io, void* workspace, size_t workspace_bytes,
                      cudaStream_t st) {
    if (!tc_supported(L, p)) return 1;
    static int sm_count = 0, max_smem = 0;
    if (!sm_count) {
        int dev = 0; SNB_CUDA(cudaGetDevice(&dev));
        int major = 0; SNB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
        if (major != 10) SNB_FAIL(-6, "the tensor-core path needs an sm_100 device (compute capability %d.x found)", major);
        SNB_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
        SNB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    if (dev_knobs().hang_mirror) { int rc = hang_mirror_init(); if (rc) return rc; }
    const bool sigma_only = (p->flags & SNB_PASS_SIGMA_ONLY) != 0;
    if (sigma_only) {
        if (L.variant == SNB_NERF) SNB_FAIL(-1, "SNB_PASS_SIGMA_ONLY: s-nerf / sat-nerf only");
        if (io->stash) SNB_FAIL(-1, "SNB_PASS_SIGMA_ONLY is an inference option");
        if (io->rgb || io->albedo || io->sun || io->sky || io->beta || io->aux_sums) SNB_FAIL(-1, "SNB_PASS_SIGMA_ONLY: only depth / weights / transparency / sigma are produced");
    }
    const bool no_beta = ((p->flags & SNB_PASS_NO_BETA) != 0 || sigma_only) && L.variant == SNB_SATNERF;
    if (no_beta && io->stash) SNB_FAIL(-1, "SNB_PASS_NO_BETA is an inference option (a training pass needs the beta head)");
    if (no_beta && io->beta) SNB_FAIL(-1, "SNB_PASS_NO_BETA: io->beta must be NULL");
    TcArgs A; memset(&A, 0, sizeof(A));
    int nfl = build_program(L, &A.prog, no_beta, sigma_only);
    TcProgram& P = A.prog;
    size_t need = (size_t)P.tables_base + (size_t)nfl * 4;
    if (need > workspace_bytes) SNB_FAIL(-4, "tensor-core path: workspace too small (%zu < %zu)", workspace_bytes, need);
    // CTA pairs (cta_group::2) are the default: each CTA streams and buffers only half of every weight tile, so the ring is
    // 4 x 16 KB deep instead of 2 x 32 KB and the L2 -> shared-memory stream per SM halves; a single CTA is bound by the
    // refill latency of its 2-stage ring (profiles/r1_phase_probe.md, mma_ring*_probe.py).  SNB_TC_CG=1 forces single CTAs.
    const int G_ = choose_group(p->n_samples);
    int cg = (p->n_rays + G_ - 1) / G_ >= 2 ? 2 : 1;
    if (p->flags & SNB_PASS_SINGLE_CTA) cg = 1;
    if (dev_knobs().cg == 1 || dev_knobs().cg == 2) cg = dev_knobs().cg;
    size_t fixed = (size_t)P.a_slabs * kSlabBytes + smem_fixed_bytes() + 1024;
    int ns = (int)(((size_t)max_smem - fixed) / (P.stage_bytes / cg)); if (ns > 8) ns = 8;
    if (ns < 2) SNB_FAIL(-6, "tensor-core path: not enough shared memory for the weight ring");
    P.n_stages = ns;
    size_t smem = fixed + (size_t)ns * (P.stage_bytes / cg);

    A.params = io->params; A.rays = io->rays; A.z = io->z_vals; A.t_emb = io->t_emb; A.noise = p->noise_std != 0.f ? io->noise : nullptr;
    A.noise_std = p->noise_std; A.xyz = io->xyz; A.aux = io->aux_dir;
    A.rgb = io->rgb; A.depth = io->depth; A.weights = io->weights; A.transparency = io->transparency; A.albedo = io->albedo;
    A.sun = io->sun; A.sky = io->sky; A.beta = io->beta; A.sigma = io->sigma; A.aux_sums = io->aux_sums; A.t_min = p->t_min; A.nerf_rgb = io->nerf_rgb;
    A.packed = (unsigned char*)workspace;
    A.R = p->n_rays; A.S = p->n_samples; A.ray_cols = p->ray_cols; A.dir_col = p->march_along_sun ? 8 : 3;
    A.G = choose_group(A.S); A.n_groups = (A.R + A.G - 1) / A.G;
    A.stash_base = L.variant == SNB_NERF ? nullptr : (unsigned char*)io->stash;      // (nerf: no tensor-core backward, nothing to stash)
    if (A.stash_base) { int tpg = (A.G * A.S + kTile - 1) / kTile; stash_layout(L, A.n_groups * tpg, tpg, &A.stash); }
    A.dbg = dev_knobs().dbg;

    MiscOffsets M; memset(&M, 0, sizeof(M));
    M.sigma_w = L.sigma.w; M.sigma_b = L.sigma.b; M.rgb0_b = L.rgb0.b; M.rgb2_w = L.rgb2.w; M.rgb2_b = L.rgb2.b;
    M.sun0_w = L.sun[0].w; M.sun0_b = L.sun[0].b; M.sun0_ld = L.sun[0].n_in; M.sun3_w = L.sun[3].w; M.sun3_b = L.sun[3].b;
    M.sky0_w = L.sky0.w; M.sky0_b = L.sky0.b; M.sky2_w = L.sky2.w; M.sky2_b = L.sky2.b;
    M.beta0_w = L.beta0.w; M.beta0_b = L.beta0.b; M.beta0_ld = L.beta0.n_in; M.beta2_w = L.beta2.w; M.beta2_b = L.beta2.b;
    M.rgb0_w = L.rgb0.w; M.rgb0_ld = L.rgb0.n_in;

    if (!p->weights_packed) {      // (caller's promise otherwise: same parameter values, workspace untouched since we packed them)
        tc_pack_kernel<<<dim3(64, P.n_gemms), 256, 0, st>>>(P, io->params, A.packed);
        SNB_CHECK_LAUNCH();
        tc_pack_misc_kernel<<<8, 256, 0, st>>>(P, M, io->params, A.packed);
        SNB_CHECK_LAUNCH();
    }
    const bool train = A.stash_base != nullptr;
    if (cg == 2) {
        auto kern = P.nerf ? tc_render_kernel<2, false, true> : (train ? tc_render_kernel<2, true, false> : tc_render_kernel<2, false, false>);
        SNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int n_pairs = (A.n_groups + 1) / 2, max_pairs = sm_count / 2;
        cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(2 * (n_pairs < max_pairs ? n_pairs : max_pairs)); cfg.blockDim = dim3(64 + 32 * (train ? kEpiWarpsFwdTrain : (P.nerf ? kEpiWarpsTrain : kEpiWarps)));
        cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        SNB_CUDA(cudaLaunchKernelEx(&cfg, kern, A));
        ++g_launches;
    } else {
        auto kern = P.nerf ? tc_render_kernel<1, false, true> : (train ? tc_render_kernel<1, true, false> : tc_render_kernel<1, false, false>);
        SNB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int grid = A.n_groups < sm_count ? A.n_groups : sm_count;
        kern<<<grid, 64 + 32 * (train ? kEpiWarpsFwdTrain : (P.nerf ? kEpiWarpsTrain : kEpiWarps)), smem, st>>>(A);
        SNB_CHECK_LAUNCH();
    }
    return 0;
}

int tc_stash_bytes(const FieldLayout& L, const snb_pass_desc* p, size_t* bytes) {
    *bytes = 0;
    if (!tc_supported(L, p) || !tc_bwd_supported(L, p)) return 0;      // no tensor-core backward for this shape: nothing to stash (the fp32 backward recomputes)
    int G = choose_group(p->n_samples), tpg = (G * p->n_samples + kTile - 1) / kTile, groups = (p->n_rays + G - 1) / G;
    TcStash S; stash_layout(L, groups * tpg, tpg, &S);
    *bytes = (size_t)S.total + 1024;
    return 0;
}

int tc_debug_read(void* dst, size_t bytes) {
    if (bytes > sizeof(long long) * 64 * 4) bytes = sizeof(long long) * 64 * 4;
    SNB_CUDA(cudaMemcpyFromSymbol(dst, g_tc_dbg, bytes));
    return 0;
}



}  // namespace snb
