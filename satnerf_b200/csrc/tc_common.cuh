// Shared tile machinery of the fused tensor-core kernels (forward: tc_field.cu, input-gradient chain: tc_bwd.cu).
#pragma once
#include "common.cuh"
#include "sm100_ptx.cuh"
#include <cstdlib>

namespace snb {

using namespace ptx;

constexpr int kTile = 128;             // points per tile = UMMA M
constexpr int kMaxGemms = 24;
constexpr int kMaxGroupRays = 4;
constexpr int kMaxGroupPts = 384;
constexpr int kSlabBytes = kTile * 128;      // one 64-wide K slab of the A tile
constexpr int kTblF = 8192, kTblV = 2048;    // epilogue table region: [N][4] or [N] floats + extra vector
// Epilogue warps (kEpi / 4 per TMEM lane quadrant).  Inference: 16 (with the producer and issuer warps 18 warps = 5 on two of
// the four sub-partitions: at most 96 registers per thread).  Training kernels (forward with stash, backward chain): 8 --
// 10 warps leave 168 registers per thread; at 96 their epilogues spilled ~50 values and re-derived loop invariants every
// 32-column block (LDL / indexed LDC stalls: measured 1.99 ms per 1024-ray step with 16, 1.70 with 12, 1.64 with 8 warps); the
// inference kernel has no spills at 96 and loses 2.5 % with 12 warps.
#ifndef SNB_EPI_WARPS
#define SNB_EPI_WARPS 16
#endif
#ifndef SNB_EPI_WARPS_TRAIN
#define SNB_EPI_WARPS_TRAIN 8
#endif
constexpr int kEpiWarps = SNB_EPI_WARPS;
constexpr int kEpiWarpsTrain = SNB_EPI_WARPS_TRAIN;
#ifndef SNB_EPI_WARPS_FWD_TRAIN
#define SNB_EPI_WARPS_FWD_TRAIN SNB_EPI_WARPS_TRAIN
#endif
constexpr int kEpiWarpsFwdTrain = SNB_EPI_WARPS_FWD_TRAIN;      // training-mode forward (the chain and the nerf kernel use kEpiWarpsTrain)
constexpr int kMaxEpiWarps = kEpiWarps > kEpiWarpsTrain ? kEpiWarps : kEpiWarpsTrain;

enum { GK_TRUNK = 0, GK_FEAT, GK_HEADA, GK_SUN1, GK_SUN2, GK_SUN3, GK_HEADN };      // HEADN: nerf's rgb_from_xyzdir.0 (per-ray view-direction bias, ReLU)
enum { TF_NONE = 0, TF_F1 = 1, TF_F4 = 4 };

struct TcGemm {
    int kind, N, K, n_chunks, chunk_n, k_slabs, skip, last, fmt, has_vec;
    int free_slabs, store2_idx;      // two-chunk GEMMs that write the tile: K-slabs the last chunk's MMAs release one by one; sequence index
    int aux;                         // 1: bias (2: + the skip layer's xyz term) comes from one extra K-step on the aux operand tiles
    int two_idx;                     // number of two-chunk GEMMs before this one in the program
    int k_early;                     // K-slabs of the input tile that are valid (and chunk-0 accumulator columns free) at the FIRST
                                     // ready signal; the rest needs the second one
    int accumulate;                  // 1: the MMAs add onto what the previous GEMM left in the accumulator (backward chain)
    int a_slab0;                     // first K-slab of the A tile this GEMM reads (nerf layer 0 reads only the positional-encoding slab)
    int relu;                        // activation of a TRUNK GEMM: 0 sin, 1 ReLU (nerf)
    int kvalid, k2_start, k2_cols, col_k2;   // weight columns along K: k < kvalid -> col0 + k; k2_start <= k < k2_start + k2_cols -> col_k2 + k - k2_start; else 0
    int tbl_off, vec_off;            // float offsets into the packed table area
    // weight sources (flat fp32 params): rows [0,rows0) from src0, the rest from src1
    long long src0, src1; int ld0, ld1, col0, col1, rows0;
};

struct TcProgram {
    int H, H2, n_gemms, n_two, n_store2, tau, has_beta, a_slabs, stage_bytes, n_stages;
    int nerf, pe_xyz, pe_dir, pe_slab;           // nerf variant: Mapping frequencies (nerf.py:36-69) and the A-tile slab that holds the encoded position
    int l0_tbl, consts, sunw, betaw, sky;        // float offsets in the table area (nerf: sunw = rgb_from_xyzdir.0 view-direction columns + bias)
    long long l0_w, l0_b;                        // flat param offsets of trunk layer 0
    long long tables_base;                       // byte offset of the table area in the packed buffer
    TcGemm g[kMaxGemms];
};

// Activation stash written by the training-mode forward (byte offsets from the stash base; every array is indexed
// by the global tile id gt = group * tiles_per_group + tile).
//   atoms : [gt][fg = f/64][pg][8 points][64 features] fp16, 128B-swizzled (see tc_backward.cu) — A-tile images
//   yb    : [gt][f/8][128 rows][8] fp16 activation DERIVATIVES w0 cos(w0 y) of the sine layers (what the backward multiplies by)
struct TcStash {
    long long a[kMaxTrunk];                 // a_l = sin(.) outputs of trunk layer l (atoms, H wide)
    long long feat, r1, s1, s2, s3, b1;     // feats_from_xyz output; first-layer activations of the rgb / sun / beta heads
    long long y[kMaxTrunk], r1y, s1y, s2y, s3y, b1y;   // yb arrays
    long long e;                            // atoms, one feature group: [x y z | sun(3) | t_emb(<=4) | 1 | 0...]
    long long total;
    int n_tiles, tiles_per_group;
};

struct TcArgs {
    TcProgram prog;
    TcStash stash; unsigned char* stash_base;    // stash_base == nullptr: inference, nothing is stashed
    const float *params, *rays, *z, *t_emb, *noise, *xyz, *aux;
    float noise_std;
    float *rgb, *depth, *weights, *transparency, *albedo, *sun, *sky, *beta, *sigma, *aux_sums, *nerf_rgb;
    float t_min;
    unsigned char* packed;
    int R, S, ray_cols, dir_col, G, n_groups;
    int dbg;      // developer knobs (env SNB_TC_DBG): 1 = skip sin, 2 = skip TMEM loads, 4 = skip activation stores,
                  // 8 = skip MMA issue, 16 = skip weight copies
};


// Developer knobs.  Product build: compile-time zeros (no getenv anywhere on the call path).  Dev build
// (libsatnerf_b200_dev.so, -DSNB_DEV_BUILD): read ONCE from the environment on first use.
struct DevKnobs { int no_early, cg, dbg, bwd_off, hang_mirror; };
#ifdef SNB_DEV_BUILD
const DevKnobs& dev_knobs();
// what-if knobs of the training kernels (SNB_TC_DBG bits, dev library only; results are wrong by design):
// 128 forward: no pre-activation stash stores, 256 forward: no activation-tile dumps, 512 chain: no pre-activation loads,
// 1024 chain: no dY dumps
#define SNB_DEV_DBG(x) (x)
#else
inline DevKnobs dev_knobs() { return DevKnobs{0, 0, 0, 0, 0}; }
#define SNB_DEV_DBG(x) 0
#endif

struct Smem {
    unsigned char* a;        // activation tile: a_slabs x [128 rows x 128 B], 128B-swizzled, K-major
    unsigned char* b;        // weight ring: n_stages x stage_bytes
    unsigned char* a_aux;    // [128 rows][16 fp16] 32B-swizzled: [1 1 | xyz hi | xyz lo | xyz hi | 0..] (aliases the upper half of tblF)
    float* tblF;             // [N][4] or [N]
    float* tblV;             // [N]
    float *z, *sg, *al0, *al1, *al2, *sn, *bt, *wt;   // per-point tables of the current group (kMaxGroupPts each); wt = xyz scratch of the tile [3][128]
    float* sunb;             // [kMaxGroupRays][H2] per-ray bias of sun_v_net.0 (bias + W[:,H:H+3] sun_d)
    float* betab;            // [kMaxGroupRays][H2] per-ray bias of beta_from_xyz.0
    float* skyc;             // [kMaxGroupRays][4]
    float* consts;           // 8 floats
    uint64_t *full, *empty, *peer_full, *acc_full, *acc_full2, *a_ready, *a_ready2, *slab_free;
    uint32_t* tmem_ptr;
};

__device__ __forceinline__ Smem carve(unsigned char* base, const TcProgram& P, int cg) {
    Smem s; unsigned char* p = base;
    s.a = p; p += (size_t)P.a_slabs * kSlabBytes;
    s.b = p; p += (size_t)P.n_stages * (P.stage_bytes / cg);
    s.tblF = (float*)p; s.a_aux = p + kTblF - kTile * 32; p += kTblF;
    s.tblV = (float*)p; p += kTblV;
    float* f = (float*)p;
    s.z = f; s.sg = f + kMaxGroupPts; s.al0 = f + 2 * kMaxGroupPts; s.al1 = f + 3 * kMaxGroupPts; s.al2 = f + 4 * kMaxGroupPts;
    s.sn = f + 5 * kMaxGroupPts; s.bt = f + 6 * kMaxGroupPts; s.wt = f + 7 * kMaxGroupPts;
    p += 8 * kMaxGroupPts * 4;
    s.sunb = (float*)p; p += kMaxGroupRays * 256 * 4;
    s.betab = (float*)p; p += kMaxGroupRays * 256 * 4;
    s.skyc = (float*)p; p += 64;
    s.consts = (float*)p; p += 32;
    s.full = (uint64_t*)p; p += 8 * 8;
    s.empty = (uint64_t*)p; p += 8 * 8;
    s.peer_full = (uint64_t*)p; p += 8 * 8;
    s.acc_full = (uint64_t*)p; p += 8;
    s.acc_full2 = (uint64_t*)p; p += 8;
    s.a_ready = (uint64_t*)p; p += 8;
    s.a_ready2 = (uint64_t*)p; p += 8;
    s.slab_free = (uint64_t*)p; p += 4 * 8;
    s.tmem_ptr = (uint32_t*)p;
    return s;
}

// bytes of one N-chunk of a GEMM in the packed weight stream: k_slabs 128B-swizzled tiles [+ one 32B-swizzled aux tile]
__host__ __device__ __forceinline__ long long chunk_stream_bytes(const TcGemm& g) { return (long long)g.k_slabs * g.chunk_n * 128 + (g.aux ? g.chunk_n * 32 : 0); }
__host__ __device__ __forceinline__ long long gemm_stream_bytes(const TcGemm& g) { return g.n_chunks * chunk_stream_bytes(g); }

__device__ __forceinline__ int group_points(const TcArgs& A, int grp) {
    int r0 = grp * A.G; int n = A.R - r0; if (n > A.G) n = A.G; return n * A.S;
}

// byte address of the 16-byte chunk holding k..k+7 (k % 8 == 0) of `row` in the swizzled A tile
__device__ __forceinline__ uint32_t a_chunk_addr(uint32_t a_base, int row, int k) {
    return a_base + (uint32_t)(k >> 6) * kSlabBytes + (uint32_t)row * 128u + ((((uint32_t)(k >> 3) & 7u) ^ ((uint32_t)row & 7u)) << 4);
}
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// store 32 consecutive activations (k0 % 32 == 0) of one row as fp16
__device__ __forceinline__ void store_act32(uint32_t a_base, int row, int k0, const float* v) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float* x = v + c * 8;
        sts128(a_chunk_addr(a_base, row, k0 + c * 8), pack_half2(x[0], x[1]), pack_half2(x[2], x[3]), pack_half2(x[4], x[5]), pack_half2(x[6], x[7]));
    }
}
// NC (multiple of 8) consecutive activations of one row as fp16 into the swizzled A tile
template <int NC>
__device__ __forceinline__ void store_act_cols(uint32_t a_base, int row, int k0, const float* v) {
#pragma unroll
    for (int c = 0; c < NC / 8; ++c) {
        const float* x = v + c * 8;
        sts128(a_chunk_addr(a_base, row, k0 + c * 8), pack_half2(x[0], x[1]), pack_half2(x[2], x[3]), pack_half2(x[4], x[5]), pack_half2(x[6], x[7]));
    }
}
// ---- training stash helpers -------------------------------------------------------------------------------
// The forward stashes the derivative of every sine activation, w0 cos(w0 y) (w0 = 30 for trunk layer 0, 1 elsewhere), as fp16: the
// range reduction is shared with the sine, the value is in [-30, 30] whatever |y| is (relative error 2^-11), and the backward's
// epilogue becomes one packed half multiply per pair of elements -- no conversion, no range reduction, no MUFU.
// yb arrays: [tile gt][8-column block n/8][row 0..127][8 fp16] -- the 32 lanes of a warp (consecutive rows) write / read 512
// contiguous bytes per 16-byte access (the round-1 layout [n/32][row][32] made every access a 64-byte-strided scatter and
// the training forward 3x slower than inference).
__device__ __forceinline__ unsigned char* yb_chunk(unsigned char* arr, int gt, int F, int n, int row) {      // n % 8 == 0
    return arr + (((size_t)gt * (F >> 3) + (n >> 3)) * kTile + row) * 16;
}
__device__ __forceinline__ uint4 yb_pack8(const float* x) {
    return make_uint4(pack_half2(x[0], x[1]), pack_half2(x[2], x[3]), pack_half2(x[4], x[5]), pack_half2(x[6], x[7]));
}
// NC (multiple of 8) consecutive pre-activations of one row starting at column n0
template <int NC>
__device__ __forceinline__ void yb_store_cols(unsigned char* arr, int gt, int F, int n0, int row, const float* y) {
#pragma unroll
    for (int c = 0; c < NC / 8; ++c) *reinterpret_cast<uint4*>(yb_chunk(arr, gt, F, n0 + c * 8, row)) = yb_pack8(y + c * 8);
}
// address of the 16-byte chunk (features k..k+7, k % 8 == 0) of point `row` of tile gt in an atoms array with `fgs` groups
__device__ __forceinline__ unsigned char* atom_chunk(unsigned char* arr, int gt, int fgs, int row, int k) {
    return arr + (((size_t)gt * fgs + (k >> 6)) * 16 + (row >> 3)) * 1024 + (size_t)(row & 7) * 128 + (size_t)((((k >> 3) & 7) ^ (row & 7)) << 4);
}
template <int NC>
__device__ __forceinline__ void atom_store_cols(unsigned char* arr, int gt, int fgs, int row, int k0, const float* v) {
#pragma unroll
    for (int c = 0; c < NC / 8; ++c) {
        const float* x = v + c * 8;
        *reinterpret_cast<uint4*>(atom_chunk(arr, gt, fgs, row, k0 + c * 8)) =
            make_uint4(pack_half2(x[0], x[1]), pack_half2(x[2], x[3]), pack_half2(x[4], x[5]), pack_half2(x[6], x[7]));
    }
}
__device__ __forceinline__ void atom_store32(unsigned char* arr, int gt, int fgs, int row, int k0, const float* v) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float* x = v + c * 8;
        *reinterpret_cast<uint4*>(atom_chunk(arr, gt, fgs, row, k0 + c * 8)) =
            make_uint4(pack_half2(x[0], x[1]), pack_half2(x[2], x[3]), pack_half2(x[4], x[5]), pack_half2(x[6], x[7]));
    }
}

__device__ __forceinline__ float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

// cooperative copy (all epilogue threads) of `bytes` (multiple of 16) from global to shared with cp.async
template <int ET>
__device__ __forceinline__ void table_copy(void* dst, const void* src, int bytes, int tid_e) {
    for (int o = tid_e * 16; o < bytes; o += ET * 16) cp_async16((char*)dst + o, (const char*)src + o);
}


// Shared-memory table read that the compiler may schedule freely (no "memory" clobber, so reads of a block
// pipeline instead of paying the LDS latency one by one).  `tok` is a value produced by fresh_token() AFTER the
// barrier that published the table: the data dependence keeps the read below that barrier and keeps reads of
// different table generations (same address, next layer) from being merged.
__device__ __forceinline__ float4 lds128(uint32_t addr, uint32_t tok) {
    float4 r;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+0];  // gen %5" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr), "r"(tok));
    return r;
}
__device__ __forceinline__ uint32_t fresh_token(uint32_t x) {
    uint32_t t;
    asm volatile("mov.u32 %0, %1;" : "=r"(t) : "r"(x) : "memory");
    return t;
}

}  // namespace snb
