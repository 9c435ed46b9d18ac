// Depth samplers of render_rays: stratified bins (rendering.py:65-78) and inverse-CDF importance
// sampling with the sorted merge (rendering.py:10-49, :121-125).  All arithmetic is fp32 with each
// operation rounded separately (no FMA contraction), like the reference's chain of eager torch ops,
// so depths and bin indices can be compared bit for bit with the CPU oracle.
#include "common.cuh"

namespace snb {

__global__ void stratified_kernel(const float* __restrict__ rays, int ray_cols, const float* __restrict__ steps,
                                  const float* __restrict__ u, float* __restrict__ z, int R, int S) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R * S) return;
    int r = idx / S, i = idx - r * S;
    float near = rays[(size_t)r * ray_cols + 6], far = rays[(size_t)r * ray_cols + 7];
    auto zc = [&](int j) {  // z = near*(1-t) + far*t                         (rendering.py:67)
        float t = steps[j];
        return __fadd_rn(__fmul_rn(near, __fsub_rn(1.0f, t)), __fmul_rn(far, t));
    };
    float zi = zc(i);
    float lower = i > 0 ? __fmul_rn(0.5f, __fadd_rn(zc(i - 1), zi)) : zi;          // :72,:75
    float upper = i < S - 1 ? __fmul_rn(0.5f, __fadd_rn(zi, zc(i + 1))) : zi;      // :72,:74
    z[idx] = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), u[idx]));         // :77-78
}

// searchsorted(cdf, u, right=True): number of cdf entries <= u                     (rendering.py:36)
__device__ __forceinline__ int upper_bound(const float* cdf, int n, float u) {
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (cdf[mid] <= u) lo = mid + 1; else hi = mid; }
    return lo;
}

__global__ void searchsorted_kernel(const float* __restrict__ cdf, const float* __restrict__ u, int64_t* __restrict__ inds,
                                    int R, int n_cdf, int n_u) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= R * n_u) return;
    int r = idx / n_u;
    inds[idx] = upper_bound(cdf + (size_t)r * n_cdf, n_cdf, u[idx]);
}

// One warp per ray.  Shared memory per warp: bins[S-1], cdf[S-1], zcat[S+N].
__global__ void importance_kernel(const float* __restrict__ zc, const float* __restrict__ wc, const float* __restrict__ u,
                                  float* __restrict__ z_out, int64_t* __restrict__ inds_out, float* __restrict__ z_new_out,
                                  float* __restrict__ cdf_out, int R, int S, int N) {
    extern __shared__ float sm[];
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int per_warp = 2 * (S - 1) + (S + N);
    float* bins = sm + (size_t)warp * per_warp;
    float* cdf = bins + (S - 1);
    float* zcat = cdf + (S - 1);
    const int M = S - 2;                       // number of weights = N_samples_      (rendering.py:22)
    for (int r = blockIdx.x * warps + warp; r < R; r += gridDim.x * warps) {
        const float* z = zc + (size_t)r * S;
        const float* w = wc + (size_t)r * S;
        for (int i = lane; i < S - 1; i += 32) bins[i] = __fmul_rn(0.5f, __fadd_rn(z[i], z[i + 1]));   // :121
        for (int i = lane; i < S; i += 32) zcat[i] = z[i];
        if (lane == 0) {
            // weights = w[1:-1] + eps (:23); pdf = weights / sum (:24); cdf = [0, cumsum(pdf)] (:25-26).
            // torch's CPU cumsum carries a double accumulator and rounds every prefix to fp32; the row sum
            // is also formed in double here (torch's vectorised fp32 cascade is not reproducible, <=1 ulp).
            double tot = 0.0;
            for (int i = 0; i < M; ++i) tot += (double)__fadd_rn(w[i + 1], 1e-5f);
            float totf = (float)tot;
            double acc = 0.0;
            cdf[0] = 0.0f;
            for (int i = 0; i < M; ++i) {
                acc += (double)__fdiv_rn(__fadd_rn(w[i + 1], 1e-5f), totf);
                cdf[i + 1] = (float)acc;
            }
        }
        __syncwarp();
        if (cdf_out) for (int i = lane; i < S - 1; i += 32) cdf_out[(size_t)r * (S - 1) + i] = cdf[i];
        for (int j = lane; j < N; j += 32) {
            float uj = u[(size_t)r * N + j];
            int k = upper_bound(cdf, M + 1, uj);                                  // :36
            int below = max(k - 1, 0), above = min(k, M);                          // :37-38
            float cb = cdf[below], ca = cdf[above], bb = bins[below], ba = bins[above];
            float den = __fsub_rn(ca, cb);                                         // :44
            if (den < 1e-5f) den = 1.0f;                                           // :45
            float s = __fadd_rn(bb, __fmul_rn(__fdiv_rn(__fsub_rn(uj, cb), den), __fsub_rn(ba, bb)));   // :48
            zcat[S + j] = s;
            if (inds_out) inds_out[(size_t)r * N + j] = k;
            if (z_new_out) z_new_out[(size_t)r * N + j] = s;
        }
        __syncwarp();
        // torch.sort(cat([z, z_new])) (:125) as a rank sort; ties broken by position (stable).
        const int T = S + N;
        for (int e = lane; e < T; e += 32) {
            float v = zcat[e]; int rank = 0;
            for (int k = 0; k < T; ++k) { float o = zcat[k]; rank += (o < v) || (o == v && k < e); }
            z_out[(size_t)r * T + rank] = v;
        }
        __syncwarp();
    }
}

}  // namespace snb

using namespace snb;

extern "C" SNB_API int snb_stratified_depths(const float* rays, int ray_cols, const float* steps, const float* u, float* z,
                                     int R, int S, void* stream) {
    if (R < 0 || S < 1 || ray_cols < 8) SNB_FAIL(-1, "snb_stratified_depths: bad shape R=%d S=%d cols=%d", R, S, ray_cols);
    if (R == 0) return 0;                       // empty batch: nothing to do (pointers may be null)
    if (!rays || !steps || !u || !z) SNB_FAIL(-1, "snb_stratified_depths: null pointer");
    int n = R * S;
    stratified_kernel<<<ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(rays, ray_cols, steps, u, z, R, S);
    SNB_CHECK_LAUNCH();
    return 0;
}

extern "C" SNB_API int snb_searchsorted_right(const float* cdf, const float* u, int64_t* inds, int R, int n_cdf, int n_u, void* stream) {
    if (R < 0 || n_cdf < 1 || n_u < 0) SNB_FAIL(-1, "snb_searchsorted_right: bad shape");
    if (R * n_u == 0) return 0;
    if (!cdf || !u || !inds) SNB_FAIL(-1, "snb_searchsorted_right: null pointer");
    searchsorted_kernel<<<ceil_div(R * n_u, 256), 256, 0, (cudaStream_t)stream>>>(cdf, u, inds, R, n_cdf, n_u);
    SNB_CHECK_LAUNCH();
    return 0;
}

extern "C" SNB_API int snb_importance_depths(const float* z_coarse, const float* weights_coarse, const float* u, float* z_out,
                                     int64_t* inds, float* z_new, float* cdf, int R, int S, int N, void* stream) {
    if (R < 0 || S < 3 || N < 1) SNB_FAIL(-1, "snb_importance_depths: bad shape R=%d S=%d N=%d", R, S, N);
    if (R == 0) return 0;
    if (!z_coarse || !weights_coarse || !u || !z_out) SNB_FAIL(-1, "snb_importance_depths: null pointer");
    const int warps = 4;
    size_t smem = (size_t)warps * (2 * (S - 1) + S + N) * sizeof(float);
    if (smem > 200 * 1024) SNB_FAIL(-1, "snb_importance_depths: S+N too large for shared memory");
    if (smem > 48 * 1024) SNB_CUDA(cudaFuncSetAttribute(importance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int blocks = ceil_div(R, warps); if (blocks > 148 * 16) blocks = 148 * 16;
    importance_kernel<<<blocks, warps * 32, smem, (cudaStream_t)stream>>>(z_coarse, weights_coarse, u, z_out, inds, z_new, cdf, R, S, N);
    SNB_CHECK_LAUNCH();
    return 0;
}

// ---- per-ray batch gather of the GPU-resident ray sampler --------------------------------------------------
// out_t[i, :] = table_t[idx[i], :] for up to 4 row-major tables sharing one index vector (the reference's
// DataLoader.__getitem__ + collate over all_rays / all_rgbs | all_depths / all_ids, datasets/satellite.py:347-350), one launch:
// a thread per 4-byte word of the concatenated output row.
namespace snb {
struct GatherArgs { const unsigned int* src[4]; unsigned int* dst[4]; int words[4], first[5]; int n_tables; const long long* idx; long long n_rows, n_src_rows; };
__global__ void gather_rows_kernel(const __grid_constant__ GatherArgs A) {
    const int wpr = A.first[A.n_tables];                   // words per concatenated row
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < A.n_rows * wpr; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / wpr; const int w = (int)(e - r * wpr);
        long long s = A.idx[r];
        if (s < 0) s += A.n_src_rows;                      // (torch indexing semantics for negative indices)
        int t = 0;
#pragma unroll
        for (int k = 1; k < 4; ++k) if (k < A.n_tables && w >= A.first[k]) t = k;
        const int c = w - A.first[t];
        A.dst[t][r * A.words[t] + c] = A.src[t][s * A.words[t] + c];
    }
}
}  // namespace snb

extern "C" SNB_API int snb_gather_rows(const void* const* tables, void* const* outs, const int32_t* row_bytes, int n_tables,
                                       const int64_t* idx, long long n_rows, long long n_src_rows, void* stream) {
    using namespace snb;
    if (!tables || !outs || !row_bytes || n_tables < 1 || n_tables > 4 || n_rows < 0 || (!idx && n_rows)) SNB_FAIL(-1, "snb_gather_rows: bad argument");
    if (n_rows == 0) return 0;
    GatherArgs A; memset(&A, 0, sizeof(A));
    int first = 0;
    for (int t = 0; t < n_tables; ++t) {
        if (!tables[t] || !outs[t] || row_bytes[t] < 4 || row_bytes[t] % 4) SNB_FAIL(-1, "snb_gather_rows: table %d: null pointer or row size not a multiple of 4 bytes", t);
        A.src[t] = (const unsigned int*)tables[t]; A.dst[t] = (unsigned int*)outs[t]; A.words[t] = row_bytes[t] / 4; A.first[t] = first; first += A.words[t];
    }
    A.first[n_tables] = first; A.n_tables = n_tables; A.idx = (const long long*)idx; A.n_rows = n_rows; A.n_src_rows = n_src_rows;
    long long total = n_rows * first; int blocks = (int)((total + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8;
    gather_rows_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(A);
    SNB_CHECK_LAUNCH();
    return 0;
}
