// fp32 CUDA-core ("SIMT") evaluation of the three fields and of their parameter gradients.
// This is the exactness path (SNB_FP32_SIMT): every contraction is an fp32 FFMA chain, every
// activation the full-precision libdevice function, so results track the reference's fp32 torch ops
// to rounding level.  It also serves as the on-device cross-check of the tcgen05 path at full size.
//
// Restates SatNeRF.forward (models/satnerf.py:156-208), ShadowNeRF.forward (models/snerf.py:148-196),
// NeRF.forward + Mapping (models/nerf.py:184-227, :36-69); the backward is what autograd derives.
#include "simt_field.cuh"
#include "sm100_ptx.cuh"

namespace snb {

// ------------------------------------------------------------------------------------------------
// small elementwise kernels
// ------------------------------------------------------------------------------------------------
// xyz = o + dir*z, products and sums rounded separately like the eager torch ops (rendering.py:81/:104)
__global__ void points_kernel(const float* __restrict__ rays, int ray_cols, int dir_col, const float* __restrict__ z,
                              float* __restrict__ xyz, int r0, int n_rays, int S) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_rays * S) return;
    int r = r0 + idx / S;
    const float* ray = rays + (size_t)r * ray_cols;
    float zz = z[(size_t)r0 * S + idx];
#pragma unroll
    for (int c = 0; c < 3; ++c) xyz[(size_t)idx * 3 + c] = __fadd_rn(ray[c], __fmul_rn(ray[dir_col + c], zz));
}

// Mapping.forward (models/nerf.py:53-69): out = [sin(2^k x), cos(2^k x)]_k, x itself excluded.
__global__ void pe_kernel(const float* __restrict__ x, int ldx, float* __restrict__ out, int n_rows, int n_freqs) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    int per = 6 * n_freqs;
    if (idx >= n_rows * per) return;
    int row = idx / per, j = idx - row * per;
    int k = j / 6, rem = j - k * 6, c = rem % 3;
    float v = __fmul_rn((float)(1 << k), x[(size_t)row * ldx + c]);
    out[idx] = rem < 3 ? sinf(v) : cosf(v);
}

__global__ void ray_sum_kernel(const float* __restrict__ per_point, float* __restrict__ per_ray, int n_rays, int S, int D) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_rays * D) return;
    int r = idx / D, d = idx - r * D;
    float acc = 0.f;
    for (int i = 0; i < S; ++i) acc += per_point[((size_t)r * S + i) * D + d];
    per_ray[idx] = acc;
}

__global__ void reduce_partials_kernel(float* __restrict__ dst, const float* __restrict__ part, int Z, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float acc = 0.f;
    for (int z = 0; z < Z; ++z) acc += part[(int64_t)z * n + i];
    dst[i] += acc;
}

// ------------------------------------------------------------------------------------------------
// activations
// ------------------------------------------------------------------------------------------------
template <int ACT> __device__ __forceinline__ float act_fwd(float x) {
    if (ACT == ACT_SIN) return sinf(x);
    if (ACT == ACT_SIN30) return sinf(__fmul_rn(30.0f, x));                 // Siren(w0=30), nerf.py:33
    if (ACT == ACT_RELU) return fmaxf(x, 0.f);
    if (ACT == ACT_SIGMOID) return 1.0f / (1.0f + expf(-x));
    if (ACT == ACT_SOFTPLUS) return x > 20.f ? x : log1pf(expf(x));         // torch Softplus(beta=1, threshold=20)
    if (ACT == ACT_SIGMOID_PAD) return __fsub_rn(__fmul_rn(1.0f / (1.0f + expf(-x)), 1.002f), 0.001f);   // satnerf.py:195
    return x;
}
template <int ACT> __device__ __forceinline__ float act_bwd(float pre) {   // d act / d pre
    if (ACT == ACT_SIN) return cosf(pre);
    if (ACT == ACT_SIN30) return 30.0f * cosf(__fmul_rn(30.0f, pre));
    if (ACT == ACT_RELU) return pre > 0.f ? 1.f : 0.f;
    return 1.f;
}

__device__ __forceinline__ float src_at(const Src& a0, const Src& a1, int m, int k) {
    if (k < a0.k) return a0.p[(size_t)(m / a0.div) * a0.ld + k];
    return a1.p[(size_t)(m / a1.div) * a1.ld + (k - a0.k)];
}

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

// ------------------------------------------------------------------------------------------------
// Y = act(cat(A0,A1) W^T + b)        A: (M,K) virtual, W: (N,K) row-major with leading dim ldw
// ------------------------------------------------------------------------------------------------
template <int ACT>
__global__ void __launch_bounds__(NT) linear_fwd_kernel(LinFwd p) {
    __shared__ float As[BK][BM + 4], Ws[BK][BN + 4];
    const int K = p.a0.k + p.a1.k;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int t = threadIdx.x, ty = t / 16, tx = t % 16;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            int idx = t + e * NT, row = idx / BK, kk = idx % BK;
            int m = m0 + row, k = k0 + kk;
            As[kk][row] = (m < p.M && k < K) ? src_at(p.a0, p.a1, m, k) : 0.f;
            int n = n0 + row;
            Ws[kk][row] = (n < p.N && k < K) ? p.W[(size_t)n * p.ldw + k] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[4], w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = As[kk][ty * 4 + i]; w[i] = Ws[kk][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n >= p.N) continue;
            float y = acc[i][j] + p.b[n];
            if (p.pre) p.pre[(size_t)m * p.N + n] = y;
            p.out[(size_t)m * p.ldo + n] = act_fwd<ACT>(y);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// dX[m,k] = (sum_n dY[m,n] W[n,k_off+k] (+ dX_old)) * act'(pre[m,k])
// ------------------------------------------------------------------------------------------------
template <int ACT>
__global__ void __launch_bounds__(NT) linear_bwd_in_kernel(LinBwdIn p) {
    __shared__ float Ys[BK][BM + 4], Ws[BK][BN + 4];
    const int m0 = blockIdx.x * BM, c0 = blockIdx.y * BN;
    const int t = threadIdx.x, ty = t / 16, tx = t % 16;
    float acc[4][4] = {};
    for (int n0 = 0; n0 < p.N; n0 += BK) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            int idx = t + e * NT;
            { int row = idx / BK, nn = idx % BK; int m = m0 + row, n = n0 + nn;
              Ys[nn][row] = (m < p.M && n < p.N) ? p.dY[(size_t)m * p.ldy + n] : 0.f; }
            { int nn = idx / BN, col = idx % BN; int n = n0 + nn, k = c0 + col;
              Ws[nn][col] = (n < p.N && k < p.K) ? p.W[(size_t)n * p.ldw + p.k_off + k] : 0.f; }
        }
        __syncthreads();
#pragma unroll
        for (int nn = 0; nn < BK; ++nn) {
            float a[4], w[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = Ys[nn][ty * 4 + i]; w[i] = Ws[nn][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int m = m0 + ty * 4 + i;
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int k = c0 + tx * 4 + j;
            if (k >= p.K) continue;
            float v = acc[i][j];
            size_t o = (size_t)m * p.ldx + k;
            if (p.accumulate) v += p.dX[o];
            if (ACT != ACT_NONE) v *= act_bwd<ACT>(p.pre[(size_t)m * p.K + k]);
            p.dX[o] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// partial[z][n][k] = sum_{m in slice z} dY[m,n] X[m,k];  partial bias appended after the N*K block
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) linear_bwd_w_kernel(LinBwdW p) {
    __shared__ float Ys[BK][BM + 4], Xs[BK][BN + 4];
    const int K = p.a0.k + p.a1.k;
    const int k0 = blockIdx.x * BN, n0 = blockIdx.y * BM, z = blockIdx.z;
    const int t = threadIdx.x, ty = t / 16, tx = t % 16;
    const int m_beg = z * p.m_per_z, m_end = min(p.M, m_beg + p.m_per_z);
    float acc[4][4] = {}; float bacc[4] = {};
    for (int mb = m_beg; mb < m_end; mb += BK) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            int idx = t + e * NT, mm = idx / BM, col = idx % BM;
            int m = mb + mm;
            int n = n0 + col, k = k0 + col;
            Ys[mm][col] = (m < m_end && n < p.N) ? p.dY[(size_t)m * p.ldy + n] : 0.f;
            Xs[mm][col] = (m < m_end && k < K) ? src_at(p.a0, p.a1, m, k) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int mm = 0; mm < BK; ++mm) {
            float a[4], x[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { a[i] = Ys[mm][ty * 4 + i]; x[i] = Xs[mm][tx * 4 + i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (tx == 0) bacc[i] += a[i];
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], x[j], acc[i][j]);
            }
        }
        __syncthreads();
    }
    float* part = p.partial + (size_t)z * ((size_t)p.N * K + p.N);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int n = n0 + ty * 4 + i;
        if (n >= p.N) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) { int k = k0 + tx * 4 + j; if (k < K) part[(size_t)n * K + k] = acc[i][j]; }
        if (tx == 0 && blockIdx.x == 0) part[(size_t)p.N * K + n] = bacc[i];
    }
}

// ------------------------------------------------------------------------------------------------
// Narrow heads (N <= 4: sigma, rgb, sun, sky, beta outputs): one warp per 4 rows, each row read once with coalesced 16-byte
// loads (4 independent loads in flight per lane), the weights in shared memory, the 4 x N dot products reduced with shuffles.
// fp32 FFMA like the tiled kernel (which would idle 60 of its 64 tile columns here: 43-66 us per head against ~8).
// ------------------------------------------------------------------------------------------------
constexpr int NARROW_MAX_N = 4, NARROW_MAX_K = 1024;
template <int ACT>
__global__ void __launch_bounds__(256) linear_fwd_narrow_kernel(LinFwd p) {
    __shared__ __align__(16) float Ws[NARROW_MAX_N * NARROW_MAX_K];
    const int K = p.a0.k + p.a1.k;
    for (int n = 0; n < NARROW_MAX_N; ++n)
        for (int k = threadIdx.x; k < K; k += 256) Ws[n * NARROW_MAX_K + k] = n < p.N ? p.W[(size_t)n * p.ldw + k] : 0.f;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool plain = p.a1.k == 0 && p.a0.div == 1;                              // (every head of the three fields: one row-major source)
    const bool vec4 = plain && (K & 3) == 0 && (p.a0.ld & 3) == 0 && (((uintptr_t)p.a0.p & 15u) == 0);
    constexpr int RW = 4;                                                         // rows per warp task: RW independent loads in flight per lane
    for (int m0 = (blockIdx.x * 8 + warp) * RW; m0 < p.M; m0 += gridDim.x * 8 * RW) {
        float acc[RW][NARROW_MAX_N] = {};
        if (vec4) {
            const float4* row[RW];
#pragma unroll
            for (int r = 0; r < RW; ++r) row[r] = reinterpret_cast<const float4*>(p.a0.p + (size_t)(m0 + r < p.M ? m0 + r : m0) * p.a0.ld);
            for (int k4 = lane; k4 < K / 4; k4 += 32) {
                float4 a[RW];
#pragma unroll
                for (int r = 0; r < RW; ++r) a[r] = row[r][k4];
#pragma unroll
                for (int n = 0; n < NARROW_MAX_N; ++n) {
                    const float4 w = *reinterpret_cast<const float4*>(&Ws[n * NARROW_MAX_K + 4 * k4]);
#pragma unroll
                    for (int r = 0; r < RW; ++r) acc[r][n] = fmaf(a[r].x, w.x, fmaf(a[r].y, w.y, fmaf(a[r].z, w.z, fmaf(a[r].w, w.w, acc[r][n]))));
                }
            }
        } else {
            for (int k = lane; k < K; k += 32) {
#pragma unroll
                for (int r = 0; r < RW; ++r) {
                    const int m = m0 + r < p.M ? m0 + r : m0;
                    const float a = plain ? p.a0.p[(size_t)m * p.a0.ld + k] : src_at(p.a0, p.a1, m, k);
#pragma unroll
                    for (int n = 0; n < NARROW_MAX_N; ++n) acc[r][n] = fmaf(a, Ws[n * NARROW_MAX_K + k], acc[r][n]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < RW; ++r)
#pragma unroll
            for (int n = 0; n < NARROW_MAX_N; ++n)
#pragma unroll
                for (int o = 16; o; o >>= 1) acc[r][n] += __shfl_xor_sync(0xffffffffu, acc[r][n], o);
        // lane 4 r + n writes output n of row m0 + r
        if (lane < RW * NARROW_MAX_N) {
            const int r = lane >> 2, n = lane & 3, m = m0 + r;
            float y = 0.f;
#pragma unroll
            for (int rr = 0; rr < RW; ++rr)
#pragma unroll
                for (int nn = 0; nn < NARROW_MAX_N; ++nn) if (rr * NARROW_MAX_N + nn == lane) y = acc[rr][nn];
            if (m < p.M && n < p.N) {
                y += p.b[n];
                if (p.pre) p.pre[(size_t)m * p.N + n] = y;
                p.out[(size_t)m * p.ldo + n] = act_fwd<ACT>(y);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// SNB_FP16X3_TC: the same Y = act(cat(A0,A1) W^T + b) on the tensor cores at (almost) fp32 operand precision.
// Every fp32 operand is split into fp16 hi + lo (x = hi + lo to 2^-22 |x|, or 2^-25 absolute below the fp16 normal range) and
// the contraction is issued as three tcgen05 MMAs per K-step -- A_hi W_hi + A_lo W_hi + A_hi W_lo, fp32 accumulation in TMEM
// (the lo x lo term is 2^-22 of the product and dropped).  One CTA = 128 rows x `nt` <= 256 columns.  K runs over the two
// sources one after the other (each padded to whole 64-wide slabs, so the loads of the wide source stay aligned whatever the
// width of the narrow one).
//   weights: split ONCE per pass by x3_pack_kernel into ready-made 128B-swizzled hi / lo tiles (scaled per row by 2^e) in the caller's
//            workspace; a producer warp brings them in with two bulk copies per slab (2-stage ring, mbarrier transaction counts);
//   activations: split on the fly -- 16 warps read the fp32 rows (one warp instruction = one 256-byte row segment, two floats
//            per lane; 8 rows per thread and slab), the loads running two slabs ahead of the split in registers;
//   an issuer warp issues the 12 MMAs of a slab once the 16 converter warps have arrived on the slab's mbarrier and the weight
//   tiles have landed; nobody meets at a CTA-wide barrier inside the K loop, and the next slab is prepared while the MMAs run.
// The epilogue transposes the accumulator through shared memory (32 x 32 blocks per warp) so that rows leave as 128-byte stores.
// For deep optical depths (trained scenes) the plain fp16-operand kernel reaches 1e-2 on the weights; this path
// stays at the fp32 level (the tensor core's truncating accumulation leaves ~4e-6 at K = 512).
// ------------------------------------------------------------------------------------------------
using namespace ptx;
constexpr int X3_NT = 256;                      // columns per CTA at most (UMMA N)
constexpr int X3_STAGE = 2 * 16384 + 2 * X3_NT * 128;      // A_hi | A_lo | W_hi | W_lo
constexpr int X3_THREADS = 512;                 // 16 converter / epilogue warps (the split stream is latency-bound with fewer) + producer warp + issuer warp
// Weight rows are pre-scaled by a power of two chosen per output row (2^e with max_k |2^e W[n][k]| in [2048, 4096): exact, no
// fp16 overflow whatever the weights' magnitude, and SIREN-scale weights (|w| ~ 4e-3 at h = 512) keep a NORMAL fp16 lo part); the
// epilogue multiplies column n of the accumulator by 2^-e (a float per column stored behind the layer's tiles).

__device__ __forceinline__ void sts_b32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void x3_split(float x0, float x1, uint32_t* hi, uint32_t* lo) {
    const __half2 hh = __floats2half2_rn(x0, x1);
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
    *hi = *reinterpret_cast<const uint32_t*>(&hh); *lo = *reinterpret_cast<const uint32_t*>(&ll);
}

// sin(y) for the epilogue of this path: k = rint(y / pi) by the magic-number add, three-term Cody-Waite reduction to
// r = y - k pi in [-pi/2, pi/2], odd degree-11 minimax polynomial, sign from k's parity -- 14 instructions against ~27 of
// libdevice's sinf (which selects between a sine and a cosine polynomial on [-pi/4, pi/4]).  Max abs error 1.2e-7 for |y| <= 8000
// (libdevice: 0.7e-7; checked against float64 on 2 M points per range); beyond that libdevice's sinf.
__device__ __forceinline__ float x3_sin(float y) {
    if (fabsf(y) > 8000.f) return sinf(y);
    const float t = fmaf(y, 0.318309886183790672f, 12582912.f);       // 1.5 * 2^23: the integer k sits in the low mantissa bits
    const float k = t - 12582912.f;
    float r = fmaf(k, -3.140625f, y);
    r = fmaf(k, -9.67502593994140625e-4f, r);
    r = fmaf(k, -1.5099580252808664e-7f, r);
    const float r2 = r * r;
    float q = fmaf(-2.3846693508744465e-08f, r2, 2.752261934801936e-06f);
    q = fmaf(q, r2, -0.00019840804452542216f);
    q = fmaf(q, r2, 0.008333330042660236f);
    q = fmaf(q, r2, -0.1666666716337204f);
    const float s = fmaf(r * r2, q, r);
    return __uint_as_float(__float_as_uint(s) ^ (__float_as_uint(t) << 31));
}
template <int ACT> __device__ __forceinline__ float x3_act(float y) {
    if (ACT == ACT_SIN) return x3_sin(y);
    if (ACT == ACT_SIN30) return x3_sin(__fmul_rn(30.0f, y));
    return act_fwd<ACT>(y);
}

static size_t x3_tile_bytes(int N, int k0, int k1) { return (size_t)(ceil_div(k0, 64) + ceil_div(k1, 64)) * 2 * ((N + 15) & ~15) * 128; }
static size_t x3_layer_bytes(int N, int k0, int k1) { return x3_tile_bytes(N, k0, k1) + (((size_t)((N + 15) & ~15) * 4 + 1023) & ~(size_t)1023); }   // + 2^-e per row

// W (N x [k_a0 | k_a1], row-major, leading dimension ldw) -> per slab: hi tile | lo tile, each n16 rows x 128 B, 128B-swizzled K-major;
// behind the tiles: inv_scale[n16] = 2^-e of every row.  One warp per row: row maximum first, then the slabs.
__global__ void __launch_bounds__(256) x3_pack_kernel(const float* __restrict__ W, int ldw, int N, int n16, int k_a0, int k_a1, unsigned char* __restrict__ dst,
                                                      float* __restrict__ inv_scale) {
    const int s0 = (k_a0 + 63) / 64, n_slabs = s0 + (k_a1 + 63) / 64, lane = threadIdx.x & 31, r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= n16) return;
    const float* row = W + (size_t)r * ldw;
    float mx = 0.f;
    if (r < N) for (int k = lane; k < k_a0 + k_a1; k += 32) mx = fmaxf(mx, fabsf(row[k]));
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    // 2^e with 2^e mx in [2048, 4096) (mx = 0, inf or nan: e = 0)
    int e = 0;
    if (mx > 0.f && mx < 3.0e38f) { int ex; frexpf(mx, &ex); e = 12 - ex; e = e < -100 ? -100 : (e > 100 ? 100 : e); }
    const float scale = ldexpf(1.f, e);
    if (lane == 0) inv_scale[r] = ldexpf(1.f, -e);
    for (int s = 0; s < n_slabs; ++s) {
        const bool second = s >= s0;
        const int k = (second ? s - s0 : s) * 64 + 2 * lane, k_end = second ? k_a1 : k_a0;
        const float* q = row + (second ? k_a0 : 0);
        const float x0 = (r < N && k < k_end) ? q[k] * scale : 0.f, x1 = (r < N && k + 1 < k_end) ? q[k + 1] * scale : 0.f;
        uint32_t hi, lo; x3_split(x0, x1, &hi, &lo);
        unsigned char* t = dst + (size_t)s * 2 * n16 * 128 + (size_t)r * 128 + ((((uint32_t)lane >> 2) ^ ((uint32_t)r & 7u)) << 4) + (((uint32_t)lane & 3u) << 2);
        *reinterpret_cast<uint32_t*>(t) = hi;
        *reinterpret_cast<uint32_t*>(t + (size_t)n16 * 128) = lo;
    }
}

// slab `it` of the CTA's A tile: this thread's 8 rows (warp + 16 j), columns 2 lane, 2 lane + 1 of the slab
constexpr int X3_ROWS = 128 * 32 / X3_THREADS;
__device__ __forceinline__ void x3_load(const LinFwd& p, int m0, int s0, int n_slabs, int it, float2 (&v)[X3_ROWS]) {
    if (it >= n_slabs) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool second = it >= s0;
    const float* base = second ? p.a1.p : p.a0.p;
    const int ld = second ? p.a1.ld : p.a0.ld, div = second ? p.a1.div : p.a0.div, k_end = second ? p.a1.k : p.a0.k;
    const int k = (second ? it - s0 : it) * 64 + 2 * lane;
    const int row0 = m0 + warp;
    const bool in2 = k + 1 < k_end, in1 = k < k_end;
#pragma unroll
    for (int j = 0; j < X3_ROWS; ++j) v[j] = make_float2(0.f, 0.f);
    if (div == 1 && (ld & 1) == 0 && (((uintptr_t)base & 7u) == 0)) {            // the wide sources: 8-byte loads, rows ld apart
        if (in2) {
            const float* q = base + (size_t)row0 * ld + k;
            const size_t step = (size_t)(X3_THREADS / 32) * ld;
#pragma unroll
            for (int j = 0; j < X3_ROWS; ++j)
                if (row0 + (X3_THREADS / 32) * j < p.M) v[j] = *reinterpret_cast<const float2*>(q + j * step);
        } else if (in1) {
#pragma unroll
            for (int j = 0; j < X3_ROWS; ++j)
                if (row0 + (X3_THREADS / 32) * j < p.M) v[j].x = base[(size_t)(row0 + (X3_THREADS / 32) * j) * ld + k];
        }
    } else if (in1) {
#pragma unroll
        for (int j = 0; j < X3_ROWS; ++j) {
            const int r = row0 + (X3_THREADS / 32) * j;
            if (r < p.M) {
                const float* q = base + (size_t)(div == 1 ? r : r / div) * ld + k;
                v[j].x = q[0];
                if (in2) v[j].y = q[1];
            }
        }
    }
}

template <int ACT>
__global__ void __launch_bounds__(X3_THREADS + 64, 1) linear_fwd_x3_kernel(LinFwd p, int nt_tile, const unsigned char* __restrict__ wpk, const float* __restrict__ inv_scale, int n16) {
    extern __shared__ unsigned char x3_raw[];
    unsigned char* base = (unsigned char*)(((uintptr_t)x3_raw + 1023) & ~(uintptr_t)1023);
    // mbarriers: [0,1] stage consumed by its MMAs; [2] accumulator complete; [3,4] weight tiles landed; [5,6] A tiles written (16 warps)
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + 2 * X3_STAGE);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 7);
    const int s0 = (p.a0.k + 63) / 64, s1 = (p.a1.k + 63) / 64, n_slabs = s0 + s1;
    const int m0 = blockIdx.x * 128, n0 = blockIdx.y * nt_tile;
    int nv = p.N - n0; if (nv > nt_tile) nv = nt_tile;                           // valid columns of this CTA
    const int nt = (nv + 15) & ~15;                                              // UMMA N: multiple of 16
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    if (t == 0) {
        for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
        mbar_init(&bars[5], X3_THREADS / 32); mbar_init(&bars[6], X3_THREADS / 32);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_ptr, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr;
    const uint32_t sbase = smem_u32(base);

    if (warp == X3_THREADS / 32) {
        // ---- producer: the slab's weight tiles (hi, lo) as two bulk copies as soon as the stage's previous MMAs are done ----
        if (lane == 0) {
            for (int it = 0; it < n_slabs; ++it) {
                const int st = it & 1;
                if (it >= 2) mbar_wait(&bars[st], (uint32_t)((it >> 1) - 1) & 1u, 33);
                unsigned char* sW_hi = base + st * X3_STAGE + 32768, *sW_lo = sW_hi + X3_NT * 128;
                const unsigned char* src = wpk + (size_t)it * 2 * n16 * 128 + (size_t)n0 * 128;
                mbar_arrive_expect_tx(&bars[3 + st], 2u * (uint32_t)nt * 128u);
                bulk_g2s(sW_hi, src, (uint32_t)nt * 128u, &bars[3 + st]);
                bulk_g2s(sW_lo, src + (size_t)n16 * 128, (uint32_t)nt * 128u, &bars[3 + st]);
            }
        }
    } else if (warp == X3_THREADS / 32 + 1) {
        // ---- issuer: 12 MMAs per slab once its A tiles are written and its weight tiles have landed ----
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_f16((uint32_t)nt);
            for (int it = 0; it < n_slabs; ++it) {
                const int st = it & 1;
                const uint32_t stage = sbase + (uint32_t)st * X3_STAGE, ph = (uint32_t)(it >> 1) & 1u;
                mbar_wait(&bars[5 + st], ph, 35);
                mbar_wait(&bars[3 + st], ph, 34);
                tc_fence_after();
                const bool second = it >= s0;
                const int k_left = (second ? p.a1.k : p.a0.k) - (second ? it - s0 : it) * 64;
                int ks = (k_left + 15) / 16; if (ks > 4) ks = 4;                           // K-steps of this slab that hold data
                const uint32_t a_hi = stage, a_lo = stage + 16384u, w_hi = stage + 32768u, w_lo = w_hi + (uint32_t)X3_NT * 128u;
                for (int k = 0; k < ks; ++k) {
                    umma_f16_ss(tmem, umma_desc_k_sw128(a_hi + k * 32), umma_desc_k_sw128(w_hi + k * 32), idesc, (it | k) != 0);
                    umma_f16_ss(tmem, umma_desc_k_sw128(a_lo + k * 32), umma_desc_k_sw128(w_hi + k * 32), idesc, 1u);
                    umma_f16_ss(tmem, umma_desc_k_sw128(a_hi + k * 32), umma_desc_k_sw128(w_lo + k * 32), idesc, 1u);
                }
                umma_commit(&bars[st]);
                if (it == n_slabs - 1) umma_commit(&bars[2]);
            }
        }
    } else {
        // this thread's rows of a slab are warp + 16 j: (row & 7) == (warp & 7), so the swizzle term of its tile offsets is a constant
        const uint32_t lane_off = (uint32_t)warp * 128u + ((((uint32_t)lane >> 2) ^ ((uint32_t)warp & 7u)) << 4) + (((uint32_t)lane & 3u) << 2);
        auto process = [&](int it, float2 (&v)[X3_ROWS]) {
            if (it >= n_slabs) return;
            const int st = it & 1;
            if (it >= 2) mbar_wait(&bars[st], (uint32_t)((it >> 1) - 1) & 1u, 31);   // the MMAs that read this stage two slabs ago are done
            const uint32_t hi_t = sbase + (uint32_t)st * X3_STAGE + lane_off;
#pragma unroll
            for (int j = 0; j < X3_ROWS; ++j) {
                uint32_t hi, lo; x3_split(v[j].x, v[j].y, &hi, &lo);
                sts_b32(hi_t + (uint32_t)j * (X3_THREADS / 32 * 128u), hi);
                sts_b32(hi_t + 16384u + (uint32_t)j * (X3_THREADS / 32 * 128u), lo);
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars[5 + st]);
        };
        {
            float2 v0[X3_ROWS], v1[X3_ROWS];                     // loads run two slabs (64 KB per SM) ahead of the split
            x3_load(p, m0, s0, n_slabs, 0, v0);
            x3_load(p, m0, s0, n_slabs, 1, v1);
            for (int it = 0; it < n_slabs; it += 2) {
                process(it, v0);     x3_load(p, m0, s0, n_slabs, it + 2, v0);
                process(it + 1, v1); x3_load(p, m0, s0, n_slabs, it + 3, v1);
            }
        }
        mbar_wait(&bars[2], 0u, 32);
        tc_fence_after();
        // epilogue: warp w reads TMEM lanes 32 (w % 4) ..; the warps of a lane quadrant take the 32-column blocks in turn.  Each
        // block goes through a 32 x 33 shared-memory scratch (the ring is free now) so that one store instruction = one row segment.
        const uint32_t scratch = sbase + (uint32_t)warp * (32 * 33 * 4);
        const int quad = warp & 3;
        int rows = p.M - (m0 + quad * 32); if (rows > 32) rows = 32;
        for (int c0 = (warp >> 2) * 32; c0 < nt; c0 += (X3_THREADS / 128) * 32) {
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) sts_b32(scratch + (uint32_t)(lane * 33 + i) * 4u, __float_as_uint(v[i]));
            __syncwarp();
            if (c0 + lane < nv) {
                const int n = n0 + c0 + lane;
                const float bias = p.b[n], inv = inv_scale[n];
                float* out = p.out + (size_t)(m0 + quad * 32) * p.ldo + n;
                float* pre = p.pre ? p.pre + (size_t)(m0 + quad * 32) * p.N + n : nullptr;
                const size_t ldo = (size_t)p.ldo, ldp = (size_t)p.N;
                const uint32_t sc = scratch + (uint32_t)lane * 4u;
#pragma unroll 8
                for (int r = 0; r < 32; ++r) {
                    if (r < rows) {
                        float a; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(a) : "r"(sc + (uint32_t)r * 132u));
                        const float y = fmaf(a, inv, bias);
                        if (pre) pre[r * ldp] = y;
                        out[r * ldo] = x3_act<ACT>(y);
                    }
                }
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

// wpk: this layer's packed weights (x3_layer_bytes); pack_now: (re)build them first (first chunk of a pass)
template <int ACT>
static int launch_x3(const LinFwd& p, cudaStream_t st, unsigned char* wpk, bool pack_now) {
    const size_t smem = 2 * (size_t)X3_STAGE + 1024 + 64;
    SNB_CUDA(cudaFuncSetAttribute(linear_fwd_x3_kernel<ACT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));     // (per device: not cached)
    const int n16 = (p.N + 15) & ~15;
    float* inv_scale = reinterpret_cast<float*>(wpk + x3_tile_bytes(p.N, p.a0.k, p.a1.k));
    if (pack_now) {
        x3_pack_kernel<<<ceil_div(n16, 8), 256, 0, st>>>(p.W, p.ldw, p.N, n16, p.a0.k, p.a1.k, wpk, inv_scale);
        SNB_CHECK_LAUNCH();
    }
    // 256-column tiles unless that leaves most SMs without a CTA (the h/2-wide head layers of one chunk)
    const int nt_tile = (p.N > 128 && ceil_div(p.M, 128) * ceil_div(p.N, X3_NT) < 100) ? 128 : X3_NT;
    dim3 g(ceil_div(p.M, 128), ceil_div(p.N, nt_tile));
    linear_fwd_x3_kernel<ACT><<<g, X3_THREADS + 64, smem, st>>>(p, nt_tile, wpk, inv_scale, n16);
    SNB_CHECK_LAUNCH();
    return 0;
}

template <int ACT>
static int launch_narrow(const LinFwd& p, cudaStream_t st) {
    int blocks = ceil_div(p.M, 32); if (blocks > 148 * 4) blocks = 148 * 4; if (blocks < 1) blocks = 1;
    linear_fwd_narrow_kernel<ACT><<<blocks, 256, 0, st>>>(p);
    SNB_CHECK_LAUNCH();
    return 0;
}

// ------------------------------------------------------------------------------------------------
// launch helpers
// ------------------------------------------------------------------------------------------------
// x3: 0 = FFMA; 1 / 2 = SNB_FP16X3_TC with the packed weights at *wpk already built / to be built now (*wpk advances layer by layer)
static int run_fwd(int act, const LinFwd& p, cudaStream_t st, int x3 = 0, unsigned char** wpk = nullptr) {
    if (p.M == 0) return 0;
    if (p.N <= NARROW_MAX_N && p.a0.k + p.a1.k <= NARROW_MAX_K && p.a0.k + p.a1.k >= 64) {
        switch (act) {
            case ACT_NONE: return launch_narrow<ACT_NONE>(p, st);
            case ACT_SIGMOID: return launch_narrow<ACT_SIGMOID>(p, st);
            case ACT_SOFTPLUS: return launch_narrow<ACT_SOFTPLUS>(p, st);
            case ACT_SIGMOID_PAD: return launch_narrow<ACT_SIGMOID_PAD>(p, st);
            default: break;
        }
    }
    if (x3 && wpk && *wpk && p.N >= 32 && p.a0.k + p.a1.k >= 16 && (act == ACT_NONE || act == ACT_SIN || act == ACT_SIN30 || act == ACT_RELU)) {
        unsigned char* w = *wpk;                              // (short contractions -- K = 3 inputs -- stay on the FFMA kernel)
        *wpk += x3_layer_bytes(p.N, p.a0.k, p.a1.k);
        switch (act) {
            case ACT_NONE: return launch_x3<ACT_NONE>(p, st, w, x3 == 2);
            case ACT_SIN: return launch_x3<ACT_SIN>(p, st, w, x3 == 2);
            case ACT_SIN30: return launch_x3<ACT_SIN30>(p, st, w, x3 == 2);
            default: return launch_x3<ACT_RELU>(p, st, w, x3 == 2);
        }
    }
    dim3 g(ceil_div(p.M, BM), ceil_div(p.N, BN));
    switch (act) {
        case ACT_NONE: linear_fwd_kernel<ACT_NONE><<<g, NT, 0, st>>>(p); break;
        case ACT_SIN: linear_fwd_kernel<ACT_SIN><<<g, NT, 0, st>>>(p); break;
        case ACT_SIN30: linear_fwd_kernel<ACT_SIN30><<<g, NT, 0, st>>>(p); break;
        case ACT_RELU: linear_fwd_kernel<ACT_RELU><<<g, NT, 0, st>>>(p); break;
        case ACT_SIGMOID: linear_fwd_kernel<ACT_SIGMOID><<<g, NT, 0, st>>>(p); break;
        case ACT_SOFTPLUS: linear_fwd_kernel<ACT_SOFTPLUS><<<g, NT, 0, st>>>(p); break;
        case ACT_SIGMOID_PAD: linear_fwd_kernel<ACT_SIGMOID_PAD><<<g, NT, 0, st>>>(p); break;
        default: SNB_FAIL(-3, "bad activation %d", act);
    }
    SNB_CHECK_LAUNCH();
    return 0;
}

static int run_bwd_in(int act, const LinBwdIn& p, cudaStream_t st) {
    if (p.M == 0) return 0;
    dim3 g(ceil_div(p.M, BM), ceil_div(p.K, BN));
    switch (act) {
        case ACT_NONE: linear_bwd_in_kernel<ACT_NONE><<<g, NT, 0, st>>>(p); break;
        case ACT_SIN: linear_bwd_in_kernel<ACT_SIN><<<g, NT, 0, st>>>(p); break;
        case ACT_SIN30: linear_bwd_in_kernel<ACT_SIN30><<<g, NT, 0, st>>>(p); break;
        case ACT_RELU: linear_bwd_in_kernel<ACT_RELU><<<g, NT, 0, st>>>(p); break;
        default: SNB_FAIL(-3, "bad backward activation %d", act);
    }
    SNB_CHECK_LAUNCH();
    return 0;
}

constexpr int kMaxZ = 16;
static int run_bwd_w(LinBwdW p, float* g_params, const Lin& l, float* partial, cudaStream_t st) {
    if (p.M == 0) return 0;
    const int K = p.a0.k + p.a1.k;
    int Z = ceil_div(p.M, 2048); if (Z > kMaxZ) Z = kMaxZ; if (Z < 1) Z = 1;
    p.m_per_z = ceil_div(ceil_div(p.M, Z), BK) * BK;
    Z = ceil_div(p.M, p.m_per_z);
    p.partial = partial; p.N = l.n_out;
    if (K != l.n_in) SNB_FAIL(-3, "internal: weight-gradient K mismatch (%d vs %d)", K, l.n_in);
    dim3 g(ceil_div(K, BN), ceil_div(p.N, BM), Z);
    linear_bwd_w_kernel<<<g, NT, 0, st>>>(p);
    SNB_CHECK_LAUNCH();
    int64_t n = (int64_t)l.n_out * l.n_in + l.n_out;      // weight block and bias are adjacent in the flat layout
    reduce_partials_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g_params + l.w, partial, Z, n);
    SNB_CHECK_LAUNCH();
    return 0;
}

static inline Src S_(const float* p, int ld, int k, int div = 1) { Src s; s.p = p; s.ld = ld; s.k = k; s.div = div; return s; }
static inline Src none() { return S_(nullptr, 0, 0, 1); }

// ------------------------------------------------------------------------------------------------
// chunk buffers
// ------------------------------------------------------------------------------------------------
size_t FieldChunk::plan(Arena& ar, const FieldLayout& L, int Pc, int Rc, bool keep) {
    const int h = L.width, h2 = h / 2, nl = L.n_layers;
    size_t before = ar.off;
    {   // SNB_FP16X3_TC: packed hi / lo weight tiles of every wide layer (bound: one extra slab per layer for the narrow source)
        size_t wb = 0;
        auto add = [&](const Lin& l) { if (l.n_out >= 32 && l.n_in >= 16) wb += x3_layer_bytes(l.n_out, l.n_in, 64); };
        for (int i = 0; i < nl; ++i) add(L.trunk[i]);
        add(L.feats); add(L.rgb0);
        if (L.variant != SNB_NERF) { for (int j = 0; j < 3; ++j) add(L.sun[j]); add(L.sky0); }
        if (L.variant == SNB_SATNERF) add(L.beta0);
        x3_w = ar.take<unsigned char>(wb);
    }
    enc = L.in_xyz != 3 ? ar.take<float>((size_t)Pc * L.in_xyz) : nullptr;
    enc_dir = L.in_dir ? ar.take<float>((size_t)Rc * L.in_dir) : nullptr;
    float* ping[2] = {nullptr, nullptr};
    if (!keep) { ping[0] = ar.take<float>((size_t)Pc * h); ping[1] = ar.take<float>((size_t)Pc * h); }
    for (int i = 0; i < nl; ++i) {
        pre[i] = keep ? ar.take<float>((size_t)Pc * h) : nullptr;
        act[i] = keep ? ar.take<float>((size_t)Pc * h) : ping[i & 1];
    }
    feat = keep ? ar.take<float>((size_t)Pc * h) : ping[nl & 1];
    auto head = [&](float*& pr, float*& ac) { pr = keep ? ar.take<float>((size_t)Pc * h2) : nullptr; ac = ar.take<float>((size_t)Pc * h2); };
    head(rgb1_pre, rgb1);
    if (L.variant != SNB_NERF) {
        if (keep) { for (int j = 0; j < 3; ++j) head(sun_pre[j], sun_act[j]); }
        else { head(sun_pre[0], sun_act[0]); sun_act[1] = rgb1; sun_pre[1] = nullptr; sun_act[2] = sun_act[0]; sun_pre[2] = nullptr; }
        head(sky1_pre, sky1);
    }
    if (L.variant == SNB_SATNERF) { if (keep) head(beta1_pre, beta1); else { beta1 = rgb1; beta1_pre = nullptr; } }
    if (keep) {
        d_feat = ar.take<float>((size_t)Pc * h); d_a = ar.take<float>((size_t)Pc * h); d_b = ar.take<float>((size_t)Pc * h);
        d_t = L.t_dims ? ar.take<float>((size_t)Pc * L.t_dims) : nullptr;
        int kmax = h + (L.in_xyz > L.in_dir ? L.in_xyz : L.in_dir) + 8 + L.t_dims;
        partial = ar.take<float>((size_t)kMaxZ * ((size_t)h * kmax + h));
    }
    return ar.off - before;
}

// ------------------------------------------------------------------------------------------------
// forward of one chunk of Pc points.  raw: (Pc, C) chunk slice.
// ------------------------------------------------------------------------------------------------
int field_forward_chunk(const FieldLayout& L, const float* P, const FieldChunk& c, const FieldInputs& in,
                        float* raw, bool sigma_only, cudaStream_t st) {
    const int h = L.width, h2 = h / 2, nl = L.n_layers, Pc = in.n_points, C = sigma_only ? 1 : L.n_channels;
    const bool siren = L.variant != SNB_NERF;
    Src x = in.xyz;
    if (L.in_xyz != 3) {
        pe_kernel<<<ceil_div(Pc * L.in_xyz, 256), 256, 0, st>>>(in.xyz.p, in.xyz.ld, c.enc, Pc, L.in_xyz / 6);
        SNB_CHECK_LAUNCH();
        x = S_(c.enc, L.in_xyz, L.in_xyz);
    }
    unsigned char* wpk = c.x3_w;
    auto fwd = [&](const Lin& l, Src a0, Src a1, int act, float* pre, float* out, int ldo) {
        LinFwd p; p.a0 = a0; p.a1 = a1; p.W = P + l.w; p.ldw = l.n_in; p.b = P + l.b; p.pre = pre; p.out = out; p.ldo = ldo; p.M = Pc; p.N = l.n_out;
        if (a0.k + a1.k != l.n_in) { set_error("internal: layer K mismatch (%d+%d vs %d)", a0.k, a1.k, l.n_in); return -3; }
        return run_fwd(act, p, st, in.x3, &wpk);
    };
    for (int i = 0; i < nl; ++i) {
        Src a0 = i == 0 ? x : (i == L.skip ? x : S_(c.act[i - 1], h, h));
        Src a1 = i == L.skip ? S_(c.act[i - 1], h, h) : none();
        int act = siren ? (i == 0 ? ACT_SIN30 : ACT_SIN) : ACT_RELU;
        SNB_TRY(fwd(L.trunk[i], a0, a1, act, c.pre[i], c.act[i], h));
    }
    Src top = S_(c.act[nl - 1], h, h);
    SNB_TRY(fwd(L.sigma, top, none(), ACT_SOFTPLUS, nullptr, raw + (sigma_only ? 0 : 3), C));
    if (sigma_only) return 0;
    SNB_TRY(fwd(L.feats, top, none(), ACT_NONE, nullptr, c.feat, h));
    Src f = S_(c.feat, h, h);
    Src dir = none();
    if (L.in_dir) {
        if (L.in_dir != 3) {
            int rows = ceil_div(Pc, in.aux.div);
            pe_kernel<<<ceil_div(rows * L.in_dir, 256), 256, 0, st>>>(in.aux.p, in.aux.ld, c.enc_dir, rows, L.in_dir / 6);
            SNB_CHECK_LAUNCH();
            dir = S_(c.enc_dir, L.in_dir, L.in_dir, in.aux.div);
        } else dir = in.aux;
    }
    SNB_TRY(fwd(L.rgb0, f, dir, siren ? ACT_SIN : ACT_RELU, c.rgb1_pre, c.rgb1, h2));
    SNB_TRY(fwd(L.rgb2, S_(c.rgb1, h2, h2), none(), ACT_SIGMOID_PAD, nullptr, raw + 0, C));
    if (L.variant != SNB_NERF) {
        SNB_TRY(fwd(L.sun[0], f, in.aux, ACT_SIN, c.sun_pre[0], c.sun_act[0], h2));
        SNB_TRY(fwd(L.sun[1], S_(c.sun_act[0], h2, h2), none(), ACT_SIN, c.sun_pre[1], c.sun_act[1], h2));
        SNB_TRY(fwd(L.sun[2], S_(c.sun_act[1], h2, h2), none(), ACT_SIN, c.sun_pre[2], c.sun_act[2], h2));
        SNB_TRY(fwd(L.sun[3], S_(c.sun_act[2], h2, h2), none(), ACT_SIGMOID, nullptr, raw + 4, C));
        SNB_TRY(fwd(L.sky0, in.aux, none(), ACT_RELU, c.sky1_pre, c.sky1, h2));
        SNB_TRY(fwd(L.sky2, S_(c.sky1, h2, h2), none(), ACT_SIGMOID, nullptr, raw + 5, C));
    }
    if (L.variant == SNB_SATNERF) {
        SNB_TRY(fwd(L.beta0, f, in.temb, ACT_SIN, c.beta1_pre, c.beta1, h2));
        SNB_TRY(fwd(L.beta2, S_(c.beta1, h2, h2), none(), ACT_SOFTPLUS, nullptr, raw + 8, C));
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// backward of one chunk; requires field_forward_chunk(keep buffers) on the same chunk just before.
// d_head: (Pc, C) gradient w.r.t. pre-activation head outputs.  g_t_ray: (Rc, tau) or null.
// ------------------------------------------------------------------------------------------------
int field_backward_chunk(const FieldLayout& L, const float* P, float* G, const FieldChunk& c, const FieldInputs& in,
                         const float* d_head, float* g_t_ray, int S, cudaStream_t st) {
    const int h = L.width, h2 = h / 2, nl = L.n_layers, Pc = in.n_points, C = L.n_channels;
    const bool siren = L.variant != SNB_NERF;
    const int hact = siren ? ACT_SIN : ACT_RELU;
    Src x = L.in_xyz != 3 ? S_(c.enc, L.in_xyz, L.in_xyz) : in.xyz;
    Src f = S_(c.feat, h, h);
    auto bw = [&](const Lin& l, const float* dY, int ldy, Src a0, Src a1) {
        LinBwdW p; p.dY = dY; p.ldy = ldy; p.a0 = a0; p.a1 = a1; p.M = Pc;
        return run_bwd_w(p, G, l, c.partial, st);
    };
    auto bi = [&](const Lin& l, const float* dY, int ldy, int k_off, int K, int act, const float* pre, float* dX, int ldx, int accumulate) {
        LinBwdIn p; p.dY = dY; p.ldy = ldy; p.N = l.n_out; p.W = P + l.w; p.ldw = l.n_in; p.k_off = k_off; p.K = K;
        p.pre = pre; p.dX = dX; p.ldx = ldx; p.accumulate = accumulate; p.M = Pc;
        return run_bwd_in(act, p, st);
    };
    bool feat_started = false;
    float* d1 = c.d_a; float* d2 = c.d_b;      // (Pc, <=h) scratch, used with leading dim h2 in the heads
    if (L.variant == SNB_SATNERF) {
        SNB_TRY(bw(L.beta2, d_head + 8, C, S_(c.beta1, h2, h2), none()));
        SNB_TRY(bi(L.beta2, d_head + 8, C, 0, h2, ACT_SIN, c.beta1_pre, d1, h2, 0));
        SNB_TRY(bw(L.beta0, d1, h2, f, in.temb));
        SNB_TRY(bi(L.beta0, d1, h2, 0, h, ACT_NONE, nullptr, c.d_feat, h, 0)); feat_started = true;
        if (g_t_ray) {
            SNB_TRY(bi(L.beta0, d1, h2, h, L.t_dims, ACT_NONE, nullptr, c.d_t, L.t_dims, 0));
            int rc = Pc / S;
            ray_sum_kernel<<<ceil_div(rc * L.t_dims, 128), 128, 0, st>>>(c.d_t, g_t_ray, rc, S, L.t_dims);
            SNB_CHECK_LAUNCH();
        }
    }
    if (L.variant != SNB_NERF) {
        SNB_TRY(bw(L.sun[3], d_head + 4, C, S_(c.sun_act[2], h2, h2), none()));
        SNB_TRY(bi(L.sun[3], d_head + 4, C, 0, h2, ACT_SIN, c.sun_pre[2], d1, h2, 0));
        SNB_TRY(bw(L.sun[2], d1, h2, S_(c.sun_act[1], h2, h2), none()));
        SNB_TRY(bi(L.sun[2], d1, h2, 0, h2, ACT_SIN, c.sun_pre[1], d2, h2, 0));
        SNB_TRY(bw(L.sun[1], d2, h2, S_(c.sun_act[0], h2, h2), none()));
        SNB_TRY(bi(L.sun[1], d2, h2, 0, h2, ACT_SIN, c.sun_pre[0], d1, h2, 0));
        SNB_TRY(bw(L.sun[0], d1, h2, f, in.aux));
        SNB_TRY(bi(L.sun[0], d1, h2, 0, h, ACT_NONE, nullptr, c.d_feat, h, feat_started ? 1 : 0)); feat_started = true;
        SNB_TRY(bw(L.sky2, d_head + 5, C, S_(c.sky1, h2, h2), none()));
        SNB_TRY(bi(L.sky2, d_head + 5, C, 0, h2, ACT_RELU, c.sky1_pre, d1, h2, 0));
        SNB_TRY(bw(L.sky0, d1, h2, in.aux, none()));
    }
    Src dir = none();
    if (L.in_dir) dir = L.in_dir != 3 ? S_(c.enc_dir, L.in_dir, L.in_dir, in.aux.div) : in.aux;
    SNB_TRY(bw(L.rgb2, d_head + 0, C, S_(c.rgb1, h2, h2), none()));
    SNB_TRY(bi(L.rgb2, d_head + 0, C, 0, h2, hact, c.rgb1_pre, d1, h2, 0));
    SNB_TRY(bw(L.rgb0, d1, h2, f, dir));
    SNB_TRY(bi(L.rgb0, d1, h2, 0, h, ACT_NONE, nullptr, c.d_feat, h, feat_started ? 1 : 0));
    // feats + sigma -> top of the trunk
    Src top = S_(c.act[nl - 1], h, h);
    SNB_TRY(bw(L.feats, c.d_feat, h, top, none()));
    SNB_TRY(bw(L.sigma, d_head + 3, C, top, none()));
    int top_act = siren ? (nl - 1 == 0 ? ACT_SIN30 : ACT_SIN) : ACT_RELU;
    SNB_TRY(bi(L.sigma, d_head + 3, C, 0, h, ACT_NONE, nullptr, d1, h, 0));
    SNB_TRY(bi(L.feats, c.d_feat, h, 0, h, top_act, c.pre[nl - 1], d1, h, 1));
    // trunk
    float* cur = d1; float* nxt = d2;
    for (int i = nl - 1; i >= 1; --i) {
        Src a0 = i == L.skip ? x : S_(c.act[i - 1], h, h);
        Src a1 = i == L.skip ? S_(c.act[i - 1], h, h) : none();
        SNB_TRY(bw(L.trunk[i], cur, h, a0, a1));
        int act = siren ? (i - 1 == 0 ? ACT_SIN30 : ACT_SIN) : ACT_RELU;
        SNB_TRY(bi(L.trunk[i], cur, h, i == L.skip ? L.in_xyz : 0, h, act, c.pre[i - 1], nxt, h, 0));
        float* t = cur; cur = nxt; nxt = t;
    }
    SNB_TRY(bw(L.trunk[0], cur, h, x, none()));
    return 0;
}

int launch_points(const float* rays, int ray_cols, int dir_col, const float* z, float* xyz, int r0, int n_rays, int S, cudaStream_t st) {
    if (n_rays == 0) return 0;
    points_kernel<<<ceil_div(n_rays * S, 256), 256, 0, st>>>(rays, ray_cols, dir_col, z, xyz, r0, n_rays, S);
    SNB_CHECK_LAUNCH();
    return 0;
}

}  // namespace snb
