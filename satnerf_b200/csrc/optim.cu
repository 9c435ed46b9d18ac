// Optimiser step on the flat parameter buffer (the reference trains with torch.optim.Adam(lr, weight_decay=0), main.py:81-94
// via train_utils.py:24-53).  One elementwise pass over params / grads / exp_avg / exp_avg_sq: HBM-bound, 28 bytes per parameter.
#include "common.cuh"
#include <cstring>
#include <cmath>

namespace snb {

struct AdamArgs { float lr, b1, b2, omb1, omb2, eps, wd, bc1, bc2_sqrt; };      // omb = 1 - beta, rounded from double like torch's Python scalars

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n, AdamArgs a) {
    const long long n4 = n >> 2;
    const float step_size = a.lr / a.bc1;
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        if (a.wd != 0.f) gg = fmaf(a.wd, pp, gg);                       // L2 penalty folded into the gradient (torch.optim.Adam)
        mm = fmaf(a.b1, mm, a.omb1 * gg);                          // lerp(m, g, 1 - b1)
        vv = fmaf(a.b2, vv, a.omb2 * (gg * gg));
        const float denom = sqrtf(vv) / a.bc2_sqrt + a.eps;
        pp -= step_size * (mm / denom);
    };
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        upd(pp.x, gg.x, mm.x, vv.x); upd(pp.y, gg.y, mm.y, vv.y); upd(pp.z, gg.z, mm.z, vv.z); upd(pp.w, gg.w, mm.w, vv.w);
        reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        upd(p[i], g[i], m[i], v[i]);
}

// Data-parallel step fused with its collective (ZeRO-1 style, one kernel per flat buffer): every rank owns a contiguous shard of the
// parameters; for its shard it SUMS the gradient shards of all ranks straight from their memory (NVLink peer loads through the
// symmetric-memory mapping, fixed rank order: deterministic, and every element is reduced by exactly one rank, so replicas stay
// bit-identical), applies Adam with its own moments, and writes the updated parameters into every rank's parameter buffer (peer
// stores).  Replaces all-reduce (10.5 MB through NCCL: ~50 us on 2 GPUs) + Adam; the caller brackets it with two device barriers
// (gradients final on every rank before; parameters delivered after).
constexpr int kMaxRanks = 8;
struct ShardedAdamArgs {
    float* params[kMaxRanks]; const float* grads[kMaxRanks];
    float *m, *v; long long lo4, hi4; int world, rank; AdamArgs a;
};
__global__ void adam_sharded_kernel(const __grid_constant__ ShardedAdamArgs A) {
    const float step_size = A.a.lr / A.a.bc1;
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        if (A.a.wd != 0.f) gg = fmaf(A.a.wd, pp, gg);
        mm = fmaf(A.a.b1, mm, A.a.omb1 * gg);
        vv = fmaf(A.a.b2, vv, A.a.omb2 * (gg * gg));
        pp -= step_size * (mm / (sqrtf(vv) / A.a.bc2_sqrt + A.a.eps));
    };
    for (long long i = A.lo4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < A.hi4; i += (long long)gridDim.x * blockDim.x) {
        float4 g[kMaxRanks];
#pragma unroll
        for (int r = 0; r < kMaxRanks; ++r) if (r < A.world) g[r] = reinterpret_cast<const float4*>(A.grads[r])[i];      // all peer loads in flight
        float4 gs = g[0];
#pragma unroll
        for (int r = 1; r < kMaxRanks; ++r) if (r < A.world) { gs.x += g[r].x; gs.y += g[r].y; gs.z += g[r].z; gs.w += g[r].w; }
        float4 pp = reinterpret_cast<const float4*>(A.params[A.rank])[i], mm = reinterpret_cast<float4*>(A.m)[i], vv = reinterpret_cast<float4*>(A.v)[i];
        upd(pp.x, gs.x, mm.x, vv.x); upd(pp.y, gs.y, mm.y, vv.y); upd(pp.z, gs.z, mm.z, vv.z); upd(pp.w, gs.w, mm.w, vv.w);
        reinterpret_cast<float4*>(A.m)[i] = mm; reinterpret_cast<float4*>(A.v)[i] = vv;
#pragma unroll
        for (int r = 0; r < kMaxRanks; ++r) if (r < A.world) reinterpret_cast<float4*>(A.params[r])[i] = pp;
    }
}


// The same step for up to kMaxShardBufs flat buffers in ONE launch (a field, a second field, the embedding table), with the NVSwitch
// doing the reduction and the broadcast when the buffers have multicast mappings: `multimem.ld_reduce` returns the sum over all
// ranks of one 16-byte gradient piece (reduced inside the switch: one NVLink round trip instead of world - 1 peer loads) and
// `multimem.st` delivers the updated parameters to every rank with one store.  Without multicast addresses: peer loads / stores as above.
constexpr int kMaxShardBufs = 4;
struct ShardBuf {
    float* params[kMaxRanks]; const float* grads[kMaxRanks];
    float* mc_params; const float* mc_grads;
    float *m, *v; long long lo4, hi4; float bc1, bc2_sqrt;
};
struct ShardedMultiArgs { ShardBuf b[kMaxShardBufs]; int n_bufs, world, rank; AdamArgs a; };

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
    float4 r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(mc) : "memory");
    return r;
}
__device__ __forceinline__ void multimem_st(float* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

__global__ void adam_sharded_multi_kernel(const __grid_constant__ ShardedMultiArgs A) {
    long long total = 0;
    for (int b = 0; b < A.n_bufs; ++b) total += A.b[b].hi4 - A.b[b].lo4;
    for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < total; j += (long long)gridDim.x * blockDim.x) {
        int bi = 0; long long i = j;
        while (i >= A.b[bi].hi4 - A.b[bi].lo4) { i -= A.b[bi].hi4 - A.b[bi].lo4; ++bi; }
        const ShardBuf& B = A.b[bi];
        i += B.lo4;
        float4 gs;
        if (B.mc_grads) gs = multimem_ld_reduce_add(B.mc_grads + 4 * i);
        else {
            float4 g[kMaxRanks];
#pragma unroll
            for (int r = 0; r < kMaxRanks; ++r) if (r < A.world) g[r] = reinterpret_cast<const float4*>(B.grads[r])[i];      // all peer loads in flight
            gs = g[0];
#pragma unroll
            for (int r = 1; r < kMaxRanks; ++r) if (r < A.world) { gs.x += g[r].x; gs.y += g[r].y; gs.z += g[r].z; gs.w += g[r].w; }
        }
        const float step_size = A.a.lr / B.bc1;
        auto upd = [&](float& pp, float gg, float& mm, float& vv) {
            if (A.a.wd != 0.f) gg = fmaf(A.a.wd, pp, gg);
            mm = fmaf(A.a.b1, mm, A.a.omb1 * gg);
            vv = fmaf(A.a.b2, vv, A.a.omb2 * (gg * gg));
            pp -= step_size * (mm / (sqrtf(vv) / B.bc2_sqrt + A.a.eps));
        };
        float4 pp = reinterpret_cast<const float4*>(B.params[A.rank])[i], mm = reinterpret_cast<float4*>(B.m)[i], vv = reinterpret_cast<float4*>(B.v)[i];
        upd(pp.x, gs.x, mm.x, vv.x); upd(pp.y, gs.y, mm.y, vv.y); upd(pp.z, gs.z, mm.z, vv.z); upd(pp.w, gs.w, mm.w, vv.w);
        reinterpret_cast<float4*>(B.m)[i] = mm; reinterpret_cast<float4*>(B.v)[i] = vv;
        if (B.mc_params) multimem_st(B.mc_params + 4 * i, pp);
        else {
#pragma unroll
            for (int r = 0; r < kMaxRanks; ++r) if (r < A.world) reinterpret_cast<float4*>(B.params[r])[i] = pp;
        }
    }
}

}  // namespace snb

using namespace snb;

extern "C" SNB_API int snb_adam_step_sharded(float* const* peer_params, const float* const* peer_grads, int world, int rank,
                                             float* exp_avg, float* exp_avg_sq, long long n,
                                             double lr, double beta1, double beta2, double eps, double weight_decay, int step, void* stream) {
    if (!peer_params || !peer_grads || !exp_avg || !exp_avg_sq || n < 0 || step < 1) SNB_FAIL(-1, "snb_adam_step_sharded: bad argument");
    if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world) SNB_FAIL(-1, "snb_adam_step_sharded: world %d / rank %d unsupported (<= %d ranks)", world, rank, kMaxRanks);
    if (n % 4) SNB_FAIL(-1, "snb_adam_step_sharded: the buffers must hold a multiple of 4 floats (pad them)");
    ShardedAdamArgs A; memset(&A, 0, sizeof(A));
    for (int r = 0; r < world; ++r) {
        if (!peer_params[r] || !peer_grads[r] || (((uintptr_t)peer_params[r] | (uintptr_t)peer_grads[r]) & 15)) SNB_FAIL(-1, "snb_adam_step_sharded: null / unaligned peer buffer %d", r);
        A.params[r] = peer_params[r]; A.grads[r] = peer_grads[r];
    }
    if (((uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) SNB_FAIL(-1, "snb_adam_step_sharded: moments must be 16-byte aligned");
    const long long n4 = n / 4, base = n4 / world, rem = n4 % world;            // contiguous shards; the first `rem` ranks take one extra float4
    A.lo4 = rank * base + (rank < rem ? rank : rem); A.hi4 = A.lo4 + base + (rank < rem ? 1 : 0);
    A.m = exp_avg; A.v = exp_avg_sq; A.world = world; A.rank = rank;
    A.a.lr = (float)lr; A.a.b1 = (float)beta1; A.a.b2 = (float)beta2; A.a.eps = (float)eps; A.a.wd = (float)weight_decay;
    A.a.omb1 = (float)(1.0 - beta1); A.a.omb2 = (float)(1.0 - beta2);
    A.a.bc1 = (float)(1.0 - pow(beta1, (double)step)); A.a.bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)step));
    if (A.hi4 == A.lo4) return 0;
    long long cnt = A.hi4 - A.lo4; int blocks = (int)((cnt + 255) / 256); if (blocks > 148 * 4) blocks = 148 * 4;
    adam_sharded_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(A);
    SNB_CHECK_LAUNCH();
    return 0;
}


extern "C" SNB_API int snb_adam_step_sharded_multi(const snb_sharded_buffer* bufs, int n_buffers, int world, int rank,
                                                   double lr, double beta1, double beta2, double eps, double weight_decay, void* stream) {
    if (!bufs || n_buffers < 1 || n_buffers > kMaxShardBufs) SNB_FAIL(-1, "snb_adam_step_sharded_multi: 1..%d buffers", kMaxShardBufs);
    if (world < 1 || world > kMaxRanks || rank < 0 || rank >= world) SNB_FAIL(-1, "snb_adam_step_sharded_multi: world %d / rank %d unsupported (<= %d ranks)", world, rank, kMaxRanks);
    ShardedMultiArgs A; memset(&A, 0, sizeof(A));
    A.n_bufs = n_buffers; A.world = world; A.rank = rank;
    A.a.lr = (float)lr; A.a.b1 = (float)beta1; A.a.b2 = (float)beta2; A.a.eps = (float)eps; A.a.wd = (float)weight_decay;
    A.a.omb1 = (float)(1.0 - beta1); A.a.omb2 = (float)(1.0 - beta2);
    long long total = 0;
    for (int b = 0; b < n_buffers; ++b) {
        const snb_sharded_buffer& u = bufs[b]; ShardBuf& B = A.b[b];
        if (!u.peer_params || !u.peer_grads || !u.exp_avg || !u.exp_avg_sq || u.n < 0 || u.step < 1) SNB_FAIL(-1, "snb_adam_step_sharded_multi: bad buffer %d", b);
        if (u.n % 4) SNB_FAIL(-1, "snb_adam_step_sharded_multi: buffer %d must hold a multiple of 4 floats (pad it)", b);
        for (int r = 0; r < world; ++r) {
            if (!u.peer_params[r] || !u.peer_grads[r] || (((uintptr_t)u.peer_params[r] | (uintptr_t)u.peer_grads[r]) & 15)) SNB_FAIL(-1, "snb_adam_step_sharded_multi: null / unaligned peer buffer %d of buffer %d", r, b);
            B.params[r] = u.peer_params[r]; B.grads[r] = u.peer_grads[r];
        }
        if (((uintptr_t)u.exp_avg | (uintptr_t)u.exp_avg_sq | (uintptr_t)u.mc_params | (uintptr_t)u.mc_grads) & 15) SNB_FAIL(-1, "snb_adam_step_sharded_multi: moments / multicast addresses must be 16-byte aligned");
        B.mc_params = u.mc_params; B.mc_grads = u.mc_grads; B.m = u.exp_avg; B.v = u.exp_avg_sq;
        const long long n4 = u.n / 4, base = n4 / world, rem = n4 % world;      // contiguous shards; the first `rem` ranks take one extra float4
        B.lo4 = rank * base + (rank < rem ? rank : rem); B.hi4 = B.lo4 + base + (rank < rem ? 1 : 0);
        B.bc1 = (float)(1.0 - pow(beta1, (double)u.step)); B.bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)u.step));
        total += B.hi4 - B.lo4;
    }
    if (total == 0) return 0;
    int blocks = (int)((total + 255) / 256); if (blocks > 148 * 4) blocks = 148 * 4;
    adam_sharded_multi_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(A);
    SNB_CHECK_LAUNCH();
    return 0;
}


extern "C" SNB_API int snb_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                                     double lr, double beta1, double beta2, double eps, double weight_decay, int step, void* stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || n < 0 || step < 1) SNB_FAIL(-1, "snb_adam_step: bad argument");
    if (((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15) SNB_FAIL(-1, "snb_adam_step: buffers must be 16-byte aligned");
    if (n == 0) return 0;
    // (hyper-parameters are doubles as torch's Python scalars are: 1 - 0.999f differs from float(1 - 0.999) by 5e-5 relative)
    const double beta1_d = beta1, beta2_d = beta2;
    AdamArgs a; a.lr = (float)lr; a.b1 = (float)beta1; a.b2 = (float)beta2; a.eps = (float)eps; a.wd = (float)weight_decay; a.omb1 = (float)(1.0 - beta1_d); a.omb2 = (float)(1.0 - beta2_d);
    a.bc1 = (float)(1.0 - pow(beta1_d, (double)step)); a.bc2_sqrt = (float)sqrt(1.0 - pow(beta2_d, (double)step));
    long long n4 = n >> 2; int blocks = (int)((n4 + 255) / 256); if (blocks > 148 * 8) blocks = 148 * 8; if (blocks < 1) blocks = 1;
    adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(params, grads, exp_avg, exp_avg_sq, n, a);
    SNB_CHECK_LAUNCH();
    return 0;
}
