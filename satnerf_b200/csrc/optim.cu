// Optimiser step on the flat parameter buffer (the reference trains with torch.optim.Adam(lr, weight_decay=0), main.py:81-94
// via train_utils.py:24-53).  One elementwise pass over params / grads / exp_avg / exp_avg_sq: HBM-bound, 28 bytes per parameter.
#include "common.cuh"

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

}  // namespace snb

using namespace snb;

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
