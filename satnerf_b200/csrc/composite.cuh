// Argument blocks of the compositing kernels (composite.cu).
#pragma once
#include "common.cuh"

namespace snb {

struct CompositeArgs {
    int R, S, C;                 // rays, samples, channels of `raw` (9 sat-nerf, 8 s-nerf, 4 nerf)
    const float* raw;            // (R*S, C) post-activation field outputs [rgb3, sigma, sun, sky3, beta]
    const float* z;              // (R,S)
    const float* noise;          // (R,S) or null
    float noise_std;
    float *rgb, *depth, *weights, *transparency, *albedo, *sun, *sky, *beta, *sigma, *nerf_rgb;
    float* aux_sums;             // (R,8) [sum w*sun, sum w*albedo(3), sum w*beta, sum w*sky(3)] or null (sat-nerf / s-nerf)
    float t_min;                 // > 0: stop once the transmittance is below t_min (remaining weights 0)
    int no_beta;                 // aux_sums[4] = 0 (SNB_PASS_NO_BETA)
};

struct CompositeBwdArgs {
    int R, S, C;
    const float *z, *noise; float noise_std;
    const float *weights, *transparency, *sigma, *albedo, *sun, *sky, *beta, *nerf_rgb;          // saved forward results
    const float *g_rgb, *g_depth, *g_weights, *g_transparency, *g_albedo, *g_sun, *g_sky, *g_beta;  // upstream (nullable)
    float* d_head;               // (R*S, C) gradient w.r.t. pre-activation head outputs
    float* absmax;               // optional device scalar (zeroed by the caller): max |d_head| (bit pattern, atomicMax) -- the loss scale of the fp16 backward
    // fused loss seed (snb_loss_desc): loss_kind != 0 replaces the upstream g_* by dL/d(outputs) computed per ray
    int loss_kind; float lambda, beta_min, inv_n;       // inv_n = 1 / n_rays_mean
    const float *target, *target_w, *g_terms;
};

struct LossFwdArgs {
    int R, S, C, kind; float lambda, beta_min, inv_n;
    const float *rgb, *depth, *weights, *transparency, *beta, *sun, *target, *target_w;
    float* per_ray;              // (R,4) workspace
    float* terms;                // (4)
};
int launch_loss_forward(const LossFwdArgs& a, cudaStream_t st);
struct LossBwdArgs {
    LossFwdArgs f; const float* g_terms;
    float *g_rgb, *g_depth, *g_weights, *g_beta, *g_sun;      // gradients w.r.t. the result-dict tensors (nullable)
};
int launch_loss_backward(const LossBwdArgs& a, cudaStream_t st);

inline int fill_loss(CompositeBwdArgs& b, const snb_loss_desc* l) {
    b.loss_kind = 0; b.lambda = 0.f; b.beta_min = 0.f; b.inv_n = 0.f; b.target = b.target_w = b.g_terms = nullptr;
    if (!l) return 0;
    if (l->kind < SNB_LOSS_COLOR_MSE || l->kind > SNB_LOSS_SOLAR) SNB_FAIL(-1, "unknown loss kind %d", l->kind);
    if (l->n_rays_mean < 1) SNB_FAIL(-1, "snb_loss_desc.n_rays_mean must be the (global) number of rays the reference's mean() divides by");
    if (l->kind != SNB_LOSS_SOLAR && !l->target) SNB_FAIL(-1, "snb_loss_desc.target is required for this loss");
    if (l->kind == SNB_LOSS_COLOR_BETA && b.C != 9) SNB_FAIL(-1, "SNB_LOSS_COLOR_BETA needs the sat-nerf beta head");
    if (l->kind == SNB_LOSS_SOLAR && b.C < 8) SNB_FAIL(-1, "SNB_LOSS_SOLAR needs a sun-visibility head (s-nerf / sat-nerf)");
    b.loss_kind = l->kind; b.lambda = l->lambda; b.beta_min = l->beta_min; b.inv_n = 1.0f / (float)l->n_rays_mean;
    b.target = l->target; b.target_w = l->target_weight; b.g_terms = l->g_terms;
    return 0;
}

int launch_composite_fwd(const CompositeArgs& a, cudaStream_t st);
int launch_composite_bwd(const CompositeBwdArgs& a, cudaStream_t st);
// warp-per-ray variant; ray_sums (R,16): per-ray sums of the d_head channels (nullable)
int launch_composite_bwd_warp(const CompositeBwdArgs& a, float* ray_sums, cudaStream_t st);

}  // namespace snb
