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
};

struct CompositeBwdArgs {
    int R, S, C;
    const float *z, *noise; float noise_std;
    const float *weights, *transparency, *sigma, *albedo, *sun, *sky, *beta, *nerf_rgb;          // saved forward results
    const float *g_rgb, *g_depth, *g_weights, *g_transparency, *g_albedo, *g_sun, *g_sky, *g_beta;  // upstream (nullable)
    float* d_head;               // (R*S, C) gradient w.r.t. pre-activation head outputs
};

int launch_composite_fwd(const CompositeArgs& a, cudaStream_t st);
int launch_composite_bwd(const CompositeBwdArgs& a, cudaStream_t st);
// warp-per-ray variant; ray_sums (R,16): per-ray sums of the d_head channels (nullable)
int launch_composite_bwd_warp(const CompositeBwdArgs& a, float* ray_sums, cudaStream_t st);

}  // namespace snb
