// Alpha compositing of one inference() pass and its gradient.
// Forward restates models/satnerf.py:43-78 (snerf.py:43-74, nerf.py:108-132); backward is the closed
// form autograd derives for those lines.  One thread per ray walks the samples in
// order, so the transmittance product has the same association order as torch.cumprod.
#include "composite.cuh"

namespace snb {

// raw: (R*S, C) field outputs [rgb3, sigma, sun, sky3, beta]; writes every non-null result array.
__global__ void composite_fwd_kernel(CompositeArgs a) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.R) return;
    const int S = a.S, C = a.C;
    const float* z = a.z + (size_t)r * S;
    const float* raw = a.raw + (size_t)r * S * C;
    float T = 1.0f, depth = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
    float xs = 0.f, xa0 = 0.f, xa1 = 0.f, xa2 = 0.f, xb = 0.f, xk0 = 0.f, xk1 = 0.f, xk2 = 0.f;
    for (int i = 0; i < S; ++i) {
        const float* o = raw + (size_t)i * C;
        float sg = o[3];
        if (a.t_min > 0.f && T < a.t_min) {          // early termination (same 32-sample granularity as the fused kernel)
            if ((i & 31) == 0) T = 0.f;
            if (T == 0.f) {
                size_t p = (size_t)r * S + i;
                if (a.weights) a.weights[p] = 0.f;
                if (a.transparency) a.transparency[p] = 0.f;
                if (a.sigma) a.sigma[p] = sg;
                continue;
            }
        }
        float nz = a.noise ? a.noise[(size_t)r * S + i] * a.noise_std : 0.f;                 // :57-58
        float delta = i < S - 1 ? __fsub_rn(z[i + 1], z[i]) : 1e10f;                          // :52-54
        float alpha = 1.0f - expf(-delta * fmaxf(sg + nz, 0.f));                              // :59
        float w = alpha * T;                                                                  // :63
        size_t p = (size_t)r * S + i;
        if (a.weights) a.weights[p] = w;
        if (a.transparency) a.transparency[p] = T;
        if (a.sigma) a.sigma[p] = sg;
        depth += w * z[i];                                                                    // :67
        float r0 = o[0], r1 = o[1], r2 = o[2];
        if (C >= 8) {
            float s = o[4], k0 = o[5], k1 = o[6], k2 = o[7];
            c0 += w * r0 * (s + (1.f - s) * k0);                                              // :68-69
            c1 += w * r1 * (s + (1.f - s) * k1);
            c2 += w * r2 * (s + (1.f - s) * k2);
            if (a.albedo) { a.albedo[p * 3] = r0; a.albedo[p * 3 + 1] = r1; a.albedo[p * 3 + 2] = r2; }
            if (a.sun) a.sun[p] = s;
            if (a.sky) { a.sky[p * 3] = k0; a.sky[p * 3 + 1] = k1; a.sky[p * 3 + 2] = k2; }
            if (C == 9 && a.beta) a.beta[p] = o[8];
            if (a.aux_sums) {
                xs = fmaf(w, s, xs); xa0 = fmaf(w, r0, xa0); xa1 = fmaf(w, r1, xa1); xa2 = fmaf(w, r2, xa2);
                if (C == 9 && !a.no_beta) xb = fmaf(w, o[8], xb);
                xk0 = fmaf(w, k0, xk0); xk1 = fmaf(w, k1, xk1); xk2 = fmaf(w, k2, xk2);
            }
        } else {
            c0 += w * r0; c1 += w * r1; c2 += w * r2;                                         // nerf.py:128
            if (a.nerf_rgb) { a.nerf_rgb[p * 3] = r0; a.nerf_rgb[p * 3 + 1] = r1; a.nerf_rgb[p * 3 + 2] = r2; }
        }
        T = T * ((1.0f - alpha) + 1e-10f);                                                    // :60-62
    }
    if (a.depth) a.depth[r] = depth;
    if (a.aux_sums && C >= 8) {
        float* x = a.aux_sums + (size_t)r * 8;
        x[0] = xs; x[1] = xa0; x[2] = xa1; x[3] = xa2; x[4] = xb; x[5] = xk0; x[6] = xk1; x[7] = xk2;
    }
    if (a.rgb) {
        if (C >= 8) { c0 = fminf(fmaxf(c0, 0.f), 1.f); c1 = fminf(fmaxf(c1, 0.f), 1.f); c2 = fminf(fmaxf(c2, 0.f), 1.f); }   // :70
        a.rgb[r * 3] = c0; a.rgb[r * 3 + 1] = c1; a.rgb[r * 3 + 2] = c2;
    }
}


// ---- fused loss seed (metrics.py:8-92) -------------------------------------------------------------------------
// Upstream gradients of one ray from the loss descriptor.  crgb = CLAMPED colour, depth, bsum = sum_i w_i beta_i.
// Per-sample: g_w_i = dbeta * beta_i, g_beta_i = dbeta * w_i (COLOR_BETA); g_sun_i = sun_a * (T_i - s_i) + sun_b * w_i (SOLAR).
struct RaySeed { float g0, g1, g2, gd, dbeta, sun_a, sun_b; };
__device__ __forceinline__ RaySeed loss_seed(const CompositeBwdArgs& a, int r, float c0, float c1, float c2, float depth, float bsum) {
    RaySeed q; q.g0 = q.g1 = q.g2 = q.gd = q.dbeta = q.sun_a = q.sun_b = 0.f;
    const float gt0 = a.g_terms ? a.g_terms[0] : 1.f, gt1 = a.g_terms ? a.g_terms[1] : 1.f;
    const float gt2 = a.g_terms ? a.g_terms[2] : 1.f, gt3 = a.g_terms ? a.g_terms[3] : 1.f;
    if (a.loss_kind == SNB_LOSS_COLOR_MSE) {                          // mean over 3N elements of (rgb - t)^2
        const float k = gt0 * 2.0f * a.inv_n / 3.0f;
        q.g0 = k * (c0 - a.target[r * 3]); q.g1 = k * (c1 - a.target[r * 3 + 1]); q.g2 = k * (c2 - a.target[r * 3 + 2]);
    } else if (a.loss_kind == SNB_LOSS_COLOR_BETA) {                  // (rgb - t)^2 / (2 beta^2), (3 + mean log beta) / 2
        const float beta = bsum + a.beta_min, ib2 = 1.0f / (beta * beta);
        const float d0 = c0 - a.target[r * 3], d1 = c1 - a.target[r * 3 + 1], d2 = c2 - a.target[r * 3 + 2];
        const float k = gt0 * a.inv_n / 3.0f;
        q.g0 = k * d0 * ib2; q.g1 = k * d1 * ib2; q.g2 = k * d2 * ib2;
        q.dbeta = -k * (d0 * d0 + d1 * d1 + d2 * d2) * ib2 / beta + gt1 * 0.5f * a.inv_n / beta;
    } else if (a.loss_kind == SNB_LOSS_DEPTH) {                       // lambda/3 * mean(weight * (depth - t)^2)
        const float wt = a.target_w ? a.target_w[r] : 1.f;
        q.gd = gt0 * (a.lambda / 3.0f) * 2.0f * wt * (depth - a.target[r]) * a.inv_n;
    } else if (a.loss_kind == SNB_LOSS_SOLAR) {                       // lambda/3 * (mean sum (T - s)^2 + mean (1 - sum w s)), T, w detached
        q.sun_a = gt2 * (a.lambda / 3.0f) * a.inv_n * 2.0f;           //   d/ds (T - s)^2 = -2 (T - s) = 2 (s - T)
        q.sun_b = -gt3 * (a.lambda / 3.0f) * a.inv_n;
    }
    return q;
}

// Writes d_head (R*S, C): gradient w.r.t. the PRE-activation outputs of the field heads.
__global__ void composite_bwd_kernel(CompositeBwdArgs a) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.R) return;
    const int S = a.S, C = a.C;
    const size_t base = (size_t)r * S;
    const float* z = a.z + base;
    const bool sat = C >= 8;
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
    if (a.g_rgb) { g0 = a.g_rgb[r * 3]; g1 = a.g_rgb[r * 3 + 1]; g2 = a.g_rgb[r * 3 + 2]; }
    float gd = a.g_depth ? a.g_depth[r] : 0.f;
    const float* col = sat ? a.albedo : a.nerf_rgb;
    RaySeed seed; seed.dbeta = seed.sun_a = seed.sun_b = 0.f;
    const bool fused = a.loss_kind != 0;
    if ((sat && a.g_rgb) || fused) {
        // clamp mask needs the un-clamped colour: re-accumulate it in the forward order (fused loss: also depth, sum w*beta)
        float c0 = 0.f, c1 = 0.f, c2 = 0.f, dep = 0.f, bsum = 0.f;
        for (int i = 0; i < S; ++i) {
            size_t p = base + i; float w = a.weights[p];
            if (sat) {
                float s = a.sun[p];
                c0 += w * col[p * 3] * (s + (1.f - s) * a.sky[p * 3]);
                c1 += w * col[p * 3 + 1] * (s + (1.f - s) * a.sky[p * 3 + 1]);
                c2 += w * col[p * 3 + 2] * (s + (1.f - s) * a.sky[p * 3 + 2]);
            } else { c0 += w * col[p * 3]; c1 += w * col[p * 3 + 1]; c2 += w * col[p * 3 + 2]; }
            dep += w * z[i];
            if (C == 9) bsum += w * a.beta[p];
        }
        if (fused) {
            const float k0 = sat ? fminf(fmaxf(c0, 0.f), 1.f) : c0, k1 = sat ? fminf(fmaxf(c1, 0.f), 1.f) : c1, k2 = sat ? fminf(fmaxf(c2, 0.f), 1.f) : c2;
            seed = loss_seed(a, r, k0, k1, k2, dep, bsum);
            g0 = seed.g0; g1 = seed.g1; g2 = seed.g2; gd = seed.gd;
        }
        if (sat) {
            if (!(c0 >= 0.f && c0 <= 1.f)) g0 = 0.f;
            if (!(c1 >= 0.f && c1 <= 1.f)) g1 = 0.f;
            if (!(c2 >= 0.f && c2 <= 1.f)) g2 = 0.f;
        }
    }
    float suffix = 0.f;   // sum_{k>i} (G_k alpha_k + gT_k) T_k
    for (int i = S - 1; i >= 0; --i) {
        size_t p = base + i;
        float w = a.weights[p], T = a.transparency[p], sg = a.sigma[p];
        float nz = a.noise ? a.noise[p] * a.noise_std : 0.f;
        float delta = i < S - 1 ? __fsub_rn(z[i + 1], z[i]) : 1e10f;
        float act = sg + nz;
        float e = expf(-delta * fmaxf(act, 0.f));
        float alpha = 1.0f - e;
        float q = (1.0f - alpha) + 1e-10f;
        float c_r = col[p * 3], c_g = col[p * 3 + 1], c_b = col[p * 3 + 2];
        float s = 0.f, k0 = 0.f, k1 = 0.f, k2 = 0.f, i0 = 1.f, i1 = 1.f, i2 = 1.f;
        if (sat) {
            s = a.sun[p]; k0 = a.sky[p * 3]; k1 = a.sky[p * 3 + 1]; k2 = a.sky[p * 3 + 2];
            i0 = s + (1.f - s) * k0; i1 = s + (1.f - s) * k1; i2 = s + (1.f - s) * k2;
        }
        float gw = fused ? (C == 9 ? seed.dbeta * a.beta[p] : 0.f) : (a.g_weights ? a.g_weights[p] : 0.f);
        float G = gw + gd * z[i] + g0 * c_r * i0 + g1 * c_g * i1 + g2 * c_b * i2;
        float gT = (!fused && a.g_transparency) ? a.g_transparency[p] : 0.f;
        float d_alpha = G * T - suffix / q;
        suffix += (G * alpha + gT) * T;
        float d_sigma = act > 0.f ? d_alpha * (delta * e) : 0.f;
        float* out = a.d_head + p * C;
        // colour head: alb = sigmoid(y)*1.002 - 0.001                                    (satnerf.py:193-195)
        float dc0 = g0 * w * i0, dc1 = g1 * w * i1, dc2 = g2 * w * i2;
        if (sat && !fused && a.g_albedo) { dc0 += a.g_albedo[p * 3]; dc1 += a.g_albedo[p * 3 + 1]; dc2 += a.g_albedo[p * 3 + 2]; }
        float s0 = (c_r + 0.001f) / 1.002f, s1 = (c_g + 0.001f) / 1.002f, s2 = (c_b + 0.001f) / 1.002f;
        out[0] = dc0 * 1.002f * s0 * (1.f - s0);
        out[1] = dc1 * 1.002f * s1 * (1.f - s1);
        out[2] = dc2 * 1.002f * s2 * (1.f - s2);
        out[3] = d_sigma * (-expm1f(-sg));              // softplus' = sigmoid(y) = 1 - exp(-softplus(y))
        if (sat) {
            float ds = g0 * w * c_r * (1.f - k0) + g1 * w * c_g * (1.f - k1) + g2 * w * c_b * (1.f - k2);
            if (fused) ds += seed.sun_a * (s - T) + seed.sun_b * w;
            else if (a.g_sun) ds += a.g_sun[p];
            out[4] = ds * s * (1.f - s);
            float dk0 = g0 * w * c_r * (1.f - s), dk1 = g1 * w * c_g * (1.f - s), dk2 = g2 * w * c_b * (1.f - s);
            if (!fused && a.g_sky) { dk0 += a.g_sky[p * 3]; dk1 += a.g_sky[p * 3 + 1]; dk2 += a.g_sky[p * 3 + 2]; }
            out[5] = dk0 * k0 * (1.f - k0); out[6] = dk1 * k1 * (1.f - k1); out[7] = dk2 * k2 * (1.f - k2);
            if (C == 9) { float b = a.beta[p]; out[8] = (fused ? seed.dbeta * w : (a.g_beta ? a.g_beta[p] : 0.f)) * (-expm1f(-b)); }
        }
    }
}

// Same gradient, one WARP per ray: lanes hold samples (coalesced loads), the suffix sum of the transmittance
// recurrence is a warp shuffle scan.  Also emits per-ray sums of every d_head channel (ray_sums: (R, 16) floats),
// which the head-bias / sky-colour gradients reduce further.  Used by the tensor-core backward.
__global__ void composite_bwd_warp_kernel(CompositeBwdArgs a, float* __restrict__ ray_sums) {
    const int lane = threadIdx.x & 31;
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= a.R) return;
    const int S = a.S, C = a.C;
    const size_t base = (size_t)r * S;
    const bool sat = C >= 8;
    float g0 = 0.f, g1 = 0.f, g2 = 0.f;
    if (a.g_rgb) { g0 = a.g_rgb[r * 3]; g1 = a.g_rgb[r * 3 + 1]; g2 = a.g_rgb[r * 3 + 2]; }
    float gd = a.g_depth ? a.g_depth[r] : 0.f;
    const float* col = sat ? a.albedo : a.nerf_rgb;
    RaySeed seed; seed.dbeta = seed.sun_a = seed.sun_b = 0.f;
    const bool fused = a.loss_kind != 0;
    if ((sat && a.g_rgb) || fused) {                   // clamp mask from the un-clamped colour (fused loss: also depth, sum w*beta)
        float c0 = 0.f, c1 = 0.f, c2 = 0.f, dep = 0.f, bsum = 0.f;
        for (int i = lane; i < S; i += 32) {
            size_t p = base + i; float w = a.weights[p];
            if (sat) {
                float s = a.sun[p];
                c0 += w * col[p * 3] * (s + (1.f - s) * a.sky[p * 3]);
                c1 += w * col[p * 3 + 1] * (s + (1.f - s) * a.sky[p * 3 + 1]);
                c2 += w * col[p * 3 + 2] * (s + (1.f - s) * a.sky[p * 3 + 2]);
            } else { c0 += w * col[p * 3]; c1 += w * col[p * 3 + 1]; c2 += w * col[p * 3 + 2]; }
            dep += w * a.z[p];
            if (C == 9) bsum += w * a.beta[p];
        }
        for (int off = 16; off; off >>= 1) {
            c0 += __shfl_xor_sync(~0u, c0, off); c1 += __shfl_xor_sync(~0u, c1, off); c2 += __shfl_xor_sync(~0u, c2, off);
            dep += __shfl_xor_sync(~0u, dep, off); bsum += __shfl_xor_sync(~0u, bsum, off);
        }
        if (fused) {
            const float k0 = sat ? fminf(fmaxf(c0, 0.f), 1.f) : c0, k1 = sat ? fminf(fmaxf(c1, 0.f), 1.f) : c1, k2 = sat ? fminf(fmaxf(c2, 0.f), 1.f) : c2;
            seed = loss_seed(a, r, k0, k1, k2, dep, bsum);
            g0 = seed.g0; g1 = seed.g1; g2 = seed.g2; gd = seed.gd;
        }
        if (sat) {
            if (!(c0 >= 0.f && c0 <= 1.f)) g0 = 0.f;
            if (!(c1 >= 0.f && c1 <= 1.f)) g1 = 0.f;
            if (!(c2 >= 0.f && c2 <= 1.f)) g2 = 0.f;
        }
    }
    float carry = 0.f;                                  // sum over samples beyond the current 32-sample window
    float sums[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float amax = 0.f;
    for (int hi = ((S + 31) / 32) * 32; hi > 0; hi -= 32) {
        const int i = hi - 32 + lane;
        const bool ok = i < S;
        size_t p = base + (ok ? i : 0);
        float w = 0.f, T = 0.f, sg = 0.f, delta = 0.f, act = 0.f, e = 1.f, alpha = 0.f, q = 1.f, G = 0.f, gT = 0.f, zi = 0.f;
        float c_r = 0.f, c_g = 0.f, c_b = 0.f, s = 0.f, k0 = 0.f, k1 = 0.f, k2 = 0.f, i0 = 1.f, i1 = 1.f, i2 = 1.f;
        if (ok) {
            w = a.weights[p]; T = a.transparency[p]; sg = a.sigma[p]; zi = a.z[p];
            float nz = a.noise ? a.noise[p] * a.noise_std : 0.f;
            delta = i < S - 1 ? __fsub_rn(a.z[p + 1], zi) : 1e10f;
            act = sg + nz; e = expf(-delta * fmaxf(act, 0.f)); alpha = 1.0f - e; q = (1.0f - alpha) + 1e-10f;
            c_r = col[p * 3]; c_g = col[p * 3 + 1]; c_b = col[p * 3 + 2];
            if (sat) {
                s = a.sun[p]; k0 = a.sky[p * 3]; k1 = a.sky[p * 3 + 1]; k2 = a.sky[p * 3 + 2];
                i0 = s + (1.f - s) * k0; i1 = s + (1.f - s) * k1; i2 = s + (1.f - s) * k2;
            }
            const float gw = fused ? (C == 9 ? seed.dbeta * a.beta[p] : 0.f) : (a.g_weights ? a.g_weights[p] : 0.f);
            G = gw + gd * zi + g0 * c_r * i0 + g1 * c_g * i1 + g2 * c_b * i2;
            gT = (!fused && a.g_transparency) ? a.g_transparency[p] : 0.f;
        }
        // suffix (exclusive) sum of term_k = (G_k alpha_k + gT_k) T_k over k > i
        float term = ok ? (G * alpha + gT) * T : 0.f;
        float incl = term;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) { float o = __shfl_down_sync(~0u, incl, off); if (lane + off < 32) incl += o; }
        const float suffix = incl - term + carry;
        carry += __shfl_sync(~0u, incl, 0);
        if (ok) {
            float d_alpha = G * T - suffix / q;
            float d_sigma = act > 0.f ? d_alpha * (delta * e) : 0.f;
            float* out = a.d_head + p * C;
            float dc0 = g0 * w * i0, dc1 = g1 * w * i1, dc2 = g2 * w * i2;
            if (sat && !fused && a.g_albedo) { dc0 += a.g_albedo[p * 3]; dc1 += a.g_albedo[p * 3 + 1]; dc2 += a.g_albedo[p * 3 + 2]; }
            float s0 = (c_r + 0.001f) / 1.002f, s1 = (c_g + 0.001f) / 1.002f, s2 = (c_b + 0.001f) / 1.002f;
            float o0 = dc0 * 1.002f * s0 * (1.f - s0), o1 = dc1 * 1.002f * s1 * (1.f - s1), o2 = dc2 * 1.002f * s2 * (1.f - s2);
            float o3 = d_sigma * (-expm1f(-sg));
            out[0] = o0; out[1] = o1; out[2] = o2; out[3] = o3;
            sums[0] += o0; sums[1] += o1; sums[2] += o2; sums[3] += o3;
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(o0), fabsf(o1)), fmaxf(fabsf(o2), fabsf(o3))));
            if (sat) {
                float ds = g0 * w * c_r * (1.f - k0) + g1 * w * c_g * (1.f - k1) + g2 * w * c_b * (1.f - k2);
                if (fused) ds += seed.sun_a * (s - T) + seed.sun_b * w;
                else if (a.g_sun) ds += a.g_sun[p];
                float o4 = ds * s * (1.f - s);
                float dk0 = g0 * w * c_r * (1.f - s), dk1 = g1 * w * c_g * (1.f - s), dk2 = g2 * w * c_b * (1.f - s);
                if (!fused && a.g_sky) { dk0 += a.g_sky[p * 3]; dk1 += a.g_sky[p * 3 + 1]; dk2 += a.g_sky[p * 3 + 2]; }
                float o5 = dk0 * k0 * (1.f - k0), o6 = dk1 * k1 * (1.f - k1), o7 = dk2 * k2 * (1.f - k2);
                out[4] = o4; out[5] = o5; out[6] = o6; out[7] = o7;
                sums[4] += o4; sums[5] += o5; sums[6] += o6; sums[7] += o7;
                amax = fmaxf(amax, fmaxf(fmaxf(fabsf(o4), fabsf(o5)), fmaxf(fabsf(o6), fabsf(o7))));
                if (C == 9) { float b = a.beta[p]; float o8 = (fused ? seed.dbeta * w : (a.g_beta ? a.g_beta[p] : 0.f)) * (-expm1f(-b)); out[8] = o8; sums[8] += o8; amax = fmaxf(amax, fabsf(o8)); }
            }
        }
    }
    if (a.absmax) {                                     // (order-independent: exact max; NaN / Inf leave the scale at 1, see loss_scale)
        for (int off = 16; off; off >>= 1) amax = fmaxf(amax, __shfl_xor_sync(~0u, amax, off));
        if (lane == 0) atomicMax(reinterpret_cast<unsigned int*>(a.absmax), __float_as_uint(amax));
    }
    if (ray_sums) {
#pragma unroll
        for (int c = 0; c < 9; ++c) {
            float v = sums[c];
            for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(~0u, v, off);
            if (lane == 0) ray_sums[(size_t)r * 16 + c] = v;
        }
    }
}


// ---- loss terms of one pass (what metrics.py computes from the result dict) ------------------------------------------
// warp per ray -> per-ray contributions (R,4); one block then sums them in a fixed order (double accumulation).
__global__ void loss_per_ray_kernel(LossFwdArgs a) {
    const int lane = threadIdx.x & 31;
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= a.R) return;
    const int S = a.S; const size_t base = (size_t)r * S;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
    if (a.kind == SNB_LOSS_COLOR_MSE || a.kind == SNB_LOSS_COLOR_BETA) {
        float bsum = 0.f;
        if (a.kind == SNB_LOSS_COLOR_BETA) {
            for (int i = lane; i < S; i += 32) bsum += a.weights[base + i] * a.beta[base + i];
            for (int off = 16; off; off >>= 1) bsum += __shfl_xor_sync(~0u, bsum, off);
        }
        if (lane == 0) {
            const float d0 = a.rgb[r * 3] - a.target[r * 3], d1 = a.rgb[r * 3 + 1] - a.target[r * 3 + 1], d2 = a.rgb[r * 3 + 2] - a.target[r * 3 + 2];
            const float sq = d0 * d0 + d1 * d1 + d2 * d2;
            if (a.kind == SNB_LOSS_COLOR_MSE) t0 = sq * a.inv_n / 3.0f;
            else { const float beta = bsum + a.beta_min; t0 = sq / (2.0f * beta * beta) * a.inv_n / 3.0f; t1 = 0.5f * logf(beta) * a.inv_n; }
        }
    } else if (a.kind == SNB_LOSS_DEPTH) {
        if (lane == 0) { const float d = a.depth[r] - a.target[r]; t0 = (a.lambda / 3.0f) * (a.target_w ? a.target_w[r] : 1.f) * d * d * a.inv_n; }
    } else if (a.kind == SNB_LOSS_SOLAR) {
        float e2 = 0.f, ws = 0.f;
        for (int i = lane; i < S; i += 32) { const float s = a.sun[base + i], d = a.transparency[base + i] - s; e2 += d * d; ws += a.weights[base + i] * s; }
        for (int off = 16; off; off >>= 1) { e2 += __shfl_xor_sync(~0u, e2, off); ws += __shfl_xor_sync(~0u, ws, off); }
        t2 = (a.lambda / 3.0f) * e2 * a.inv_n; t3 = (a.lambda / 3.0f) * (1.0f - ws) * a.inv_n;
    }
    if (lane == 0) *reinterpret_cast<float4*>(a.per_ray + (size_t)r * 4) = make_float4(t0, t1, t2, t3);
}

__global__ void loss_reduce_kernel(LossFwdArgs a) {
    __shared__ double sh[4][256];
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (int r = threadIdx.x; r < a.R; r += blockDim.x) {
        const float4 v = *reinterpret_cast<const float4*>(a.per_ray + (size_t)r * 4);
        acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
    }
    for (int k = 0; k < 4; ++k) sh[k][threadIdx.x] = acc[k];
    __syncthreads();
    for (int st = 128; st; st >>= 1) {
        if ((int)threadIdx.x < st) for (int k = 0; k < 4; ++k) sh[k][threadIdx.x] += sh[k][threadIdx.x + st];
        __syncthreads();
    }
    if (threadIdx.x < 4) {
        double v = sh[threadIdx.x][0];
        // the constant of (3 + mean log beta) / 2 counts once per GLOBAL batch: R * inv_n is this call's share of it
        if (threadIdx.x == 1 && a.kind == SNB_LOSS_COLOR_BETA) v += 1.5 * (double)a.R * (double)a.inv_n;
        a.terms[threadIdx.x] = (float)v;
    }
}

// dL/d(result-dict tensors) of the same terms (the autograd backward of the loss classes of metrics.py): warp per ray.
__global__ void loss_grad_kernel(LossBwdArgs b) {
    const LossFwdArgs& a = b.f;
    const int lane = threadIdx.x & 31;
    const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (r >= a.R) return;
    const int S = a.S; const size_t base = (size_t)r * S;
    CompositeBwdArgs q{};
    q.loss_kind = a.kind; q.lambda = a.lambda; q.beta_min = a.beta_min; q.inv_n = a.inv_n; q.target = a.target; q.target_w = a.target_w; q.g_terms = b.g_terms;
    float bsum = 0.f;
    if (a.kind == SNB_LOSS_COLOR_BETA) {
        for (int i = lane; i < S; i += 32) bsum += a.weights[base + i] * a.beta[base + i];
        for (int off = 16; off; off >>= 1) bsum += __shfl_xor_sync(~0u, bsum, off);
    }
    const bool col = a.kind == SNB_LOSS_COLOR_MSE || a.kind == SNB_LOSS_COLOR_BETA;
    const RaySeed sd = loss_seed(q, r, col ? a.rgb[r * 3] : 0.f, col ? a.rgb[r * 3 + 1] : 0.f, col ? a.rgb[r * 3 + 2] : 0.f,
                                 a.kind == SNB_LOSS_DEPTH ? a.depth[r] : 0.f, bsum);
    if (lane == 0) {
        if (b.g_rgb) { b.g_rgb[r * 3] = sd.g0; b.g_rgb[r * 3 + 1] = sd.g1; b.g_rgb[r * 3 + 2] = sd.g2; }
        if (b.g_depth) b.g_depth[r] = sd.gd;
    }
    for (int i = lane; i < S; i += 32) {
        const size_t p = base + i;
        if (b.g_weights) b.g_weights[p] = a.kind == SNB_LOSS_COLOR_BETA ? sd.dbeta * a.beta[p] : 0.f;
        if (b.g_beta) b.g_beta[p] = a.kind == SNB_LOSS_COLOR_BETA ? sd.dbeta * a.weights[p] : 0.f;
        if (b.g_sun) b.g_sun[p] = a.kind == SNB_LOSS_SOLAR ? sd.sun_a * (a.sun[p] - a.transparency[p]) + sd.sun_b * a.weights[p] : 0.f;
    }
}

int launch_loss_backward(const LossBwdArgs& a, cudaStream_t st) {
    if (a.f.R == 0) return 0;
    loss_grad_kernel<<<ceil_div(a.f.R, 4), 128, 0, st>>>(a);
    SNB_CHECK_LAUNCH();
    return 0;
}

int launch_loss_forward(const LossFwdArgs& a, cudaStream_t st) {
    if (a.R == 0) { SNB_CUDA(cudaMemsetAsync(a.terms, 0, 16, st)); return 0; }
    loss_per_ray_kernel<<<ceil_div(a.R, 4), 128, 0, st>>>(a);
    SNB_CHECK_LAUNCH();
    loss_reduce_kernel<<<1, 256, 0, st>>>(a);
    SNB_CHECK_LAUNCH();
    return 0;
}

int launch_composite_bwd_warp(const CompositeBwdArgs& a, float* ray_sums, cudaStream_t st) {
    if (a.R == 0) return 0;
    composite_bwd_warp_kernel<<<ceil_div(a.R, 4), 128, 0, st>>>(a, ray_sums);
    SNB_CHECK_LAUNCH();
    return 0;
}

int launch_composite_fwd(const CompositeArgs& a, cudaStream_t st) {
    if (a.R == 0) return 0;
    composite_fwd_kernel<<<ceil_div(a.R, 128), 128, 0, st>>>(a);
    SNB_CHECK_LAUNCH();
    return 0;
}

int launch_composite_bwd(const CompositeBwdArgs& a, cudaStream_t st) {
    if (a.R == 0) return 0;
    composite_bwd_kernel<<<ceil_div(a.R, 128), 128, 0, st>>>(a);
    SNB_CHECK_LAUNCH();
    return 0;
}

}  // namespace snb
