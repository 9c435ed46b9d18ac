// extern "C" SNB_API entry points of the render pass (declared in include/satnerf_b200.h) and the
// orchestration of the fp32 SIMT path: rays are processed in chunks of whole rays so that the
// per-layer activations of a chunk fit the caller's workspace (the reference bounds memory the same
// way with args.chunk, models/satnerf.py:30-40).
#include "common.cuh"
#include "composite.cuh"
#include "simt_field.cuh"
#include "tc_field.cuh"

using namespace snb;

namespace {

constexpr int kChunkPoints = 148 * 64;      // 74 row blocks of 128: one CTA per SM in the hi+lo layer kernel (2 column blocks per layer)

struct PassPlan {
    int rays_per_chunk;
    float *raw, *d_head, *xyz, *raw_chunk;
    FieldChunk chunk;
};

int check_pass(const FieldLayout& L, const snb_pass_desc* p) {
    if (!p) SNB_FAIL(-1, "null pass descriptor");
    if (p->n_rays < 0 || p->n_samples < 1) SNB_FAIL(-1, "bad pass shape R=%d S=%d", p->n_rays, p->n_samples);
    int need = L.variant == SNB_NERF ? 8 : 11;
    if (p->ray_cols != 0 && p->ray_cols < need) SNB_FAIL(-1, "rays need %d columns for this variant, got %d", need, p->ray_cols);
    if (p->march_along_sun && L.variant == SNB_NERF) SNB_FAIL(-1, "solar-correction pass is undefined for nerf");
    if (p->precision != SNB_FP32_SIMT && p->precision != SNB_FP16_TC && p->precision != SNB_FP16X3_TC) SNB_FAIL(-1, "unknown precision %d", p->precision);
    if ((int64_t)p->n_rays * p->n_samples > (int64_t)1 << 30) SNB_FAIL(-1, "too many points in one pass");
    return 0;
}

void plan_pass(Arena& ar, const FieldLayout& L, const snb_pass_desc* p, bool backward, PassPlan* pl) {
    const int S = p->n_samples, C = L.n_channels;
    int rc = kChunkPoints / S; if (rc < 1) rc = 1; if (rc > p->n_rays) rc = p->n_rays > 0 ? p->n_rays : 1;
    pl->rays_per_chunk = rc;
    const size_t P = (size_t)p->n_rays * S, Pc = (size_t)rc * S;
    pl->raw = backward ? nullptr : ar.take<float>(P * C);
    pl->d_head = backward ? ar.take<float>(P * C) : nullptr;
    pl->raw_chunk = backward ? ar.take<float>(Pc * C) : nullptr;
    pl->xyz = ar.take<float>(Pc * 3);
    pl->chunk.plan(ar, L, (int)Pc, rc, backward);
}

FieldInputs chunk_inputs(const FieldLayout& L, const snb_pass_desc* p, const snb_render_io* io, const float* xyz, int r0, int rc) {
    const int S = p->n_samples;
    FieldInputs in;
    in.n_points = rc * S;
    in.xyz = io->xyz ? Src{io->xyz + (size_t)r0 * S * 3, 3, 3, 1} : Src{xyz, 3, 3, 1};
    // sat-nerf / s-nerf: sun direction rays[:, 8:11] (rendering.py:87/:99); nerf: view direction rays[:, 3:6] (:112)
    int aux_col = L.variant == SNB_NERF ? 3 : 8;
    if (io->aux_dir) in.aux = Src{io->aux_dir + (size_t)r0 * 3, 3, 3, S};
    else in.aux = Src{io->rays + (size_t)r0 * p->ray_cols + aux_col, p->ray_cols, 3, S};
    in.temb = L.t_dims ? Src{io->t_emb + (size_t)r0 * L.t_dims, L.t_dims, L.t_dims, S} : Src{nullptr, 0, 0, 1};
    return in;
}

}  // namespace

extern "C" SNB_API int snb_render_workspace(const snb_field_desc* f, const snb_pass_desc* p, int backward, size_t* bytes) {
    FieldLayout L; SNB_TRY(build_layout(f, &L)); SNB_TRY(check_pass(L, p));
    if (!bytes) SNB_FAIL(-1, "null output pointer");
    size_t simt = 0, tc = 0;
    { Arena ar(nullptr, 0); PassPlan pl; plan_pass(ar, L, p, backward != 0, &pl); simt = ar.off; }
    if (p->precision == SNB_FP16_TC) SNB_TRY(tc_workspace(L, p, backward != 0, &tc));
    *bytes = (simt > tc ? simt : tc) + 256;
    return 0;
}

extern "C" SNB_API int snb_render_stash_bytes(const snb_field_desc* f, const snb_pass_desc* p, size_t* bytes) {
    FieldLayout L; SNB_TRY(build_layout(f, &L)); SNB_TRY(check_pass(L, p));
    if (!bytes) SNB_FAIL(-1, "null output pointer");
    *bytes = 0;
    if (p->precision != SNB_FP16_TC) return 0;
    return tc_stash_bytes(L, p, bytes);
}

extern "C" SNB_API int snb_render_forward(const snb_field_desc* f, const snb_pass_desc* p, const snb_render_io* io,
                                  void* workspace, size_t workspace_bytes, void* stream) {
    FieldLayout L; SNB_TRY(build_layout(f, &L)); SNB_TRY(check_pass(L, p));
    if (p->n_rays == 0) return 0;               // empty batch (pointers of empty tensors may be null)
    if (!io || !io->params || !io->z_vals) SNB_FAIL(-1, "snb_render_forward: params and z_vals are required");
    if (!io->rays && !(io->xyz && io->aux_dir)) SNB_FAIL(-1, "snb_render_forward: rays (or xyz + aux_dir) are required");
    if (L.t_dims && !io->t_emb) SNB_FAIL(-1, "snb_render_forward: sat-nerf needs t_emb (the reference raises on ts=None, satnerf.py:204)");
    if (p->n_rays == 0) return 0;
    if (!workspace) SNB_FAIL(-1, "snb_render_forward: null workspace");
    cudaStream_t st = (cudaStream_t)stream;
    if (p->precision == SNB_FP16_TC) {
        int r = tc_render_forward(L, p, io, workspace, workspace_bytes, st);
        if (r != 1) return r;      // 1 = configuration not covered by the tensor-core kernel: fp32 CUDA-core path below
    }

    if (p->flags & SNB_PASS_SIGMA_ONLY) SNB_FAIL(-1, "SNB_PASS_SIGMA_ONLY is provided by the fused tensor-core path only (precision SNB_FP16_TC, a shape it covers)");
    Arena ar(workspace, workspace_bytes); PassPlan pl; plan_pass(ar, L, p, false, &pl);
    if (ar.overflow) SNB_FAIL(-4, "snb_render_forward: workspace too small (%zu bytes given)", workspace_bytes);
    const int S = p->n_samples, C = L.n_channels;
    const int dir_col = p->march_along_sun ? 8 : 3;
    for (int r0 = 0; r0 < p->n_rays; r0 += pl.rays_per_chunk) {
        int rc = p->n_rays - r0 < pl.rays_per_chunk ? p->n_rays - r0 : pl.rays_per_chunk;
        if (!io->xyz) SNB_TRY(launch_points(io->rays, p->ray_cols, dir_col, io->z_vals, pl.xyz, r0, rc, S, st));
        FieldInputs in = chunk_inputs(L, p, io, pl.xyz, r0, rc);
        in.x3 = p->precision == SNB_FP16X3_TC ? (r0 == 0 ? 2 : 1) : 0;
        SNB_TRY(field_forward_chunk(L, io->params, pl.chunk, in, pl.raw + (size_t)r0 * S * C, false, st));
    }
    CompositeArgs a{};
    a.R = p->n_rays; a.S = S; a.C = C; a.raw = pl.raw; a.z = io->z_vals; a.noise = io->noise; a.noise_std = p->noise_std;
    a.rgb = io->rgb; a.depth = io->depth; a.weights = io->weights; a.transparency = io->transparency;
    a.albedo = io->albedo; a.sun = io->sun; a.sky = io->sky; a.beta = io->beta; a.sigma = io->sigma; a.nerf_rgb = io->nerf_rgb;
    a.aux_sums = io->aux_sums; a.t_min = p->t_min; a.no_beta = (p->flags & SNB_PASS_NO_BETA) ? 1 : 0;
    return launch_composite_fwd(a, st);
}

extern "C" SNB_API int snb_render_backward(const snb_field_desc* f, const snb_pass_desc* p, const snb_render_io* io,
                                   const snb_render_grads* g, void* workspace, size_t workspace_bytes, void* stream) {
    FieldLayout L; SNB_TRY(build_layout(f, &L)); SNB_TRY(check_pass(L, p));
    if (p->n_rays == 0) return 0;
    if (!io || !g || !io->params || !io->z_vals || !g->g_params) SNB_FAIL(-1, "snb_render_backward: missing required pointer");
    if (!io->rays && !(io->xyz && io->aux_dir)) SNB_FAIL(-1, "snb_render_backward: rays (or xyz + aux_dir) are required");
    if (!io->weights || !io->transparency || !io->sigma) SNB_FAIL(-1, "snb_render_backward: forward stash (weights, transparency, sigma) missing");
    if (L.variant != SNB_NERF && (!io->albedo || !io->sun || !io->sky)) SNB_FAIL(-1, "snb_render_backward: saved albedo/sun/sky missing");
    if (L.variant == SNB_SATNERF && (!io->beta || !io->t_emb)) SNB_FAIL(-1, "snb_render_backward: saved beta / t_emb missing");
    if (L.variant == SNB_NERF && !io->nerf_rgb) SNB_FAIL(-1, "snb_render_backward: saved per-sample rgb missing");
    if (p->n_rays == 0) return 0;
    if (!workspace) SNB_FAIL(-1, "snb_render_backward: null workspace");
    cudaStream_t st = (cudaStream_t)stream;
    if (p->precision == SNB_FP16_TC) {
        int r = tc_render_backward(L, p, io, g, workspace, workspace_bytes, st);
        if (r != 1) return r;      // 1 = "not available for this shape": use the fp32 chain below
    }
    Arena ar(workspace, workspace_bytes); PassPlan pl; plan_pass(ar, L, p, true, &pl);
    if (ar.overflow) SNB_FAIL(-4, "snb_render_backward: workspace too small (%zu bytes given)", workspace_bytes);
    const int S = p->n_samples, C = L.n_channels;
    CompositeBwdArgs b{};
    b.R = p->n_rays; b.S = S; b.C = C; b.z = io->z_vals; b.noise = io->noise; b.noise_std = p->noise_std;
    b.weights = io->weights; b.transparency = io->transparency; b.sigma = io->sigma; b.albedo = io->albedo; b.sun = io->sun;
    b.sky = io->sky; b.beta = io->beta; b.nerf_rgb = io->nerf_rgb;
    b.g_rgb = g->g_rgb; b.g_depth = g->g_depth; b.g_weights = g->g_weights; b.g_transparency = g->g_transparency;
    b.g_albedo = g->g_albedo; b.g_sun = g->g_sun; b.g_sky = g->g_sky; b.g_beta = g->g_beta; b.d_head = pl.d_head;
    SNB_TRY(fill_loss(b, g->loss));
    SNB_TRY(launch_composite_bwd(b, st));
    const int dir_col = p->march_along_sun ? 8 : 3;
    for (int r0 = 0; r0 < p->n_rays; r0 += pl.rays_per_chunk) {
        int rc = p->n_rays - r0 < pl.rays_per_chunk ? p->n_rays - r0 : pl.rays_per_chunk;
        if (!io->xyz) SNB_TRY(launch_points(io->rays, p->ray_cols, dir_col, io->z_vals, pl.xyz, r0, rc, S, st));
        FieldInputs in = chunk_inputs(L, p, io, pl.xyz, r0, rc);
        in.x3 = p->precision == SNB_FP16X3_TC ? (r0 == 0 ? 2 : 1) : 0;       // the recompute runs the arithmetic the forward ran
        SNB_TRY(field_forward_chunk(L, io->params, pl.chunk, in, pl.raw_chunk, false, st));
        float* gt = (g->g_t_emb && L.t_dims) ? g->g_t_emb + (size_t)r0 * L.t_dims : nullptr;
        SNB_TRY(field_backward_chunk(L, io->params, g->g_params, pl.chunk, in, pl.d_head + (size_t)r0 * S * C, gt, S, st));
    }
    return 0;
}

static int loss_args(const snb_pass_desc* p, const snb_render_io* io, const snb_loss_desc* loss, LossFwdArgs* a, const char* who) {
    if (!p || !io || !loss) SNB_FAIL(-1, "%s: null argument", who);
    if (p->n_rays < 0 || p->n_samples < 1) SNB_FAIL(-1, "bad pass shape R=%d S=%d", p->n_rays, p->n_samples);
    if (loss->kind < SNB_LOSS_COLOR_MSE || loss->kind > SNB_LOSS_SOLAR) SNB_FAIL(-1, "unknown loss kind %d", loss->kind);
    if (loss->n_rays_mean < 1) SNB_FAIL(-1, "snb_loss_desc.n_rays_mean must be >= 1");
    a->R = p->n_rays; a->S = p->n_samples; a->kind = loss->kind; a->lambda = loss->lambda; a->beta_min = loss->beta_min; a->inv_n = 1.0f / (float)loss->n_rays_mean;
    a->rgb = io->rgb; a->depth = io->depth; a->weights = io->weights; a->transparency = io->transparency; a->beta = io->beta; a->sun = io->sun;
    a->target = loss->target; a->target_w = loss->target_weight;
    if (p->n_rays == 0) return 0;
    switch (loss->kind) {
        case SNB_LOSS_COLOR_BETA: if (!io->weights || !io->beta) SNB_FAIL(-1, "%s: weights and beta are required", who); /* fallthrough */
        case SNB_LOSS_COLOR_MSE: if (!io->rgb || !loss->target) SNB_FAIL(-1, "%s: rgb and target are required", who); break;
        case SNB_LOSS_DEPTH: if (!io->depth || !loss->target) SNB_FAIL(-1, "%s: depth and target are required", who); break;
        default: if (!io->sun || !io->weights || !io->transparency) SNB_FAIL(-1, "%s: sun, weights and transparency of the solar-correction pass are required", who);
    }
    return 0;
}

extern "C" SNB_API int snb_loss_backward(const snb_pass_desc* p, const snb_render_io* io, const snb_loss_desc* loss,
                                         float* g_rgb, float* g_depth, float* g_weights, float* g_beta, float* g_sun, void* stream) {
    LossBwdArgs b{};
    SNB_TRY(loss_args(p, io, loss, &b.f, "snb_loss_backward"));
    b.g_terms = loss->g_terms; b.g_rgb = g_rgb; b.g_depth = g_depth; b.g_weights = g_weights; b.g_beta = g_beta; b.g_sun = g_sun;
    return launch_loss_backward(b, (cudaStream_t)stream);
}

extern "C" SNB_API int snb_loss_forward(const snb_pass_desc* p, const snb_render_io* io, const snb_loss_desc* loss, float* terms,
                                        void* workspace, size_t workspace_bytes, void* stream) {
    if (!terms) SNB_FAIL(-1, "snb_loss_forward: null output");
    LossFwdArgs a{};
    SNB_TRY(loss_args(p, io, loss, &a, "snb_loss_forward"));
    if ((size_t)p->n_rays * 16 > workspace_bytes || (!workspace && p->n_rays)) SNB_FAIL(-4, "snb_loss_forward: workspace too small (%zu bytes given)", workspace_bytes);
    a.per_ray = (float*)workspace; a.terms = terms;
    return launch_loss_forward(a, (cudaStream_t)stream);
}

extern "C" SNB_API int snb_field_workspace(const snb_field_desc* f, int n_points, size_t* bytes) {
    FieldLayout L; SNB_TRY(build_layout(f, &L));
    if (!bytes || n_points < 0) SNB_FAIL(-1, "snb_field_workspace: bad arguments");
    int pc = n_points < kChunkPoints ? (n_points > 0 ? n_points : 1) : kChunkPoints;
    Arena ar(nullptr, 0); FieldChunk c; c.plan(ar, L, pc, pc, false);
    *bytes = ar.off + 256;
    return 0;
}

extern "C" SNB_API int snb_field_forward(const snb_field_desc* f, const float* params, const float* xyz, const float* aux_dir,
                                 const float* t_emb, float* out, int n_points, int sigma_only, int precision,
                                 void* workspace, size_t workspace_bytes, void* stream) {
    FieldLayout L; SNB_TRY(build_layout(f, &L));
    if (!params || !xyz || !out || n_points < 0) SNB_FAIL(-1, "snb_field_forward: bad arguments");
    if (!sigma_only && L.variant != SNB_NERF && !aux_dir) SNB_FAIL(-1, "snb_field_forward: input_sun_dir is required");
    if (!sigma_only && L.variant == SNB_NERF && !aux_dir) SNB_FAIL(-1, "snb_field_forward: input_dir is required");
    if (!sigma_only && L.t_dims && !t_emb) SNB_FAIL(-1, "snb_field_forward: input_t is required (satnerf.py:204)");
    if (precision != SNB_FP32_SIMT && precision != SNB_FP16X3_TC)
        SNB_FAIL(-1, "snb_field_forward: the per-point API runs at SNB_FP32_SIMT or SNB_FP16X3_TC (the fused SNB_FP16_TC kernel takes per-ray inputs)");
    if (n_points == 0) return 0;
    if (!workspace) SNB_FAIL(-1, "snb_field_forward: null workspace");
    int pc = n_points < kChunkPoints ? n_points : kChunkPoints;
    Arena ar(workspace, workspace_bytes); FieldChunk c; c.plan(ar, L, pc, pc, false);
    if (ar.overflow) SNB_FAIL(-4, "snb_field_forward: workspace too small");
    const int C = sigma_only ? 1 : L.n_channels;
    for (int p0 = 0; p0 < n_points; p0 += pc) {
        int n = n_points - p0 < pc ? n_points - p0 : pc;
        FieldInputs in; in.n_points = n;
        in.xyz = Src{xyz + (size_t)p0 * 3, 3, 3, 1};
        in.aux = aux_dir ? Src{aux_dir + (size_t)p0 * 3, 3, 3, 1} : Src{nullptr, 0, 0, 1};
        in.temb = (L.t_dims && t_emb) ? Src{t_emb + (size_t)p0 * L.t_dims, L.t_dims, L.t_dims, 1} : Src{nullptr, 0, 0, 1};
        in.x3 = precision == SNB_FP16X3_TC ? (p0 == 0 ? 2 : 1) : 0;
        SNB_TRY(field_forward_chunk(L, params, c, in, out + (size_t)p0 * C, sigma_only != 0, (cudaStream_t)stream));
    }
    return 0;
}

// d(loss)/d(pre-activation head outputs) from d(loss)/d(outputs) and the outputs themselves (satnerf.py:183-206): softplus' =
// 1 - exp(-out), sigmoid' = s (1 - s), the colour padding contributes 1.002 (rgb = s * 1.002 - 0.001)
__global__ void field_head_grad_kernel(const float* __restrict__ d_out, const float* __restrict__ out, float* __restrict__ d_head, long long n_points,
                                       int C, int sigma_only) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_points * C; i += (long long)gridDim.x * blockDim.x) {
        const long long p = i / C; const int c = (int)(i - p * C);
        float g = 0.f;
        if (sigma_only) { if (c == 3) { const float o = out[p]; g = d_out[p] * (1.f - expf(-o)); } }
        else {
            const float o = out[i], d = d_out[i];
            if (c < 3) { const float sgm = (o + 0.001f) / 1.002f; g = d * 1.002f * sgm * (1.f - sgm); }
            else if (c == 3 || c == 8) g = d * (1.f - expf(-o));
            else g = d * o * (1.f - o);
        }
        d_head[i] = g;
    }
}

extern "C" SNB_API int snb_field_backward_workspace(const snb_field_desc* f, int n_points, size_t* bytes) {
    FieldLayout L; SNB_TRY(build_layout(f, &L));
    if (!bytes || n_points < 0) SNB_FAIL(-1, "snb_field_backward_workspace: bad arguments");
    int pc = n_points < kChunkPoints ? (n_points > 0 ? n_points : 1) : kChunkPoints;
    Arena ar(nullptr, 0); FieldChunk c; c.plan(ar, L, pc, pc, true);
    ar.take<float>((size_t)pc * L.n_channels * 2);
    *bytes = ar.off + 512;
    return 0;
}

extern "C" SNB_API int snb_field_backward(const snb_field_desc* f, const float* params, const float* xyz, const float* aux_dir, const float* t_emb,
                                          const float* out, const float* d_out, float* g_params, float* g_t_emb, int n_points, int sigma_only,
                                          void* workspace, size_t workspace_bytes, void* stream) {
    FieldLayout L; SNB_TRY(build_layout(f, &L));
    if (!params || !xyz || !out || !d_out || !g_params || n_points < 0) SNB_FAIL(-1, "snb_field_backward: bad arguments");
    if (!aux_dir && L.variant != SNB_NERF) SNB_FAIL(-1, "snb_field_backward: input_sun_dir is required");
    if (!aux_dir && L.variant == SNB_NERF && L.in_dir) SNB_FAIL(-1, "snb_field_backward: input_dir is required");
    if (L.t_dims && !t_emb) SNB_FAIL(-1, "snb_field_backward: input_t is required");
    if (n_points == 0) return 0;
    if (!workspace) SNB_FAIL(-1, "snb_field_backward: null workspace");
    cudaStream_t st = (cudaStream_t)stream;
    const int pc = n_points < kChunkPoints ? n_points : kChunkPoints, C = L.n_channels, Co = sigma_only ? 1 : C;
    Arena ar(workspace, workspace_bytes); FieldChunk c; c.plan(ar, L, pc, pc, true);
    float* raw = ar.take<float>((size_t)pc * C);
    float* d_head = ar.take<float>((size_t)pc * C);
    if (ar.overflow) SNB_FAIL(-4, "snb_field_backward: workspace too small");
    for (int p0 = 0; p0 < n_points; p0 += pc) {
        const int n = n_points - p0 < pc ? n_points - p0 : pc;
        FieldInputs in; in.n_points = n;
        in.xyz = Src{xyz + (size_t)p0 * 3, 3, 3, 1};
        in.aux = aux_dir ? Src{aux_dir + (size_t)p0 * 3, 3, 3, 1} : Src{nullptr, 0, 0, 1};
        in.temb = (L.t_dims && t_emb) ? Src{t_emb + (size_t)p0 * L.t_dims, L.t_dims, L.t_dims, 1} : Src{nullptr, 0, 0, 1};
        SNB_TRY(field_forward_chunk(L, params, c, in, raw, false, st));            // recompute with the per-layer buffers kept
        field_head_grad_kernel<<<148 * 4, 256, 0, st>>>(d_out + (size_t)p0 * Co, out + (size_t)p0 * Co, d_head, n, C, sigma_only);
        SNB_CHECK_LAUNCH();
        SNB_TRY(field_backward_chunk(L, params, g_params, c, in, d_head, (L.t_dims && g_t_emb) ? g_t_emb + (size_t)p0 * L.t_dims : nullptr, 1, st));
    }
    return 0;
}

#ifdef SNB_DEV_BUILD
#include "satnerf_b200_dev.h"
// Developer aid (libsatnerf_b200_dev.so only): phase timestamps recorded by the fused kernel (see tc_field.cu); host buffer of int64.
extern "C" SNB_API int snb_debug_read(void* host_dst, size_t bytes) { return tc_debug_read(host_dst, bytes); }
extern "C" SNB_API int snb_debug_hang_info(unsigned int* out192) { return tc_debug_hang_info(out192); }
#endif
