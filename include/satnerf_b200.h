/*
 * satnerf_b200 — C-ABI of the B200-native Sat-NeRF volumetric-rendering hot path.
 *
 * The reference (centreborelli/satnerf @ 700a5919) is pure Python and has no FFI layer; its
 * boundary for this path is two Python callables:
 *     rendering.render_rays(models, args, rays, ts)            rendering.py:52
 *     models.<variant>.inference(model, args, xyz, z_vals, …)  models/satnerf.py:4, snerf.py:4, nerf.py:71
 *     <Field>.forward(input_xyz, input_dir, input_sun_dir, input_t)   satnerf.py:156, snerf.py:148, nerf.py:184
 * The entry points below are what a ctypes binding for those callables binds (the binding itself is
 * satnerf_b200/capi.py; INTEGRATION.md shows the stub a reference maintainer would add).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 (int64 where stated), row-major, contiguous;
 *     NULL means "absent" (input) or "not wanted" (output);
 *   - no allocation and no hidden streams inside the library: the caller passes a workspace
 *     (size from snb_render_workspace) and the cudaStream_t (as void*) to launch on;
 *   - every function returns 0 on success; otherwise a negative code, and snb_last_error()
 *     (thread-local) describes it.  Nothing falls back to the CPU.
 *   - parameters live in ONE flat fp32 buffer in the reference's state_dict() order, each
 *     nn.Linear as weight[out][in] followed by bias[out] (snb_param_layout lists the offsets).
 */
#ifndef SATNERF_B200_H
#define SATNERF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SNB_ABI_VERSION 5

#if defined(__GNUC__)
#define SNB_API __attribute__((visibility("default")))
#else
#define SNB_API
#endif

/* args.model (opt.py:34): which field / inference() variant */
enum { SNB_NERF = 0, SNB_SNERF = 1, SNB_SATNERF = 2 };
/* arithmetic of the MLP contractions */
enum { SNB_FP32_SIMT = 0,   /* fp32 FFMA on CUDA cores: the exactness path                        */
       SNB_FP16_TC = 1,     /* fp16 operands / fp32 accumulate on tcgen05 tensor cores (sm_100a);
                               first trunk layer and all head outputs stay fp32                    */
       SNB_FP16X3_TC = 2 }; /* forward: the layer-by-layer path with every wide contraction on the tensor cores at fp16 hi+lo
                               operand precision (3 MMAs per K-step, 2^-22 relative; ~14x the FFMA path); backward: as
                               SNB_FP32_SIMT                                                        */

/* Field architecture = constructor arguments of models.load_model (models/__init__.py:6-15). */
typedef struct snb_field_desc {
    int32_t variant;     /* SNB_NERF | SNB_SNERF | SNB_SATNERF                                  */
    int32_t n_layers;    /* args.fc_layers (8)                                                  */
    int32_t width;       /* args.fc_units  (512)                                                */
    int32_t skip_layer;  /* trunk layer whose input is cat([x, h]) (skips=[4]); -1 = none       */
    int32_t t_dims;      /* args.t_embbeding_tau (sat-nerf), else 0                             */
    int32_t pe_xyz;      /* Mapping frequencies for xyz (nerf: 10), 0 = identity                */
    int32_t pe_dir;      /* Mapping frequencies for the view direction (nerf: 4)                */
} snb_field_desc;

/* One inference() pass (models/satnerf.py:4-79) over R rays x S samples. */
typedef struct snb_pass_desc {
    int32_t n_rays;          /* R                                                               */
    int32_t n_samples;       /* S (coarse: args.n_samples; fine: n_samples + n_importance)      */
    int32_t ray_cols;        /* 8 (nerf) or 11: o(3) d(3) near far [sun(3)]  (rendering.py:62)  */
    int32_t march_along_sun; /* 1 = solar-correction pass, points = o + sun_d*z (rendering.py:104) */
    int32_t precision;       /* SNB_FP32_SIMT | SNB_FP16_TC | SNB_FP16X3_TC                     */
    float   noise_std;       /* args.noise_std; multiplies `noise` (satnerf.py:57-58)           */
    int32_t weights_packed;  /* SNB_FP16_TC forward only: 1 = `workspace` still holds the packed fp16 weight tiles this
                                library wrote there on an earlier call with the SAME parameter values (the caller
                                guarantees both): the per-call repacking (two small kernels) is skipped.  Evaluation
                                loops (batched_inference, eval_satnerf.py:46-66) render many batches per weight set. */
    int32_t flags;           /* SNB_PASS_* bits below                                                              */
    float   t_min;           /* > 0: compositing stops along a ray once the transmittance has dropped below t_min (the
                                remaining weights are written as 0; their sum is < t_min).  0 = exact reference behaviour */
} snb_pass_desc;

/* snb_pass_desc.flags */
enum { SNB_PASS_SINGLE_CTA = 1,   /* tensor-core path: one CTA per 128-point tile instead of CTA pairs (A/B testing; bit-identical) */
       SNB_PASS_NO_BETA    = 2,   /* sat-nerf, inference: the caller does not consume the uncertainty head (create_satnerf_dsm.py:78
                                     uses depth only): beta_from_xyz is not evaluated; io->beta must be NULL and aux_sums[4] is 0  */
       SNB_PASS_SIGMA_ONLY = 4 }; /* inference: only the density is evaluated (the fields' `sigma_only=True`, satnerf.py:184-185):
                                     trunk + sigma head, no feature / colour / sun / sky / uncertainty layers.  Outputs: depth,
                                     weights, transparency, sigma; every other output pointer must be NULL.  Tensor-core path,
                                     s-nerf / sat-nerf                                                                          */

typedef struct snb_render_io {
    /* inputs */
    const float* params;        /* flat parameter buffer of the field                           */
    const float* rays;          /* (R, ray_cols); may be NULL when xyz and aux_dir are given    */
    const float* z_vals;        /* (R, S) sample depths                                         */
    const float* t_emb;         /* (R, t_dims) = models['t'](ts)  (rendering.py:100); sat-nerf  */
    const float* noise;         /* (R, S) standard-normal draws, or NULL for none               */
    const float* xyz;           /* (R,S,3) optional explicit sample positions — the `rays_xyz`
                                   argument of inference() (satnerf.py:4); NULL: o + dir*z      */
    const float* aux_dir;       /* (R,3) optional `sun_d` / `rays_d` argument of inference();
                                   NULL: taken from rays[:, 8:11] / rays[:, 3:6]                */
    /* outputs = the dict returned by inference(); any may be NULL                              */
    float* rgb;                 /* (R,3)                                                        */
    float* depth;               /* (R)                                                          */
    float* weights;             /* (R,S)                                                        */
    float* transparency;        /* (R,S)                                                        */
    float* albedo;              /* (R,S,3)  (nerf: unused)                                      */
    float* sun;                 /* (R,S,1)                                                      */
    float* sky;                 /* (R,S,3)                                                      */
    float* beta;                /* (R,S,1)  sat-nerf only                                       */
    float* sigma;               /* (R,S) densities; stash needed by snb_render_backward         */
    float* nerf_rgb;            /* (R,S,3) per-sample colour of the nerf variant (stash)        */
    void*  stash;               /* training only: activation stash of snb_render_stash_bytes() bytes written by the
                                   tensor-core forward and consumed by its backward; NULL = inference / fp32 path  */
    float* aux_sums;            /* (R,8) per-ray weighted sums the evaluation scripts form from the per-sample outputs
                                   (eval_satnerf.py:125-146): [sum w*sun, sum w*albedo (3), sum w*beta, sum w*sky (3)];
                                   with it an evaluation pass needs none of the (R,S,.) outputs                    */
} snb_render_io;

/* Losses of metrics.py evaluated per ray INSIDE the backward seed: with snb_render_grads.loss set, the
 * upstream gradients g_* are not read; dL/d(rgb, depth, weights, beta, sun) come from the saved forward results and the targets. */
enum { SNB_LOSS_COLOR_MSE = 1,   /* NerfLoss / SNerfLoss colour term: mean((rgb - target)^2)              (metrics.py:8-19, :36-55) */
       SNB_LOSS_COLOR_BETA = 2,  /* SatNerfLoss: uncertainty_aware_loss, beta = sum w*beta + beta_min       (metrics.py:21-25)      */
       SNB_LOSS_DEPTH = 3,       /* DepthLoss: lambda/3 * mean(weight * (depth - target)^2)                 (metrics.py:75-92)      */
       SNB_LOSS_SOLAR = 4 };     /* solar_correction on a march_along_sun pass: lambda/3 * (mean sum (T - s)^2 + mean (1 - sum w s)),
                                    T and w detached                                                          (metrics.py:27-34)      */
typedef struct snb_loss_desc {
    int32_t kind;               /* SNB_LOSS_*                                                                  */
    int32_t n_rays_mean;        /* denominator of the reference's mean(): rays of the GLOBAL batch (all ranks)  */
    float   lambda;             /* lambda_sc (SOLAR) / lambda_ds (DEPTH); unused otherwise                      */
    float   beta_min;           /* COLOR_BETA: 0.05 (metrics.py:21)                                             */
    const float* target;        /* (R,3) colours | (R) depths; unused for SOLAR                                */
    const float* target_weight; /* DEPTH: (R) per-ray weights or NULL (= 1)                                    */
    const float* g_terms;       /* (4) device: upstream gradient of each loss term below (autograd), NULL = 1   */
} snb_loss_desc;

typedef struct snb_render_grads {
    /* upstream gradients w.r.t. the outputs above; NULL = zero */
    const float* g_rgb;  const float* g_depth;  const float* g_weights;  const float* g_transparency;
    const float* g_albedo;  const float* g_sun;  const float* g_sky;  const float* g_beta;
    /* results */
    float* g_params;            /* flat, same layout as params; ACCUMULATED into (+=)           */
    float* g_t_emb;             /* (R, t_dims), overwritten; NULL if not wanted                 */
    const snb_loss_desc* loss;  /* non-NULL: fused loss seed (the g_* inputs above are ignored) */
} snb_render_grads;

SNB_API int         snb_abi_version(void);
SNB_API const char* snb_last_error(void);
/* number of CUDA kernels this library has launched since the last reset (bench accounting) */
SNB_API int64_t     snb_launch_count(int reset);
/* 1 when the device of the current context can run the tcgen05 path (compute capability 10.x). */
SNB_API int         snb_device_supports_tc(void);

/* Flat parameter layout: returns the number of nn.Linear layers L (or <0); fills up to `cap`
 * entries: weight offset, bias offset (in floats), rows (=out) and cols (=in).
 * Order = state_dict() order of models/satnerf.py:104-153 / snerf.py / nerf.py:157-177. */
SNB_API int     snb_param_layout(const snb_field_desc* f, int64_t* w_off, int64_t* b_off,
                         int32_t* n_out, int32_t* n_in, int cap);
SNB_API int64_t snb_param_count(const snb_field_desc* f);

/* rendering.py:65-78: z = lower + (upper-lower)*u over the stratified bins of [near, far].
 * steps = torch.linspace(0,1,S) (S floats, device); u (R,S) uniform draws; z (R,S) out.        */
SNB_API int snb_stratified_depths(const float* rays, int ray_cols, const float* steps, const float* u,
                          float* z, int n_rays, int n_samples, void* stream);

/* rendering.py:121-125 + sample_pdf (:10-49): importance-sample n_imp depths from the coarse
 * weights and merge them (sorted) with the coarse depths.  u (R,n_imp) uniform draws.
 * z_out (R, S+n_imp).  Optional debug outputs: inds (R,n_imp) int64 = searchsorted result (:36),
 * z_new (R,n_imp) unsorted samples, cdf (R,S-1).                                               */
SNB_API int snb_importance_depths(const float* z_coarse, const float* weights_coarse, const float* u,
                          float* z_out, int64_t* inds, float* z_new, float* cdf,
                          int n_rays, int n_samples, int n_imp, void* stream);

/* searchsorted(cdf, u, right=True) alone (rendering.py:36) — pure comparisons, bit-exact.      */
SNB_API int snb_searchsorted_right(const float* cdf, const float* u, int64_t* inds,
                           int n_rays, int n_cdf, int n_u, void* stream);

/* Workspace (bytes) needed by forward / backward of a pass. */
SNB_API int snb_render_workspace(const snb_field_desc* f, const snb_pass_desc* p, int backward, size_t* bytes);

/* Bytes of the activation stash the tensor-core forward fills for its backward (0 when the pass runs on the
 * fp32 path, which recomputes instead). */
SNB_API int snb_render_stash_bytes(const snb_field_desc* f, const snb_pass_desc* p, size_t* bytes);

/* inference(): field at every sample + alpha compositing (satnerf.py:4-79, snerf.py:4-75, nerf.py:71-133) */
SNB_API int snb_render_forward(const snb_field_desc* f, const snb_pass_desc* p, const snb_render_io* io,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Gradient of the same pass w.r.t. the flat parameters and t_emb (what autograd derives for the
 * reference).  `io` must carry the forward's inputs and saved outputs
 * (weights, transparency, albedo, sun, sky, beta, sigma; nerf: nerf_rgb).                      */
SNB_API int snb_render_backward(const snb_field_desc* f, const snb_pass_desc* p, const snb_render_io* io,
                        const snb_render_grads* g, void* workspace, size_t workspace_bytes, void* stream);

/* Loss terms of one pass from its forward results (what metrics.py computes from the result dict): terms (4) device floats
 *   COLOR_MSE: [colour, 0, 0, 0]   COLOR_BETA: [colour, logbeta, 0, 0]   DEPTH: [ds, 0, 0, 0]   SOLAR: [0, 0, sc_term2, sc_term3]
 * io carries rgb / depth / weights / beta (SOLAR: transparency, weights, sun of the march_along_sun pass).  Deterministic
 * (fixed-order reduction).  workspace: n_rays * 16 bytes.                                                               */
SNB_API int snb_loss_forward(const snb_pass_desc* p, const snb_render_io* io, const snb_loss_desc* loss, float* terms,
                     void* workspace, size_t workspace_bytes, void* stream);

/* Gradient of those terms w.r.t. the result-dict tensors they read (autograd of the loss classes): any output may be NULL.
 * g_rgb (R,3), g_depth (R), g_weights (R,S), g_beta (R,S), g_sun (R,S); loss->g_terms = upstream gradient of the 4 terms. */
SNB_API int snb_loss_backward(const snb_pass_desc* p, const snb_render_io* io, const snb_loss_desc* loss,
                      float* g_rgb, float* g_depth, float* g_weights, float* g_beta, float* g_sun, void* stream);

/* <Field>.forward on B independent points (satnerf.py:156-208): xyz (B,3); aux_dir (B,3) = sun
 * direction (sat-nerf / s-nerf) or view direction (nerf); t_emb (B,t_dims); out (B,C) with
 * C = 9 / 8 / 4 = [rgb3, sigma, sun, sky3, beta]; sigma_only -> out (B,1) (satnerf.py:184-185).
 * precision: SNB_FP32_SIMT or SNB_FP16X3_TC (tensor cores, fp16 hi+lo operands; per-POINT sun / embedding inputs). */
SNB_API int snb_field_workspace(const snb_field_desc* f, int n_points, size_t* bytes);
SNB_API int snb_field_forward(const snb_field_desc* f, const float* params, const float* xyz,
                      const float* aux_dir, const float* t_emb, float* out, int n_points,
                      int sigma_only, int precision, void* workspace, size_t workspace_bytes, void* stream);

/* Gradient of snb_field_forward (autograd of <Field>.forward, satnerf.py:156-208): d_out (B,C | B,1) = gradient w.r.t. the
 * outputs, out = the outputs the forward returned; ACCUMULATES into g_params (flat, like the parameters) and writes g_t_emb
 * (B,t_dims; optional).  fp32 CUDA-core path (the forward is recomputed chunk by chunk with its buffers kept).  No gradient
 * w.r.t. xyz / directions (the render path never needs one).                                                          */
SNB_API int snb_field_backward_workspace(const snb_field_desc* f, int n_points, size_t* bytes);
SNB_API int snb_field_backward(const snb_field_desc* f, const float* params, const float* xyz, const float* aux_dir, const float* t_emb,
                       const float* out, const float* d_out, float* g_params, float* g_t_emb, int n_points, int sigma_only,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ---- geometry either side of the render path -------------------------------------------------
 * RPC camera model of one satellite image: the fields of rpcm.RPCModel (projection coefficients in RPC00B order; the
 * reference builds it from the image's json, datasets/satellite.py:191) after sat_utils.rescale_rpc.                  */
typedef struct snb_rpc_model {
    double row_num[20], row_den[20], col_num[20], col_den[20];
    double row_offset, row_scale, col_offset, col_scale, lat_offset, lat_scale, lon_offset, lon_scale, alt_offset, alt_scale;
} snb_rpc_model;

/* datasets/satellite.py:18-65 get_rays (+ :218-227 normalize_rays when center3 != NULL, + :229-244 sun direction when
 * sun_dir3 != NULL): rays (n_pixels, ray_cols = 8 | 11) fp32 DEVICE = [origin, unit direction, near = 0, far, (sun)].
 * Pixels: cols / rows (n_pixels doubles each, DEVICE) or NULL for the row-major grid of `width` columns
 * (np.meshgrid(arange(w), arange(h)), :193).  rpc, center3, sun_dir3 are HOST pointers (copied into the launch).
 * max_iters (DEVICE int, optional, zeroed by the caller) receives the largest iteration count of the inverse RPC
 * (> 100: rpcm raises MaxLocalizationIterationsError).                                                              */
SNB_API int snb_rpc_rays(const snb_rpc_model* rpc, const double* cols, const double* rows, int width, long long n_pixels,
                 double min_alt, double max_alt, const double* center3, double range, const float* sun_dir3,
                 float* rays, int ray_cols, int* max_iters, void* stream);

/* datasets/satellite.py:246-274 get_latlonalt_from_nerf_prediction + sat_utils.utm_from_latlon (:99-113): cloud (n,3)
 * doubles DEVICE = [east, north, alt] in UTM zone `utm_zone` (<= 0: [lon, lat, alt], no projection); latlon (n,2)
 * optional [lat, lon] degrees.  center3 HOST.                                                                       */
SNB_API int snb_dsm_points(const float* rays, int ray_cols, const float* depth, long long n_rays, const double* center3, double range,
                   int utm_zone, double* cloud, double* latlon, void* stream);

/* plyflatten(cloud, xoff, yoff, resolution, xsize, ysize, radius, sigma=inf) as called at datasets/satellite.py:308:
 * dsm (ysize, xsize) fp32 DEVICE, mean altitude of the points within `radius` cells of each cell, NaN where none.
 * Deterministic (fixed-point sums).  workspace: xsize * ysize * 12 + 512 bytes.                                     */
SNB_API int snb_dsm_rasterize(const double* cloud, long long n_points, double xoff, double yoff, double resolution, int xsize, int ysize,
                      int radius, float* dsm, void* workspace, size_t workspace_bytes, void* stream);

/* Batch assembly of the GPU-resident ray sampler (replaces the per-ray DataLoader of main.py:96-110 over the dataset's
 * __getitem__, datasets/satellite.py:347-350 / satellite_depth.py:138-141): outs[t][i, :] = tables[t][idx[i], :] for up to 4
 * row-major DEVICE tables (all_rays, all_rgbs | all_depths, all_ids ...) that share the int64 index vector idx (n_rows, DEVICE);
 * row_bytes[t] multiples of 4.  tables / outs / row_bytes are HOST arrays.  One launch.                                   */
SNB_API int snb_gather_rows(const void* const* tables, void* const* outs, const int32_t* row_bytes, int n_tables,
                    const int64_t* idx, long long n_rows, long long n_src_rows, void* stream);

/* Optimiser step of the training loop on a flat parameter buffer (main.py:81-94 builds torch.optim.Adam(lr, weight_decay=0)
 * through train_utils.py:24-53; Lightning calls its step after every training_step): torch.optim.Adam arithmetic (amsgrad
 * off, L2 weight decay folded into the gradient; hyper-parameters as doubles like torch's Python scalars), `step` = 1-based
 * count of this update.  All buffers n floats, 16-byte
 * aligned; updates params / exp_avg / exp_avg_sq in place.                                                          */
SNB_API int snb_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                  double lr, double beta1, double beta2, double eps, double weight_decay, int step, void* stream);

/* Data-parallel optimiser step fused with its collective (multi-GPU training; the reference is single-GPU): rank
 * `rank` of `world` (<= 8) owns a contiguous shard of the n floats (n % 4 == 0); for its shard it sums the gradient shards of ALL
 * ranks through peer_grads[r] (device pointers valid on this GPU: symmetric / peer-mapped memory over NVLink; fixed order), applies
 * snb_adam_step's arithmetic with its own moments, and stores the new parameters into peer_params[r] of every rank.  peer_params /
 * peer_grads are HOST arrays of `world` device pointers.  The caller provides the two cross-rank barriers around it (all gradients
 * final before; all parameters delivered after).  Replaces all-reduce + Adam: replicas stay bit-identical.                    */
SNB_API int snb_adam_step_sharded(float* const* peer_params, const float* const* peer_grads, int world, int rank,
                          float* exp_avg, float* exp_avg_sq, long long n,
                          double lr, double beta1, double beta2, double eps, double weight_decay, int step, void* stream);

/* The same step for up to 4 flat buffers (fields, embedding table) in ONE launch.  mc_params / mc_grads: multicast (NVLS) addresses of
 * the symmetric allocations, or NULL: with them the gradient sum is one `multimem.ld_reduce` per 16 bytes (reduced inside the
 * NVSwitch) and the parameter delivery one `multimem.st`; without them peer loads / stores as in snb_adam_step_sharded.  The
 * in-switch sum has no specified association order: replicas stay bit-identical (every element is reduced once, by its owner), the
 * result may differ from the fixed-order sum in the last bit.                                                              */
typedef struct snb_sharded_buffer {
    float* const* peer_params;      /* host array of `world` device pointers (valid on this GPU)  */
    const float* const* peer_grads;
    float* mc_params;               /* multicast address of the parameter buffers, or NULL        */
    const float* mc_grads;          /* multicast address of the gradient buffers, or NULL         */
    float* exp_avg;                 /* this rank's moments (n floats; only its shard is touched)  */
    float* exp_avg_sq;
    long long n;                    /* floats in the buffer, a multiple of 4                      */
    int32_t step;                   /* 1-based count of this update (bias correction)             */
} snb_sharded_buffer;
SNB_API int snb_adam_step_sharded_multi(const snb_sharded_buffer* bufs, int n_buffers, int world, int rank,
                                double lr, double beta1, double beta2, double eps, double weight_decay, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SATNERF_B200_H */
