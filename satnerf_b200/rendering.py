"""Drop-in for the reference's `rendering.py`: `render_rays(models, args, rays, ts)` and `sample_pdf`,
plus the per-variant `inference()` (models/satnerf.py:4, snerf.py:4, nerf.py:71).

Python here only sequences passes and draws the random tensors in the reference's order
(rand_like -> randn [-> randn for the solar-correction pass] -> rand -> randn ...);
sampling, the MLP, compositing and their gradients run in libsatnerf_b200.so through the C ABI.

`args` is the reference's argparse Namespace.  Extra, optional attributes understood here:
  args.precision : 'tc'  (default on sm_100: fp16 operands / fp32 accumulate on tcgen05 tensor cores)
                   'fp32' (fp32 FFMA CUDA-core path, matches the reference to rounding level)
                   'tcx3' (layer-by-layer path with every wide contraction on the tensor cores at fp16 hi+lo operand precision:
                           fp32-level results -- for checkpoints whose optical depths push the fp16-operand kernel past 1e-3)
  args.render_outputs : 'full' (default: the reference's result dict)
                   'eval'  (no_grad only) per-ray outputs: rgb_*, depth_* and the weighted images eval_satnerf.py:125-146 builds
                           from the per-sample tensors -- sun_w_*, albedo_w_*, beta_w_*, sky_w_* = sum_i w_i x_i -- computed in-kernel
                   'depth' (no_grad only) rgb_*, depth_* only; sat-nerf skips the uncertainty head (create_satnerf_dsm.py:78)
                   'depth_only' (no_grad only, tensor-core path) depth_* only: density trunk + sigma head, what the fields'
                           `sigma_only=True` evaluates (satnerf.py:184-185) -- all a DSM needs
  args.t_min     : > 0 stops compositing along a ray once its transmittance is below t_min (dropped weights sum to < t_min)
  args.tc_cta_group : 1 forces single-CTA tiles on the tensor-core path (default 2: CTA pairs; results are bit-identical)
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import capi

_SAVED = ("weights", "transparency", "albedo", "sun", "sky", "beta", "sigma", "nerf_rgb")


def _precision(args) -> int:
    p = getattr(args, "precision", None)
    if p is None:
        p = "tc" if capi.device_supports_tc() else "fp32"
    if p not in ("tc", "fp32", "tcx3"):
        raise ValueError(f"precision {p!r} is not valid (tc | tcx3 | fp32)")
    return {"tc": capi.FP16_TC, "tcx3": capi.FP16X3_TC, "fp32": capi.FP32_SIMT}[p]


def _out_shapes(variant: str, R: int, S: int):
    sh = {"rgb": (R, 3), "depth": (R,), "weights": (R, S), "transparency": (R, S)}
    if variant != "nerf":
        sh.update(albedo=(R, S, 3), sun=(R, S, 1), sky=(R, S, 3))
    if variant == "sat-nerf":
        sh["beta"] = (R, S, 1)
    return sh


def _pass_desc(cfg, R, S, ray_cols) -> capi.PassDesc:
    return capi.PassDesc(R, S, ray_cols, int(cfg["sc"]), cfg["precision"], float(cfg["noise_std"]), 0, int(cfg.get("flags", 0)), float(cfg.get("t_min", 0.0)))


def _packed_workspace(field, pd, params, dev):
    """Inference on the tensor-core path: the packed fp16 weight tiles live in a buffer owned by the field and are reused while
    the parameters are untouched (storage, autograd version counters and `field.invalidate_packed()` epoch unchanged) --
    batched_inference / DSM extraction render many ray batches per weight set.  Sets pd.weights_packed; returns the buffer."""
    flat = field.flat_params()
    # (the parameters alias `flat` through `.data`, so each keeps its own version counter)
    key = (flat.data_ptr(), flat._version, tuple(p._version for p in params), str(dev), field._packed_epoch,
           pd.flags & (capi.PASS_NO_BETA | capi.PASS_SIGMA_ONLY))
    need = capi.render_workspace_bytes(field.desc, pd)
    cache = getattr(field, "_tc_packed", None)
    if cache is None or cache[0].device != dev or cache[0].numel() < need:
        cache = [torch.empty(max(need, 1), dtype=torch.uint8, device=dev), None]
        field._tc_packed = cache
    pd.weights_packed = 1 if (need and cache[1] == key) else 0
    cache[1] = key
    return cache[0]


class _Pass(torch.autograd.Function):
    """One inference() pass.  Differentiable w.r.t. the field parameters and the per-ray embedding."""

    @staticmethod
    def forward(ctx, field, cfg, rays, z, t_emb, noise, xyz, aux_dir, *params):
        variant = field.variant
        R, S = z.shape
        dev = z.device
        ctx.set_materialize_grads(False)       # outputs the loss does not touch arrive as None in backward (no zero-filled (R,S,.) tensors)
        pd = _pass_desc(cfg, R, S, rays.shape[1] if rays is not None else 0)
        outs = {k: torch.empty(s, device=dev, dtype=torch.float32) for k, s in _out_shapes(variant, R, S).items()}
        stash = {"sigma": torch.empty(R, S, device=dev, dtype=torch.float32)}
        if variant == "nerf":
            stash["nerf_rgb"] = torch.empty(R, S, 3, device=dev, dtype=torch.float32)
        # training: the tensor-core forward stashes its activations for the tensor-core backward (the fp32 path recomputes)
        act_stash = None
        if cfg["train"] and cfg["precision"] == capi.FP16_TC:
            nbytes = capi.render_stash_bytes(field.desc, pd)
            if nbytes:
                act_stash = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        tensors = dict(params=field.flat_params(), rays=rays, z_vals=z, t_emb=t_emb,
                       noise=noise if cfg["noise_std"] != 0 else None, xyz=xyz, aux_dir=aux_dir, stash=act_stash, **outs, **stash)
        ws = None
        if not cfg["train"] and cfg["precision"] == capi.FP16_TC:
            ws = _packed_workspace(field, pd, params, dev)
        capi.render_forward(field.desc, pd, tensors, workspace=ws)
        pd.weights_packed = 0
        ctx.act_stash = act_stash
        ctx.field, ctx.pd, ctx.variant = field, pd, variant
        ctx.keys = list(outs)
        ctx.use_noise = cfg["noise_std"] != 0
        ctx.has = (rays is not None, t_emb is not None, noise is not None, xyz is not None, aux_dir is not None)
        opt = [t for t in (rays, t_emb, noise, xyz, aux_dir) if t is not None]
        saved = [outs.get(k, stash.get(k)) for k in _SAVED if k in outs or k in stash]
        ctx.saved_names = [k for k in _SAVED if k in outs or k in stash]
        ctx.save_for_backward(z, *opt, *saved)
        return tuple(outs[k] for k in ctx.keys)

    @staticmethod
    def backward(ctx, *gouts):
        field = ctx.field
        sv = list(ctx.saved_tensors)
        z = sv.pop(0)
        rays, t_emb, noise, xyz, aux_dir = (sv.pop(0) if h else None for h in ctx.has)
        saved = dict(zip(ctx.saved_names, sv))
        flat = field.flat_params()
        g_flat = torch.zeros_like(flat)
        g_t = torch.empty_like(t_emb) if t_emb is not None else None
        tensors = dict(params=flat, rays=rays, z_vals=z, t_emb=t_emb, noise=noise if ctx.use_noise else None,
                       xyz=xyz, aux_dir=aux_dir, stash=ctx.act_stash, **saved)
        grads = {"g_params": g_flat, "g_t_emb": g_t}
        for k, g in zip(ctx.keys, gouts):
            grads["g_" + k] = None if g is None else g.to(torch.float32).contiguous()
        capi.render_backward(field.desc, ctx.pd, tensors, grads)
        field._flat_grad = g_flat       # candidate flat gradient buffer (flat_grads() checks that .grad really aliases it)
        gp, off = [], 0
        for p in field.ordered_params():
            n = p.numel()
            gp.append(g_flat[off:off + n].view(p.shape))
            off += n
        return (None, None, None, None, g_t, None, None, None, *gp)


_LITE_KEYS = {"depth": ("rgb", "depth"), "eval": ("rgb", "depth", "sun_w", "albedo_w", "beta_w", "sky_w")}


def _run_pass_lite(field, args, rays, z, t_emb, noise, mode, want_weights=False) -> Dict[str, torch.Tensor]:
    """Evaluation pass with per-RAY outputs only: rgb, depth and -- mode 'eval' -- the weighted sums
    sum_i w_i * {sun, albedo, beta, sky} that eval_satnerf.py:125-146 forms from the (R,S,.) tensors; none of those tensors is
    written.  mode 'depth' (create_satnerf_dsm.py:78 consumes the depth only) also skips the uncertainty head."""
    if torch.is_grad_enabled() and any(p.requires_grad for p in field.parameters()):
        raise RuntimeError("render_outputs='eval'/'depth'/'depth_only' are inference modes: call under torch.no_grad()")
    if field.variant == "nerf":
        raise NotImplementedError("render_outputs='eval'/'depth' exist for s-nerf / sat-nerf (the outputs of eval_satnerf.py:125-146)")
    R, S = z.shape
    dev = z.device
    flags = _flags(args) | (capi.PASS_NO_BETA if (mode == "depth" and field.variant == "sat-nerf") else 0)
    if mode == "depth_only":
        if _precision(args) != capi.FP16_TC:
            raise NotImplementedError("render_outputs='depth_only' runs on the tensor-core path (args.precision='tc')")
        flags |= capi.PASS_SIGMA_ONLY
    cfg = {"sc": False, "precision": _precision(args), "noise_std": float(args.noise_std), "flags": flags, "t_min": float(getattr(args, "t_min", 0.0))}
    pd = _pass_desc(cfg, R, S, rays.shape[1])
    f32 = lambda t: None if t is None else t.to(torch.float32).contiguous()
    outs = {"depth": torch.empty(R, device=dev)}
    if mode != "depth_only":
        outs["rgb"] = torch.empty(R, 3, device=dev)
    if want_weights:
        outs["weights"] = torch.empty(R, S, device=dev)
    aux = torch.empty(R, 8, device=dev) if mode == "eval" else None
    tensors = dict(params=field.flat_params(), rays=f32(rays), z_vals=f32(z), t_emb=f32(t_emb), noise=f32(noise) if cfg["noise_std"] != 0 else None,
                   aux_sums=aux, **outs)
    ws = _packed_workspace(field, pd, field.ordered_params(), dev) if cfg["precision"] == capi.FP16_TC else None
    capi.render_forward(field.desc, pd, tensors, workspace=ws)
    if aux is not None:
        outs.update(sun_w=aux[:, 0:1], albedo_w=aux[:, 1:4], sky_w=aux[:, 5:8])
        if field.variant == "sat-nerf":
            outs["beta_w"] = aux[:, 4:5]
    return outs


def _flags(args) -> int:
    return capi.PASS_SINGLE_CTA if getattr(args, "tc_cta_group", 2) == 1 else 0


def _run_pass(field, args, rays, z, t_emb, noise, sc=False, xyz=None, aux_dir=None) -> Dict[str, torch.Tensor]:
    if not z.is_cuda:
        raise RuntimeError("satnerf_b200 renders CUDA tensors only (no CPU fallback); move rays and models to the GPU")
    # stash activations only when a backward can follow (Function.forward itself always runs with grad mode off)
    train = torch.is_grad_enabled() and (any(p.requires_grad for p in field.parameters()) or (t_emb is not None and t_emb.requires_grad))
    cfg = {"sc": sc, "precision": _precision(args), "noise_std": float(args.noise_std), "train": train, "flags": _flags(args),
           "t_min": float(getattr(args, "t_min", 0.0))}
    f32 = lambda t: None if t is None else t.to(torch.float32).contiguous()
    outs = _Pass.apply(field, cfg, f32(rays), f32(z), f32(t_emb), f32(noise), f32(xyz), f32(aux_dir), *field.ordered_params())
    return dict(zip(_out_shapes(field.variant, z.shape[0], z.shape[1]).keys(), outs))


# ------------------------------------------------------------------------------------------------
# public API (reference signatures)
# ------------------------------------------------------------------------------------------------
def inference(model, args, rays_xyz, z_vals, rays_d=None, sun_d=None, rays_t=None):
    """models/satnerf.py:4 / snerf.py:4 / nerf.py:71 — explicit sample positions."""
    variant = model.variant
    aux = rays_d if variant == "nerf" else sun_d
    if aux is None:
        raise TypeError("rays_d is required" if variant == "nerf" else "sun_d is required")
    if variant == "sat-nerf" and rays_t is None:
        raise TypeError("sat-nerf needs rays_t (models/satnerf.py:204)")
    noise = torch.randn(z_vals.shape, device=z_vals.device)                      # satnerf.py:58
    return _run_pass(model, args, None, z_vals, rays_t if variant == "sat-nerf" else None, noise,
                     xyz=rays_xyz, aux_dir=aux)


def sample_pdf(bins, weights, N_importance, det=False, eps=1e-5):
    """rendering.py:10-49.  Returns the (unsorted) importance samples for explicit (bins, weights).

    render_rays does not come through here: it uses the fused importance kernel (snb_importance_depths),
    which builds bins, pdf and cdf from the coarse depths / weights itself.  This entry point keeps the
    reference's stand-alone signature; its bin search is the library's bit-exact snb_searchsorted_right."""
    if eps != 1e-5:
        raise NotImplementedError("eps is fixed to 1e-5 (the only value the reference uses)")
    if not bins.is_cuda:
        raise RuntimeError("satnerf_b200 runs on CUDA tensors only (no CPU fallback)")
    R, M = weights.shape
    u = (torch.linspace(0, 1, N_importance, device=bins.device).expand(R, N_importance) if det
         else torch.rand(R, N_importance, device=bins.device)).contiguous()
    w = weights.to(torch.float32) + eps
    pdf = w / w.sum(-1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[:, :1]), torch.cumsum(pdf, -1)], -1).contiguous()
    k = capi.searchsorted_right(cdf, u)
    below, above = (k - 1).clamp_min(0), k.clamp_max(M)
    c0, c1 = cdf.gather(1, below), cdf.gather(1, above)
    b0, b1 = bins.gather(1, below), bins.gather(1, above)
    den = c1 - c0
    den = torch.where(den < eps, torch.ones_like(den), den)
    return b0 + (u - c0) / den * (b1 - b0)


_LINSPACE = {}


def _linspace01(S: int, dev) -> torch.Tensor:
    """torch.linspace(0, 1, S) of rendering.py:65 — deterministic, so one tensor per (S, device) serves every call."""
    key = (S, str(dev))
    t = _LINSPACE.get(key)
    if t is None:
        t = _LINSPACE[key] = torch.linspace(0, 1, S, device=dev)
    return t


def render_rays(models, args, rays, ts, _draws: Optional[List[torch.Tensor]] = None):
    """rendering.py:52-158.  Same arguments and result dict as the reference.

    `_draws` (tests only) replays recorded random tensors instead of drawing them."""
    variant = args.model
    if variant not in capi.VARIANTS:
        raise ValueError(f"model {variant} is not valid")
    S, n_imp = args.n_samples, args.n_importance
    sc = variant != "nerf" and args.sc_lambda > 0
    if n_imp > 0 and (variant == "s-nerf" or sc):
        # the reference crashes here (NameError rendering.py:134 / result dict overwritten :138,:149)
        raise NotImplementedError("fine pass with s-nerf or with solar correction is broken in the reference and not provided")
    if not rays.is_cuda:
        raise RuntimeError("satnerf_b200 renders CUDA tensors only (no CPU fallback)")
    R, dev = rays.shape[0], rays.device
    tape = None if _draws is None else [t.to(dev) for t in _draws]

    def draw(kind, *shape):
        if tape is not None:
            t = tape.pop(0)
            assert tuple(t.shape) == tuple(shape)
            return t
        return torch.rand(*shape, device=dev) if kind == "u" else torch.randn(*shape, device=dev)

    rays = rays.to(torch.float32).contiguous()
    if variant == "sat-nerf" and ts is None:
        raise TypeError("sat-nerf needs ts (the reference fails in torch.cat at models/satnerf.py:204)")
    t_emb = models["t"](ts) if variant == "sat-nerf" else None                    # rendering.py:100
    steps = _linspace01(S, dev)                                                   # :65
    z = capi.stratified_depths(rays, steps, draw("u", R, S).contiguous())         # :67-78

    mode = getattr(args, "render_outputs", "full")
    if mode not in ("full", "eval", "depth", "depth_only"):
        raise ValueError(f"render_outputs {mode!r} is not valid (full | eval | depth | depth_only)")

    def level(name, zz):
        if mode != "full":                     # per-ray outputs only; the solar-correction pass feeds the loss alone and is not evaluated
            res = _run_pass_lite(models[name], args, rays, zz, t_emb, draw("n", R, zz.shape[1]), mode, want_weights=(name == "coarse" and n_imp > 0))
            if sc:
                draw("n", R, zz.shape[1])      # keep the generator position of the reference (rendering.py:95/:107 draws even when unused)
            return {f"{k}_{name}": v for k, v in res.items()}
        res = _run_pass(models[name], args, rays, zz, t_emb, draw("n", R, zz.shape[1]))          # :89/:101/:112
        if sc:                                                                                    # :90-96 / :102-108
            r2 = _run_pass(models[name], args, rays, zz, t_emb, draw("n", R, zz.shape[1]), sc=True)
            res["weights_sc"], res["transparency_sc"], res["sun_sc"] = r2["weights"], r2["transparency"], r2["sun"]
        return {f"{k}_{name}": v for k, v in res.items()}                                         # :113-115

    result = level("coarse", z)
    if n_imp > 0:                                                                                 # :118-156
        u = draw("u", R, n_imp).contiguous()
        z_fine = capi.importance_depths(z, result["weights_coarse"].detach().contiguous(), u)    # :121-125
        result.update(level("fine", z_fine))
    return result


# ------------------------------------------------------------------------------------------------
# training: render + loss + backward in one go
# ------------------------------------------------------------------------------------------------
def render_loss_backward(models, args, rays, ts, color=None, depth=None, n_rays_mean=None, backward=True,
                         _draws: Optional[List[torch.Tensor]] = None):
    """`render_rays` + the losses of metrics.py + `loss.backward()` of one ray batch without autograd: the loss gradient is formed
    per ray inside the compositing backward (snb_render_grads.loss), parameter gradients are ACCUMULATED into each field's flat
    gradient buffer (`field.flat_grads()`, aliased by the parameters' `.grad`) and the embedding's `.grad`.

      color = (kind, target_rgb (R,3))   kind 'mse' (NerfLoss / SNerfLoss, metrics.py:8-55) | 'beta' (SatNerfLoss, :57-73);
                                         with args.sc_lambda > 0 the solar-correction terms (:27-34) ride on the second pass
      depth = (target (R), weight (R) | None, lambda_ds)     DepthLoss (:75-92) on this batch instead of a colour loss
      n_rays_mean   denominator of the reference's mean(): rays of the GLOBAL batch (default: this call's R) -- a rank of a
                    data-parallel job passes world * R and sum-all-reduces the gradients afterwards (no division)
      backward=False evaluates the loss terms only.

    Same RNG draw order as render_rays.  Returns (loss_dict of 0-dim tensors keyed like metrics.py, {'rgb_<level>', 'depth_<level>'})."""
    variant = args.model
    if variant not in capi.VARIANTS:
        raise ValueError(f"model {variant} is not valid")
    if (color is None) == (depth is None):
        raise ValueError("give exactly one of color= / depth=")
    S, n_imp = args.n_samples, args.n_importance
    sc = variant != "nerf" and args.sc_lambda > 0
    if n_imp > 0 and (variant == "s-nerf" or sc):
        raise NotImplementedError("fine pass with s-nerf or with solar correction is broken in the reference and not provided")
    if color is not None and color[0] == "beta" and n_imp > 0:
        raise RuntimeError("SatNerfLoss on a coarse+fine sat-nerf result fails in the reference (beta_coarse vs weights_fine, SURVEY.md App. B)")
    if not rays.is_cuda:
        raise RuntimeError("satnerf_b200 renders CUDA tensors only (no CPU fallback)")
    R, dev = rays.shape[0], rays.device
    n_mean = int(n_rays_mean or R)
    tape = None if _draws is None else [t.to(dev) for t in _draws]

    def draw(kind, *shape):
        if tape is not None:
            t = tape.pop(0)
            assert tuple(t.shape) == tuple(shape)
            return t
        return torch.rand(*shape, device=dev) if kind == "u" else torch.randn(*shape, device=dev)

    rays = rays.to(torch.float32).contiguous()
    if variant == "sat-nerf" and ts is None:
        raise TypeError("sat-nerf needs ts (the reference fails in torch.cat at models/satnerf.py:204)")
    emb = models["t"] if variant == "sat-nerf" else None
    t_emb = emb.weight.detach()[ts].contiguous() if emb is not None else None                # rendering.py:100
    g_t_total = None
    z = capi.stratified_depths(rays, _linspace01(S, dev), draw("u", R, S).contiguous())
    precision, noise_std = _precision(args), float(args.noise_std)
    loss_dict, results = {}, {}
    f32c = lambda t: None if t is None else t.to(device=dev, dtype=torch.float32).contiguous()

    def one_pass(field, zz, noise, march_sun, loss):
        """forward (training mode) -> loss terms -> backward seeded with the loss; returns (terms (4,), outs)."""
        nonlocal g_t_total
        Sz = zz.shape[1]
        cfg = {"sc": march_sun, "precision": precision, "noise_std": noise_std, "flags": _flags(args)}
        pd = _pass_desc(cfg, R, Sz, rays.shape[1])
        outs = {k: torch.empty(s, device=dev, dtype=torch.float32) for k, s in _out_shapes(variant, R, Sz).items()}
        outs["sigma"] = torch.empty(R, Sz, device=dev, dtype=torch.float32)
        if variant == "nerf":
            outs["nerf_rgb"] = torch.empty(R, Sz, 3, device=dev, dtype=torch.float32)
        stash = None
        if backward and precision == capi.FP16_TC:
            nbytes = capi.render_stash_bytes(field.desc, pd)
            stash = torch.empty(nbytes, dtype=torch.uint8, device=dev) if nbytes else None
        tensors = dict(params=field.flat_params(), rays=rays, z_vals=zz, t_emb=t_emb, noise=noise if noise_std != 0 else None, stash=stash, **outs)
        capi.render_forward(field.desc, pd, tensors)
        terms = capi.loss_forward(pd, tensors, loss)
        if backward:
            g_t = torch.empty_like(t_emb) if t_emb is not None else None
            capi.render_backward(field.desc, pd, tensors, {"g_params": field.flat_grads(zero=False), "g_t_emb": g_t}, loss=loss)
            if g_t is not None:
                g_t_total = g_t if g_t_total is None else g_t_total.add_(g_t)
        return terms, outs

    def level(name, zz):
        field = models[name]
        if color is not None:
            kind = capi.LOSS_COLOR_BETA if color[0] == "beta" else capi.LOSS_COLOR_MSE
            loss = capi.loss_desc(kind, n_mean, f32c(color[1]))
        else:
            loss = capi.loss_desc(capi.LOSS_DEPTH, n_mean, f32c(depth[0]), f32c(depth[1]), None, float(depth[2]))
        terms, outs = one_pass(field, zz, draw("n", R, zz.shape[1]), False, loss)
        if color is not None:
            loss_dict[f"{name}_color"] = terms[0]
            if color[0] == "beta":
                loss_dict[f"{name}_logbeta"] = terms[1]
        else:
            loss_dict[f"{name}_ds"] = terms[0]
        results[f"rgb_{name}"], results[f"depth_{name}"] = outs["rgb"], outs["depth"]
        if sc:                                                                                  # rendering.py:90-96 / :102-108
            noise2 = draw("n", R, zz.shape[1])
            if color is not None:                                                               # (DepthLoss ignores the SC outputs)
                t2, _ = one_pass(field, zz, noise2, True, capi.loss_desc(capi.LOSS_SOLAR, n_mean, lam=float(args.sc_lambda)))
                loss_dict[f"{name}_sc_term2"], loss_dict[f"{name}_sc_term3"] = t2[2], t2[3]
        return outs

    outs = level("coarse", z)
    if n_imp > 0:
        u = draw("u", R, n_imp).contiguous()
        level("fine", capi.importance_depths(z, outs["weights"], u))
    if backward and g_t_total is not None:
        if emb.weight.grad is None:
            emb.weight.grad = torch.zeros_like(emb.weight)
        emb.weight.grad.index_add_(0, ts.reshape(-1), g_t_total)
    return loss_dict, results


def batched_inference(models, rays, ts, args):
    """eval_satnerf.py:46-66 — the no-grad chunk loop.  The fused path needs no activation chunking, but the
    loop is kept so memory stays bounded by args.chunk rays exactly like the reference."""
    with torch.no_grad():
        chunks = [render_rays(models, args, rays[i:i + args.chunk], None if ts is None else ts[i:i + args.chunk])
                  for i in range(0, rays.shape[0], args.chunk)]
    return {k: torch.cat([c[k] for c in chunks], 0) for k in chunks[0]}
