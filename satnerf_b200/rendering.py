"""Drop-in for the reference's `rendering.py`: `render_rays(models, args, rays, ts)` and `sample_pdf`,
plus the per-variant `inference()` (models/satnerf.py:4, snerf.py:4, nerf.py:71).

Python here only sequences passes and draws the random tensors in the reference's order
(rand_like -> randn [-> randn for the solar-correction pass] -> rand -> randn ...);
sampling, the MLP, compositing and their gradients run in libsatnerf_b200.so through the C ABI.

`args` is the reference's argparse Namespace.  Extra, optional attributes understood here:
  args.precision : 'tc'  (default on sm_100: fp16 operands / fp32 accumulate on tcgen05 tensor cores)
                   'fp32' (fp32 FFMA CUDA-core path, matches the reference to rounding level)
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
    if p not in ("tc", "fp32"):
        raise ValueError(f"precision {p!r} is not valid (tc | fp32)")
    return capi.FP16_TC if p == "tc" else capi.FP32_SIMT


def _out_shapes(variant: str, R: int, S: int):
    sh = {"rgb": (R, 3), "depth": (R,), "weights": (R, S), "transparency": (R, S)}
    if variant != "nerf":
        sh.update(albedo=(R, S, 3), sun=(R, S, 1), sky=(R, S, 3))
    if variant == "sat-nerf":
        sh["beta"] = (R, S, 1)
    return sh


class _Pass(torch.autograd.Function):
    """One inference() pass.  Differentiable w.r.t. the field parameters and the per-ray embedding."""

    @staticmethod
    def forward(ctx, field, cfg, rays, z, t_emb, noise, xyz, aux_dir, *params):
        variant = field.variant
        R, S = z.shape
        dev = z.device
        ctx.set_materialize_grads(False)       # outputs the loss does not touch arrive as None in backward (no zero-filled (R,S,.) tensors)
        pd = capi.PassDesc(R, S, rays.shape[1] if rays is not None else 0, int(cfg["sc"]), cfg["precision"], float(cfg["noise_std"]), 0)
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
        # Inference on the tensor-core path: the packed fp16 weight tiles are kept in a buffer owned by the field and
        # reused while the parameters are untouched (their autograd version counter and storage are unchanged) --
        # batched_inference / DSM extraction render many ray batches per weight set.
        ws = None
        if not cfg["train"] and cfg["precision"] == capi.FP16_TC:
            flat = tensors["params"]
            # (the parameters alias `flat` through `.data`, so each keeps its own version counter)
            key = (flat.data_ptr(), flat._version, tuple(p._version for p in params), str(dev))
            need = capi.render_workspace_bytes(field.desc, pd)
            cache = getattr(field, "_tc_packed", None)
            if cache is None or cache[0].device != dev or cache[0].numel() < need:
                cache = [torch.empty(max(need, 1), dtype=torch.uint8, device=dev), None]
                field._tc_packed = cache
            ws = cache[0]
            if need and cache[1] == key:
                pd.weights_packed = 1
            cache[1] = key
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


def _run_pass(field, args, rays, z, t_emb, noise, sc=False, xyz=None, aux_dir=None) -> Dict[str, torch.Tensor]:
    if not z.is_cuda:
        raise RuntimeError("satnerf_b200 renders CUDA tensors only (no CPU fallback); move rays and models to the GPU")
    # stash activations only when a backward can follow (Function.forward itself always runs with grad mode off)
    train = torch.is_grad_enabled() and (any(p.requires_grad for p in field.parameters()) or (t_emb is not None and t_emb.requires_grad))
    cfg = {"sc": sc, "precision": _precision(args), "noise_std": float(args.noise_std), "train": train}
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

    def level(name, zz):
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


def batched_inference(models, rays, ts, args):
    """eval_satnerf.py:46-66 — the no-grad chunk loop.  The fused path needs no activation chunking, but the
    loop is kept so memory stays bounded by args.chunk rays exactly like the reference."""
    with torch.no_grad():
        chunks = [render_rays(models, args, rays[i:i + args.chunk], None if ts is None else ts[i:i + args.chunk])
                  for i in range(0, rays.shape[0], args.chunk)]
    return {k: torch.cat([c[k] for c in chunks], 0) for k in chunks[0]}
