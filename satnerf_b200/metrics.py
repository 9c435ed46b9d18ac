"""Losses of the training step — the classes of the reference's metrics.py:8-103 with the same names, constructor arguments,
call signature `(inputs, targets) -> (loss, loss_dict)` and loss_dict keys; ssim (kornia, an evaluation metric outside the
render hot path) is left out.

The arithmetic is not done with torch ops on the result dict: every term is evaluated per ray inside libsatnerf_b200.so
(`snb_loss_forward`, csrc/composite.cu) and its gradient w.r.t. the dict tensors comes from `snb_loss_backward`, so these
classes are thin autograd bindings.  The training harness (satnerf_b200/train.py) goes one step further and seeds the
render backward with the loss gradient directly (`snb_render_grads.loss`): no (R,S,.) gradient tensor exists at all.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import capi

_KIND_TERMS = {capi.LOSS_COLOR_MSE: ("color",), capi.LOSS_COLOR_BETA: ("color", "logbeta"), capi.LOSS_DEPTH: ("ds",),
               capi.LOSS_SOLAR: ("sc_term2", "sc_term3")}
_TERM_SLOT = {"color": 0, "logbeta": 1, "ds": 0, "sc_term2": 2, "sc_term3": 3}


class _LossTerms(torch.autograd.Function):
    """(4,) loss terms of one level; differentiable w.r.t. rgb / depth / weights / beta / sun as the reference's losses are
    (weights and transparency of the solar-correction pass are detached there, metrics.py:30-31)."""

    @staticmethod
    def forward(ctx, kind, lam, n_mean, target, target_w, rgb, depth, weights, beta, sun, transparency):
        ref = next(t for t in (rgb, depth, sun) if t is not None)
        R = ref.shape[0]
        S = weights.shape[1] if weights is not None else 1
        f32 = lambda t: None if t is None else t.detach().to(torch.float32).contiguous()
        io = dict(rgb=f32(rgb), depth=f32(depth), weights=f32(weights), beta=f32(beta), sun=f32(sun), transparency=f32(transparency))
        pd = capi.PassDesc(R, S, 0, 0, 0, 0.0, 0, 0, 0.0)
        tgt, tw = f32(target), f32(target_w)
        terms = capi.loss_forward(pd, io, capi.loss_desc(kind, n_mean, tgt, tw, None, lam))
        ctx.pd, ctx.io, ctx.kind, ctx.lam, ctx.n_mean, ctx.tgt, ctx.tw = pd, io, kind, lam, n_mean, tgt, tw
        ctx.shapes = {k: (None if v is None else v.shape) for k, v in dict(rgb=rgb, depth=depth, weights=weights, beta=beta, sun=sun).items()}
        return terms

    @staticmethod
    def backward(ctx, g_terms):
        want = {"rgb": ctx.kind in (capi.LOSS_COLOR_MSE, capi.LOSS_COLOR_BETA), "depth": ctx.kind == capi.LOSS_DEPTH,
                "weights": ctx.kind == capi.LOSS_COLOR_BETA, "beta": ctx.kind == capi.LOSS_COLOR_BETA, "sun": ctx.kind == capi.LOSS_SOLAR}
        g = capi.loss_backward(ctx.pd, ctx.io, capi.loss_desc(ctx.kind, ctx.n_mean, ctx.tgt, ctx.tw, g_terms.to(torch.float32).contiguous(), ctx.lam), want)
        view = lambda k: None if (g[k] is None or ctx.shapes[k] is None) else g[k].view(ctx.shapes[k])
        return (None, None, None, None, None, view("rgb"), view("depth"), view("weights"), view("beta"), view("sun"), None)


def _terms(kind, inputs: Dict[str, torch.Tensor], typ: str, target=None, target_w=None, lam: float = 0.0, n_mean: Optional[int] = None) -> Dict[str, torch.Tensor]:
    sc = kind == capi.LOSS_SOLAR
    get = lambda k: inputs.get(f"{k}_sc_{typ}" if sc else f"{k}_{typ}")
    rgb, depth, weights = (None if sc else inputs.get(f"rgb_{typ}")), (None if sc else inputs.get(f"depth_{typ}")), get("weights")
    beta = inputs.get("beta_coarse") if kind == capi.LOSS_COLOR_BETA else None            # metrics.py:22 reads beta_coarse for both levels
    sun = get("sun").squeeze(-1) if sc else None
    trans = get("transparency") if sc else None
    if beta is not None:
        beta = beta.squeeze(-1)
        if beta.shape != weights.shape:
            raise RuntimeError(f"The size of tensor a ({weights.shape[1]}) must match the size of tensor b ({beta.shape[1]}) at non-singleton "
                               "dimension 1")          # the reference's own failure for sat-nerf coarse+fine
    R = (sun if sc else (rgb if rgb is not None else depth)).shape[0]
    t = _LossTerms.apply(kind, float(lam), int(n_mean or R), target, target_w, rgb, depth, weights, beta, sun, trans)
    return {f"{typ}_{name}": t[_TERM_SLOT[name]] for name in _KIND_TERMS[kind]}


class NerfLoss(torch.nn.Module):                                               # metrics.py:8-19
    def forward(self, inputs, targets):
        d = {}
        for typ in ("coarse", "fine"):
            if f"rgb_{typ}" in inputs:
                d.update(_terms(capi.LOSS_COLOR_MSE, inputs, typ, targets))
        return sum(d.values()), d


class SNerfLoss(torch.nn.Module):                                              # metrics.py:36-55
    def __init__(self, lambda_sc=0.05):
        super().__init__()
        self.lambda_sc = lambda_sc

    color_kind = capi.LOSS_COLOR_MSE

    def forward(self, inputs, targets):
        d = {}
        for typ in ("coarse", "fine"):
            if f"rgb_{typ}" not in inputs:
                continue
            d.update(_terms(self.color_kind, inputs, typ, targets))
            if self.lambda_sc > 0:                                             # solar_correction, metrics.py:27-34
                d.update(_terms(capi.LOSS_SOLAR, inputs, typ, lam=self.lambda_sc))
        return sum(d.values()), d


class SatNerfLoss(SNerfLoss):                                                  # metrics.py:57-73 (uncertainty_aware_loss :21-25)
    def __init__(self, lambda_sc=0.0):
        super().__init__(lambda_sc)

    color_kind = capi.LOSS_COLOR_BETA


class DepthLoss(torch.nn.Module):                                              # metrics.py:75-92
    def __init__(self, lambda_ds=1.0):
        super().__init__()
        self.lambda_ds = lambda_ds / 3.0

    def forward(self, inputs, targets, weights=1.0):
        d = {}
        tw = weights if torch.is_tensor(weights) else None
        scale = 1.0 if tw is not None else float(weights)
        for typ in ("coarse", "fine"):
            if f"depth_{typ}" in inputs:
                d.update(_terms(capi.LOSS_DEPTH, inputs, typ, targets, tw, lam=3.0 * self.lambda_ds * scale))
        return sum(d.values()), d


def load_loss(args):                                                           # metrics.py:94-103
    if args.model == "nerf":
        return NerfLoss()
    if args.model == "s-nerf":
        return SNerfLoss(lambda_sc=args.sc_lambda)
    if args.model == "sat-nerf":
        return SatNerfLoss(lambda_sc=args.sc_lambda)
    raise ValueError(f"model {args.model} is not valid")


def mse(image_pred, image_gt, valid_mask=None, reduction="mean"):              # metrics.py:105-111
    err = torch.square(image_pred - image_gt)
    err = err if valid_mask is None else err[valid_mask]
    return err.mean() if reduction == "mean" else err


def psnr(image_pred, image_gt, valid_mask=None, reduction="mean"):             # metrics.py:113-114
    return -10.0 * torch.log10(mse(image_pred, image_gt, valid_mask, reduction))
