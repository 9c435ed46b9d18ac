"""Host-side mirror of the reference's `models` package for the render hot path.

Same public names, constructor arguments, parameter names / shapes / init as the reference so that
released checkpoints load by name and `load_model(args)` is a drop-in
(models/__init__.py:6-15, models/satnerf.py:81-153, models/snerf.py:78-146, models/nerf.py:9-33,:135-182).
The modules own fp32 master parameters only; they contain no eager math.  All arithmetic runs in
libsatnerf_b200.so (hand-written sm_100a CUDA); calling `forward` on CPU tensors raises.

Parameters of a field are kept as views into ONE flat fp32 buffer (`flat_params`) in state_dict
order — the layout the C ABI consumes directly (snb_param_layout) and the buffer the multi-GPU path
all-reduces in a single NCCL call.
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
from torch import nn

from . import capi


class Siren(nn.Module):
    """Marker for the sin(w0*x) activation (models/nerf.py:23-33); evaluated inside the CUDA kernels."""

    def __init__(self, w0: float = 1.0):
        super().__init__()
        self.w0 = w0

    def extra_repr(self):
        return f"w0={self.w0}"

    def forward(self, x):
        raise RuntimeError("satnerf_b200 activations are fused into the CUDA kernels; call the field or render_rays")


def _uniform_(w: torch.Tensor, bound: float):
    with torch.no_grad():
        w.uniform_(-bound, bound)


def _siren_init(seq: nn.Sequential):
    """sine_init on every Linear, then first_layer_sine_init on the first (models/nerf.py:9-21,
    applied as in models/satnerf.py:145-149) — same order of generator draws as the reference."""
    for m in seq:
        if isinstance(m, nn.Linear):
            _uniform_(m.weight, math.sqrt(6 / m.weight.size(-1)))
    _uniform_(seq[0].weight, 1 / seq[0].weight.size(-1))


class _Field(nn.Module):
    """Shared machinery: trunk/head construction, flat parameter storage, forward dispatch."""

    variant = None            # 'nerf' | 's-nerf' | 'sat-nerf'
    number_of_outputs = 0

    def _build(self, layers, feat, mapping, mapping_sizes, skips, siren, t_dims):
        self.layers, self.skips, self.feat = layers, list(skips), feat
        self.siren = siren
        self.rgb_padding = 0.001
        self.mapping_sizes = list(mapping_sizes)
        self.use_mapping = bool(mapping)
        in_xyz = 2 * mapping_sizes[0] * self.input_sizes[0] if mapping else self.input_sizes[0]
        in_dir = 2 * mapping_sizes[1] * self.input_sizes[1] if mapping else self.input_sizes[1]
        act = (lambda w0=1.0: Siren(w0)) if siren else (lambda w0=1.0: nn.ReLU())

        trunk: List[nn.Module] = [nn.Linear(in_xyz, feat), act(30.0)]
        for i in range(1, layers):
            trunk += [nn.Linear(feat + in_xyz if i in self.skips else feat, feat), act()]
        self.fc_net = nn.Sequential(*trunk)
        self.sigma_from_xyz = nn.Sequential(nn.Linear(feat, 1), nn.Softplus())
        self.feats_from_xyz = nn.Linear(feat, feat)
        self.rgb_from_xyzdir = nn.Sequential(nn.Linear(feat + in_dir, feat // 2), act(), nn.Linear(feat // 2, 3), nn.Sigmoid())
        if self.variant != "nerf":
            sun: List[nn.Module] = [nn.Linear(feat + 3, feat // 2), act()]
            for _ in range(2):
                sun += [nn.Linear(feat // 2, feat // 2), act()]
            sun += [nn.Linear(feat // 2, 1), nn.Sigmoid()]
            self.sun_v_net = nn.Sequential(*sun)
            self.sky_color = nn.Sequential(nn.Linear(3, feat // 2), nn.ReLU(), nn.Linear(feat // 2, 3), nn.Sigmoid())
        if siren:
            _siren_init(self.fc_net)
            if self.variant != "nerf":
                _siren_init(self.sun_v_net)
        if self.variant == "sat-nerf":
            self.t_embedding_dims = t_dims
            self.beta_from_xyz = nn.Sequential(nn.Linear(t_dims + feat, feat // 2), act(), nn.Linear(feat // 2, 1), nn.Softplus())
        if self.use_mapping != (self.variant == "nerf") or siren != (self.variant != "nerf"):
            raise NotImplementedError("satnerf_b200 builds the configurations load_model() constructs: "
                                      "nerf = mapping+ReLU, s-nerf / sat-nerf = identity mapping + SIREN")
        self._desc = capi.field_desc(self.variant, layers, feat, self.skips, t_dims, self.mapping_sizes, self.use_mapping)
        self._flat: Optional[torch.Tensor] = None
        self._flat_grad: Optional[torch.Tensor] = None
        self._packed_epoch = 0

    def invalidate_packed(self):
        """Drops the cached fp16 weight tiles of the inference path.  In-place updates through the parameters (optimizer steps,
        load_state_dict) are detected by their version counters; call this after writing through `.data` / `flat_params()`
        (EMA swaps, dist.broadcast(p.data), manual copies), which autograd cannot see."""
        self._packed_epoch += 1

    # ---- flat parameter storage -------------------------------------------------------------
    @property
    def desc(self) -> capi.FieldDesc:
        return self._desc

    def ordered_params(self) -> List[nn.Parameter]:
        return list(self.parameters())         # registration order == state_dict order == snb_param_layout order

    def flat_params(self) -> torch.Tensor:
        """Returns the flat fp32 parameter buffer, (re)building it when `.to()/.cuda()` replaced the storages."""
        ps = self.ordered_params()
        flat = self._flat
        ok = flat is not None and flat.device == ps[0].device
        if ok:
            off, base = 0, flat.data_ptr()
            for p in ps:
                if p.data_ptr() != base + 4 * off:
                    ok = False
                    break
                off += p.numel()
        if not ok:
            total = sum(p.numel() for p in ps)
            flat = torch.empty(total, dtype=torch.float32, device=ps[0].device)
            off = 0
            for p in ps:
                n = p.numel()
                flat[off:off + n].copy_(p.data.reshape(-1))
                p.data = flat[off:off + n].view(p.shape)
                off += n
            self._flat, self._flat_grad = flat, None
            self._packed_epoch += 1
        return flat

    def adopt_flat_storage(self, flat_new: torch.Tensor, grad_new: torch.Tensor) -> None:
        """Moves the parameters (values kept) and their gradients into caller-provided flat buffers of the same size -- e.g.
        symmetric memory that the other ranks of a data-parallel job can address (dist.make_symmetric)."""
        old = self.flat_params()
        if flat_new.numel() != old.numel() or grad_new.numel() != old.numel():
            raise ValueError("adopt_flat_storage: buffers must hold exactly the field's parameters")
        with torch.no_grad():
            flat_new.copy_(old)
            grad_new.zero_()
            off = 0
            for p in self.ordered_params():
                n = p.numel()
                p.data = flat_new[off:off + n].view(p.shape)
                p.grad = grad_new[off:off + n].view(p.shape)
                off += n
        self._flat, self._flat_grad = flat_new, grad_new
        object.__setattr__(self, "_flat_param", None)
        self._packed_epoch += 1

    def flat_parameter(self) -> nn.Parameter:
        """The flat buffer as ONE nn.Parameter (shares storage and version counter with `flat_params()`; `.grad` aliases
        `flat_grads()`): what the training harness hands to Adam, so the update is one launch over one tensor.  The module's own
        parameters are views of the same storage and see every update."""
        flat = self.flat_params()
        fp = getattr(self, "_flat_param", None)
        if fp is None or fp.data_ptr() != flat.data_ptr() or fp.device != flat.device:
            fp = nn.Parameter(flat, requires_grad=True)
            object.__setattr__(self, "_flat_param", fp)      # not registered: state_dict / parameters() stay the reference's
        fp.grad = self.flat_grads(zero=False)
        return fp

    def flat_grads(self, zero: bool = True) -> torch.Tensor:
        """One flat gradient buffer whose slices are the parameters' `.grad` (a single NCCL all-reduce target)."""
        flat = self.flat_params()
        g = self._flat_grad
        ps = self.ordered_params()
        bound = g is not None and g.device == flat.device and all(p.grad is not None for p in ps)
        if bound:
            off, base = 0, g.data_ptr()
            for p in ps:
                if p.grad.data_ptr() != base + 4 * off:
                    bound = False
                    break
                off += p.numel()
        if not bound:
            g = torch.zeros_like(flat)
            off = 0
            for p in ps:
                n = p.numel()
                if p.grad is not None:
                    g[off:off + n].copy_(p.grad.reshape(-1))
                p.grad = g[off:off + n].view(p.shape)
                off += n
            self._flat_grad = g
        elif zero:
            g.zero_()
        fp = getattr(self, "_flat_param", None)
        if fp is not None and (fp.grad is None or fp.grad.data_ptr() != g.data_ptr()):
            fp.grad = g
        return g

    # ---- <Field>.forward: per-point evaluation ------------------------------------------------
    def _points(self, input_xyz, aux, input_t, sigma_only):
        if not input_xyz.is_cuda:
            raise RuntimeError("satnerf_b200 fields run on CUDA tensors only (no CPU fallback)")
        f32 = lambda t: None if t is None else t.detach().to(torch.float32).contiguous()
        if not sigma_only:
            if aux is None:
                raise TypeError("direction input is required" if self.variant == "nerf" else "input_sun_dir is required")
            if self.variant == "sat-nerf" and input_t is None:
                raise TypeError("sat-nerf needs input_t (torch.cat with None in the reference, models/satnerf.py:204)")
        xyz, aux_, t_ = f32(input_xyz), f32(aux), f32(input_t)
        if sigma_only and aux_ is None:
            aux_ = torch.zeros_like(xyz)                     # (unused by the sigma head; the backward's buffers want a direction)
        needs_grad = torch.is_grad_enabled() and (any(p.requires_grad for p in self.parameters()) or (input_t is not None and input_t.requires_grad))
        prec = self._points_precision()
        if not needs_grad:
            return capi.field_forward(self._desc, self.flat_params(), xyz, aux_, t_, sigma_only, self.number_of_outputs, prec)
        if sigma_only and self.variant == "sat-nerf" and t_ is None:
            t_ = torch.zeros(xyz.shape[0], self._desc.t_dims, device=xyz.device)
        t_in = input_t if (input_t is not None and not sigma_only) else None
        return _PointsFn.apply(self, xyz, aux_, t_, sigma_only, t_in, prec, *self.ordered_params())

    def _points_precision(self) -> int:
        """`field.points_precision`: 'tcx3' (default on sm_100: contractions on the tensor cores, fp16 hi+lo operands, fp32-level
        results) | 'fp32' (FFMA).  The gradients are the fp32 CUDA-core chain either way."""
        p = getattr(self, "points_precision", None)
        if p is None:
            p = "tcx3" if capi.device_supports_tc() else "fp32"
        if p not in ("tcx3", "fp32"):
            raise ValueError(f"points_precision {p!r} is not valid (tcx3 | fp32)")
        return capi.FP16X3_TC if p == "tcx3" else capi.FP32_SIMT


class _PointsFn(torch.autograd.Function):
    """<Field>.forward on B points, differentiable w.r.t. the field parameters and input_t (models/satnerf.py:156-208 under
    autograd).  fp32 CUDA-core path: the backward recomputes the forward chunk by chunk (snb_field_backward).  No gradient flows
    to input_xyz / the directions -- the reference's training never asks for one (rays are data)."""

    @staticmethod
    def forward(ctx, field, xyz, aux, t_emb, sigma_only, t_in, prec, *params):
        out = capi.field_forward(field._desc, field.flat_params(), xyz, aux, t_emb, sigma_only, field.number_of_outputs, prec)
        ctx.field, ctx.sigma_only, ctx.has_t = field, sigma_only, t_in is not None
        ctx.save_for_backward(xyz, aux, out, *([t_emb] if t_emb is not None else []))
        return out

    @staticmethod
    def backward(ctx, d_out):
        field = ctx.field
        xyz, aux, out, *rest = ctx.saved_tensors
        t_emb = rest[0] if rest else None
        flat = field.flat_params()
        g_flat = torch.zeros_like(flat)
        g_t = capi.field_backward(field._desc, flat, xyz, aux, t_emb, out, d_out.to(torch.float32).contiguous(), g_flat, ctx.sigma_only)
        gp, off = [], 0
        for p in field.ordered_params():
            n = p.numel()
            gp.append(g_flat[off:off + n].view(p.shape))
            off += n
        return (None, None, None, None, None, g_t if ctx.has_t else None, None, *gp)


class NeRF(_Field):
    """models/nerf.py:135-227.  forward(input_xyz, input_dir, sigma_only) -> (B,4) = [rgb3, sigma]."""
    variant = "nerf"
    number_of_outputs = 4

    def __init__(self, layers=8, feat=256, mapping=True, mapping_sizes=[10, 4], skips=[4], siren=False):
        super().__init__()
        self.input_sizes = [3, 3]
        self._build(layers, feat, mapping, mapping_sizes, skips, siren, 0)

    def forward(self, input_xyz, input_dir=None, sigma_only=False):
        return self._points(input_xyz, input_dir, None, sigma_only)


class ShadowNeRF(_Field):
    """models/snerf.py:78-196.  forward(...) -> (B,8) = [rgb3, sigma, sun, sky3]."""
    variant = "s-nerf"
    number_of_outputs = 8

    def __init__(self, layers=8, feat=256, mapping=False, mapping_sizes=[10, 4], skips=[4], siren=True):
        super().__init__()
        self.input_sizes = [3, 0]
        self._build(layers, feat, mapping, mapping_sizes, skips, siren, 0)

    def forward(self, input_xyz, input_dir=None, input_sun_dir=None, sigma_only=False):
        return self._points(input_xyz, input_sun_dir, None, sigma_only)


class SatNeRF(_Field):
    """models/satnerf.py:81-208.  forward(...) -> (B,9) = [rgb3, sigma, sun, sky3, beta]."""
    variant = "sat-nerf"
    number_of_outputs = 9

    def __init__(self, layers=8, feat=256, mapping=False, mapping_sizes=[10, 4], skips=[4], siren=True, t_embedding_dims=16):
        super().__init__()
        self.input_sizes = [3, 0]
        self._build(layers, feat, mapping, mapping_sizes, skips, siren, t_embedding_dims)

    def forward(self, input_xyz, input_dir=None, input_sun_dir=None, input_t=None, sigma_only=False):
        return self._points(input_xyz, input_sun_dir, input_t, sigma_only)


def load_model(args):
    """models/__init__.py:6-15."""
    if args.model == "nerf":
        return NeRF(layers=args.fc_layers, feat=args.fc_units)
    if args.model == "s-nerf":
        return ShadowNeRF(layers=args.fc_layers, feat=args.fc_units)
    if args.model == "sat-nerf":
        return SatNeRF(layers=args.fc_layers, feat=args.fc_units, t_embedding_dims=args.t_embbeding_tau)
    raise ValueError(f'model {args.model} is not valid')
