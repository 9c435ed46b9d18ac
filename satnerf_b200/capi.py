"""ctypes binding of libsatnerf_b200.so (C ABI declared in include/satnerf_b200.h).

PyTorch is used for device memory and streams only: every call passes raw device pointers
(`tensor.data_ptr()`) and the current CUDA stream handle.  There is no CPU or eager fallback:
a missing library, a CPU tensor or a non-zero return code raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional

import torch

_LIB_PATH = os.environ.get("SNB_LIBRARY_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libsatnerf_b200.so")   # (override: developer A/B builds)

NERF, SNERF, SATNERF = 0, 1, 2
FP32_SIMT, FP16_TC, FP16X3_TC = 0, 1, 2
VARIANTS = {"nerf": NERF, "s-nerf": SNERF, "sat-nerf": SATNERF}

c_float_p = C.POINTER(C.c_float)


class FieldDesc(C.Structure):
    _fields_ = [("variant", C.c_int32), ("n_layers", C.c_int32), ("width", C.c_int32), ("skip_layer", C.c_int32),
                ("t_dims", C.c_int32), ("pe_xyz", C.c_int32), ("pe_dir", C.c_int32)]


class PassDesc(C.Structure):
    _fields_ = [("n_rays", C.c_int32), ("n_samples", C.c_int32), ("ray_cols", C.c_int32),
                ("march_along_sun", C.c_int32), ("precision", C.c_int32), ("noise_std", C.c_float), ("weights_packed", C.c_int32),
                ("flags", C.c_int32), ("t_min", C.c_float)]


PASS_SINGLE_CTA, PASS_NO_BETA, PASS_SIGMA_ONLY = 1, 2, 4
LOSS_COLOR_MSE, LOSS_COLOR_BETA, LOSS_DEPTH, LOSS_SOLAR = 1, 2, 3, 4


class LossDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_rays_mean", C.c_int32), ("lambda_", C.c_float), ("beta_min", C.c_float),
                ("target", C.c_void_p), ("target_weight", C.c_void_p), ("g_terms", C.c_void_p)]


_IO_FIELDS = ["params", "rays", "z_vals", "t_emb", "noise", "xyz", "aux_dir", "rgb", "depth", "weights", "transparency",
              "albedo", "sun", "sky", "beta", "sigma", "nerf_rgb", "stash", "aux_sums"]
_GRAD_FIELDS = ["g_rgb", "g_depth", "g_weights", "g_transparency", "g_albedo", "g_sun", "g_sky", "g_beta",
                "g_params", "g_t_emb"]       # + `loss` (pointer to LossDesc)


class RpcModel(C.Structure):
    """snb_rpc_model: the fields of rpcm.RPCModel (projection direction)."""
    _fields_ = [(n, C.c_double * 20) for n in ("row_num", "row_den", "col_num", "col_den")] + \
               [(n, C.c_double) for n in ("row_offset", "row_scale", "col_offset", "col_scale", "lat_offset", "lat_scale",
                                          "lon_offset", "lon_scale", "alt_offset", "alt_scale")]


class RenderIO(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _IO_FIELDS]


class RenderGrads(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in _GRAD_FIELDS] + [("loss", C.POINTER(LossDesc))]


_SIGNATURES = {
    "snb_abi_version": (C.c_int, []),
    "snb_last_error": (C.c_char_p, []),
    "snb_device_supports_tc": (C.c_int, []),
    "snb_launch_count": (C.c_int64, [C.c_int]),
    "snb_param_layout": (C.c_int, [C.POINTER(FieldDesc), C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                   C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int]),
    "snb_param_count": (C.c_int64, [C.POINTER(FieldDesc)]),
    "snb_stratified_depths": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "snb_importance_depths": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "snb_searchsorted_right": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "snb_render_workspace": (C.c_int, [C.POINTER(FieldDesc), C.POINTER(PassDesc), C.c_int, C.POINTER(C.c_size_t)]),
    "snb_render_stash_bytes": (C.c_int, [C.POINTER(FieldDesc), C.POINTER(PassDesc), C.POINTER(C.c_size_t)]),
    "snb_render_forward": (C.c_int, [C.POINTER(FieldDesc), C.POINTER(PassDesc), C.POINTER(RenderIO), C.c_void_p, C.c_size_t, C.c_void_p]),
    "snb_render_backward": (C.c_int, [C.POINTER(FieldDesc), C.POINTER(PassDesc), C.POINTER(RenderIO), C.POINTER(RenderGrads),
                                      C.c_void_p, C.c_size_t, C.c_void_p]),
    "snb_loss_forward": (C.c_int, [C.POINTER(PassDesc), C.POINTER(RenderIO), C.POINTER(LossDesc), C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "snb_loss_backward": (C.c_int, [C.POINTER(PassDesc), C.POINTER(RenderIO), C.POINTER(LossDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "snb_field_backward_workspace": (C.c_int, [C.POINTER(FieldDesc), C.c_int, C.POINTER(C.c_size_t)]),
    "snb_field_backward": (C.c_int, [C.POINTER(FieldDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
    "snb_rpc_rays": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_double, C.c_double, C.c_void_p, C.c_double,
                               C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "snb_dsm_points": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "snb_dsm_rasterize": (C.c_int, [C.c_void_p, C.c_longlong, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                    C.c_void_p, C.c_size_t, C.c_void_p]),
    "snb_gather_rows": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p]),
    "snb_adam_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_double, C.c_double, C.c_double, C.c_double,
                                C.c_double, C.c_int, C.c_void_p]),
    "snb_adam_step_sharded": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_double, C.c_double,
                                        C.c_double, C.c_double, C.c_double, C.c_int, C.c_void_p]),
    "snb_adam_step_sharded_multi": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_void_p]),
    "snb_field_workspace": (C.c_int, [C.POINTER(FieldDesc), C.c_int, C.POINTER(C.c_size_t)]),
    "snb_field_forward": (C.c_int, [C.POINTER(FieldDesc), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
}

_lib = None


def lib():
    """Loads the shared library once; raises (never falls back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(
                f"{_LIB_PATH} is missing: build it with `python -m satnerf_b200.build` "
                "(satnerf_b200 has no CPU or eager fallback)")
        handle = C.CDLL(_LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)          # AttributeError if the symbol is not exported
            fn.restype, fn.argtypes = res, args
        if handle.snb_abi_version() != 5:
            raise RuntimeError("libsatnerf_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def _check(code: int, what: str):
    if code != 0:
        raise RuntimeError(f"{what} failed ({code}): {lib().snb_last_error().decode()}")


def _ptr(t: Optional[torch.Tensor], dtype=torch.float32):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("satnerf_b200 runs on CUDA tensors only (no CPU fallback)")
    if t.dtype != dtype or not t.is_contiguous():
        raise RuntimeError(f"expected contiguous {dtype} tensor, got {t.dtype} contiguous={t.is_contiguous()}")
    return C.c_void_p(t.data_ptr())


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def field_desc(variant: str, n_layers: int, width: int, skips, t_dims: int = 0, mapping_sizes=(10, 4), mapping=None) -> FieldDesc:
    if variant not in VARIANTS:
        raise ValueError(f"model {variant} is not valid")      # same error as models/__init__.py:14
    skips = list(skips)
    if len(skips) > 1:
        raise ValueError("only a single skip connection is supported")
    if mapping is None:
        mapping = variant == "nerf"
    return FieldDesc(VARIANTS[variant], n_layers, width, skips[0] if skips else -1, t_dims if variant == "sat-nerf" else 0,
                     mapping_sizes[0] if mapping else 0, mapping_sizes[1] if mapping else 0)


def param_layout(desc: FieldDesc):
    """[(w_off, b_off, n_out, n_in)] in state_dict order — host arithmetic only, usable without a GPU."""
    cap = 64
    w, b = (C.c_int64 * cap)(), (C.c_int64 * cap)()
    no, ni = (C.c_int32 * cap)(), (C.c_int32 * cap)()
    n = lib().snb_param_layout(C.byref(desc), w, b, no, ni, cap)
    if n < 0:
        _check(n, "snb_param_layout")
    return [(w[i], b[i], no[i], ni[i]) for i in range(n)]


def param_count(desc: FieldDesc) -> int:
    n = lib().snb_param_count(C.byref(desc))
    if n < 0:
        _check(int(n), "snb_param_count")
    return int(n)


def launch_count(reset: bool = False) -> int:
    return int(lib().snb_launch_count(int(reset)))


def device_supports_tc() -> bool:
    return bool(lib().snb_device_supports_tc())


# ---- workspace cache (one growing buffer per device; calls are ordered on the current stream) ----
_workspaces: Dict[int, torch.Tensor] = {}


def _workspace(device, nbytes: int) -> torch.Tensor:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    ws = _workspaces.get(idx)
    if ws is None or ws.numel() < nbytes:
        _workspaces[idx] = ws = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=device)
    return ws


def stratified_depths(rays, steps, u):
    R, S = u.shape
    z = torch.empty_like(u)
    with torch.cuda.device(rays.device):
        _check(lib().snb_stratified_depths(_ptr(rays), rays.shape[1], _ptr(steps), _ptr(u), _ptr(z), R, S, _stream(rays.device)),
               "snb_stratified_depths")
    return z


def importance_depths(z_coarse, weights_coarse, u, debug=False):
    R, S = z_coarse.shape
    N = u.shape[1]
    z_out = torch.empty(R, S + N, device=u.device, dtype=torch.float32)
    inds = torch.empty(R, N, device=u.device, dtype=torch.int64) if debug else None
    z_new = torch.empty(R, N, device=u.device, dtype=torch.float32) if debug else None
    cdf = torch.empty(R, S - 1, device=u.device, dtype=torch.float32) if debug else None
    with torch.cuda.device(u.device):
        _check(lib().snb_importance_depths(_ptr(z_coarse), _ptr(weights_coarse), _ptr(u), _ptr(z_out), _ptr(inds, torch.int64),
                                           _ptr(z_new), _ptr(cdf), R, S, N, _stream(u.device)), "snb_importance_depths")
    return (z_out, inds, z_new, cdf) if debug else z_out


def searchsorted_right(cdf, u):
    R, n = cdf.shape
    inds = torch.empty(u.shape, device=u.device, dtype=torch.int64)
    with torch.cuda.device(u.device):
        _check(lib().snb_searchsorted_right(_ptr(cdf), _ptr(u), _ptr(inds, torch.int64), R, n, u.shape[1], _stream(u.device)),
               "snb_searchsorted_right")
    return inds


def render_stash_bytes(desc: FieldDesc, pd: PassDesc) -> int:
    n = C.c_size_t(0)
    _check(lib().snb_render_stash_bytes(C.byref(desc), C.byref(pd), C.byref(n)), "snb_render_stash_bytes")
    return int(n.value)


def _fill(struct, names, tensors: Dict[str, Optional[torch.Tensor]]):
    for n in names:
        t = tensors.get(n)
        setattr(struct, n, _ptr(t, torch.uint8) if n == "stash" else _ptr(t))
    return struct


def render_forward(desc: FieldDesc, pd: PassDesc, tensors: Dict[str, Optional[torch.Tensor]], workspace: Optional[torch.Tensor] = None):
    """`workspace`: a caller-owned uint8 buffer of render_workspace_bytes() bytes (needed for pd.weights_packed, which relies
    on the buffer's contents surviving between calls); default: the shared per-device scratch."""
    dev = tensors["rays"].device if tensors.get("rays") is not None else tensors["xyz"].device
    io = _fill(RenderIO(), _IO_FIELDS, tensors)
    n = C.c_size_t(0)
    with torch.cuda.device(dev):
        _check(lib().snb_render_workspace(C.byref(desc), C.byref(pd), 0, C.byref(n)), "snb_render_workspace")
        ws = workspace if workspace is not None else _workspace(dev, n.value)
        if ws.numel() < n.value:
            raise RuntimeError(f"render_forward: workspace of {ws.numel()} bytes given, {n.value} needed")
        _check(lib().snb_render_forward(C.byref(desc), C.byref(pd), C.byref(io), C.c_void_p(ws.data_ptr()), ws.numel(), _stream(dev)),
               "snb_render_forward")


def render_workspace_bytes(desc: FieldDesc, pd: PassDesc, backward: bool = False) -> int:
    n = C.c_size_t(0)
    _check(lib().snb_render_workspace(C.byref(desc), C.byref(pd), int(backward), C.byref(n)), "snb_render_workspace")
    return int(n.value)


def loss_desc(kind: int, n_rays_mean: int, target=None, target_weight=None, g_terms=None, lam: float = 0.0, beta_min: float = 0.05) -> LossDesc:
    """snb_loss_desc over device tensors (kept alive by the caller for the duration of the call)."""
    return LossDesc(kind, int(n_rays_mean), float(lam), float(beta_min), _ptr(target), _ptr(target_weight), _ptr(g_terms))


def loss_forward(pd: PassDesc, tensors: Dict[str, Optional[torch.Tensor]], loss: LossDesc) -> torch.Tensor:
    """(4,) loss terms of one pass from its forward results (metrics.py:8-92 evaluated in the library)."""
    ref = next(t for t in tensors.values() if t is not None)
    dev = ref.device
    io = _fill(RenderIO(), _IO_FIELDS, tensors)
    terms = torch.empty(4, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        ws = _workspace(dev, pd.n_rays * 16 + 256)
        _check(lib().snb_loss_forward(C.byref(pd), C.byref(io), C.byref(loss), _ptr(terms), C.c_void_p(ws.data_ptr()), ws.numel(), _stream(dev)),
               "snb_loss_forward")
    return terms


def loss_backward(pd: PassDesc, tensors: Dict[str, Optional[torch.Tensor]], loss: LossDesc, want: Dict[str, bool]):
    """Gradients of the loss terms w.r.t. the result-dict tensors: {'rgb','depth','weights','beta','sun'} -> tensors (or None)."""
    ref = next(t for t in tensors.values() if t is not None)
    dev, R, S = ref.device, pd.n_rays, pd.n_samples
    io = _fill(RenderIO(), _IO_FIELDS, tensors)
    shapes = {"rgb": (R, 3), "depth": (R,), "weights": (R, S), "beta": (R, S), "sun": (R, S)}
    out = {k: (torch.empty(shapes[k], device=dev, dtype=torch.float32) if want.get(k) else None) for k in shapes}
    with torch.cuda.device(dev):
        _check(lib().snb_loss_backward(C.byref(pd), C.byref(io), C.byref(loss), _ptr(out["rgb"]), _ptr(out["depth"]), _ptr(out["weights"]),
                                       _ptr(out["beta"]), _ptr(out["sun"]), _stream(dev)), "snb_loss_backward")
    return out


def render_backward(desc: FieldDesc, pd: PassDesc, tensors: Dict[str, Optional[torch.Tensor]], grads: Dict[str, Optional[torch.Tensor]],
                    loss: Optional[LossDesc] = None):
    dev = tensors["params"].device
    io = _fill(RenderIO(), _IO_FIELDS, tensors)
    g = _fill(RenderGrads(), _GRAD_FIELDS, grads)
    if loss is not None:
        g.loss = C.pointer(loss)
    n = C.c_size_t(0)
    with torch.cuda.device(dev):
        _check(lib().snb_render_workspace(C.byref(desc), C.byref(pd), 1, C.byref(n)), "snb_render_workspace")
        ws = _workspace(dev, n.value)
        _check(lib().snb_render_backward(C.byref(desc), C.byref(pd), C.byref(io), C.byref(g), C.c_void_p(ws.data_ptr()), ws.numel(),
                                         _stream(dev)), "snb_render_backward")


def field_forward(desc: FieldDesc, params, xyz, aux_dir, t_emb, sigma_only: bool, n_channels: int, precision=FP32_SIMT):
    B = xyz.shape[0]
    out = torch.empty(B, 1 if sigma_only else n_channels, device=xyz.device, dtype=torch.float32)
    n = C.c_size_t(0)
    with torch.cuda.device(xyz.device):
        _check(lib().snb_field_workspace(C.byref(desc), B, C.byref(n)), "snb_field_workspace")
        ws = _workspace(xyz.device, n.value)
        _check(lib().snb_field_forward(C.byref(desc), _ptr(params), _ptr(xyz), _ptr(aux_dir), _ptr(t_emb), _ptr(out), B,
                                       int(sigma_only), precision, C.c_void_p(ws.data_ptr()), ws.numel(), _stream(xyz.device)),
               "snb_field_forward")
    return out


def adam_step(params, grads, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, weight_decay, step):
    """torch.optim.Adam update of one flat fp32 buffer, in place (one launch)."""
    for t in (params, grads, exp_avg, exp_avg_sq):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.numel() == params.numel()):
            raise ValueError("adam_step: contiguous fp32 CUDA buffers of one size expected")
    with torch.cuda.device(params.device):
        _check(lib().snb_adam_step(C.c_void_p(params.data_ptr()), C.c_void_p(grads.data_ptr()), C.c_void_p(exp_avg.data_ptr()),
                                   C.c_void_p(exp_avg_sq.data_ptr()), params.numel(), lr, beta1, beta2, eps, weight_decay, int(step),
                                   _stream(params.device)), "snb_adam_step")


def field_backward(desc: FieldDesc, params, xyz, aux_dir, t_emb, out, d_out, g_params, sigma_only: bool):
    """Accumulates d(loss)/d(params) of <Field>.forward into g_params (flat); returns d(loss)/d(input_t) (B,tau) or None."""
    B = xyz.shape[0]
    g_t = torch.empty_like(t_emb) if (t_emb is not None and not sigma_only) else None
    n = C.c_size_t(0)
    with torch.cuda.device(xyz.device):
        _check(lib().snb_field_backward_workspace(C.byref(desc), B, C.byref(n)), "snb_field_backward_workspace")
        ws = _workspace(xyz.device, n.value)
        _check(lib().snb_field_backward(C.byref(desc), _ptr(params), _ptr(xyz), _ptr(aux_dir), _ptr(t_emb), _ptr(out), _ptr(d_out), _ptr(g_params),
                                        _ptr(g_t), B, int(sigma_only), C.c_void_p(ws.data_ptr()), ws.numel(), _stream(xyz.device)), "snb_field_backward")
    return g_t


def adam_step_sharded(peer_param_ptrs, peer_grad_ptrs, rank, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, step, device):
    """One rank's part of the fused reduce-scatter + Adam + all-gather step (snb_adam_step_sharded); the caller issues the barriers."""
    world = len(peer_param_ptrs)
    pp = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in peer_param_ptrs])
    gp = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in peer_grad_ptrs])
    with torch.cuda.device(device):
        _check(lib().snb_adam_step_sharded(pp, gp, world, int(rank), C.c_void_p(exp_avg.data_ptr()), C.c_void_p(exp_avg_sq.data_ptr()), int(n),
                                           lr, beta1, beta2, eps, weight_decay, int(step), _stream(device)), "snb_adam_step_sharded")


class ShardedBuffer(C.Structure):
    """snb_sharded_buffer (include/satnerf_b200.h)."""
    _fields_ = [("peer_params", C.c_void_p), ("peer_grads", C.c_void_p), ("mc_params", C.c_void_p), ("mc_grads", C.c_void_p),
                ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p), ("n", C.c_longlong), ("step", C.c_int32)]


def adam_step_sharded_multi(buffers, rank, lr, beta1, beta2, eps, weight_decay, device):
    """One launch of the fused reduce + Adam + deliver step over several flat buffers (snb_adam_step_sharded_multi).
    buffers: list of dicts {param_ptrs, grad_ptrs, mc_params, mc_grads (0 = no multicast), exp_avg, exp_avg_sq, n, step}."""
    world = len(buffers[0]["param_ptrs"])
    arr = (ShardedBuffer * len(buffers))()
    keep = []
    for i, b in enumerate(buffers):
        pp = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in b["param_ptrs"]])
        gp = (C.c_void_p * world)(*[C.c_void_p(int(p)) for p in b["grad_ptrs"]])
        keep += [pp, gp]
        arr[i].peer_params = C.cast(pp, C.c_void_p); arr[i].peer_grads = C.cast(gp, C.c_void_p)
        arr[i].mc_params = C.c_void_p(int(b.get("mc_params") or 0) or None); arr[i].mc_grads = C.c_void_p(int(b.get("mc_grads") or 0) or None)
        arr[i].exp_avg = C.c_void_p(b["exp_avg"].data_ptr()); arr[i].exp_avg_sq = C.c_void_p(b["exp_avg_sq"].data_ptr())
        arr[i].n = int(b["n"]); arr[i].step = int(b["step"])
    with torch.cuda.device(device):
        _check(lib().snb_adam_step_sharded_multi(C.cast(arr, C.c_void_p), len(buffers), world, int(rank), lr, beta1, beta2, eps, weight_decay,
                                                 _stream(device)), "snb_adam_step_sharded_multi")


def gather_rows(tables: Dict[str, torch.Tensor], idx: torch.Tensor) -> Dict[str, torch.Tensor]:
    """{k: tables[k][idx]} for up to 4 contiguous row-major CUDA tables of 4- or 8-byte elements sharing the int64 index `idx`, in
    one launch (the batch assembly of data.DeviceRaySampler)."""
    keys = list(tables)
    if not 1 <= len(keys) <= 4:
        raise ValueError("gather_rows: 1..4 tables")
    dev = idx.device
    idx = idx.to(torch.int64).contiguous()
    n, n_src = idx.numel(), tables[keys[0]].shape[0]
    outs = {}
    for k in keys:
        t = tables[k]
        if not (t.is_cuda and t.is_contiguous() and t.shape[0] == n_src and t.element_size() in (4, 8)):
            raise ValueError(f"gather_rows: table {k!r} must be a contiguous CUDA tensor of 4- or 8-byte elements with {n_src} rows")
        outs[k] = torch.empty((n, *t.shape[1:]), dtype=t.dtype, device=dev)
    src = (C.c_void_p * len(keys))(*[C.c_void_p(tables[k].data_ptr()) for k in keys])
    dst = (C.c_void_p * len(keys))(*[C.c_void_p(outs[k].data_ptr()) for k in keys])
    rb = (C.c_int32 * len(keys))(*[int(tables[k][0].numel() * tables[k].element_size()) if n_src else 4 for k in keys])
    with torch.cuda.device(dev):
        _check(lib().snb_gather_rows(src, dst, rb, len(keys), C.c_void_p(idx.data_ptr()), n, n_src, _stream(dev)), "snb_gather_rows")
    return outs
