"""
TEST INFRASTRUCTURE — NOT PRODUCT CODE.

CPU restatement (torch, fp32 or fp64, autograd-capable) of the Sat-NeRF volumetric rendering hot
path of the reference (centreborelli/satnerf @ 700a5919).  Only `tests/`, `__graft_entry__.smoke()`
and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import this module; the product
(`satnerf_b200/`) never does.

Parity status: PINNED.  The reference has no golden vectors of its own, so this file
is pinned against outputs of the reference itself: `tests/golden/make_golden.py` imports
`/root/reference/rendering.py` + `models/` in the build container, captures every random draw, and
commits inputs/outputs under `tests/golden/*.npz`; `tests/test_oracle_golden.py` replays them here.
`tests/test_oracle_vs_reference.py` additionally compares live (bit-identical values, gradients to 1e-6) whenever
`/root/reference` is present (the build container).

Every function cites the reference lines it restates.  Parameters are passed as a plain
`dict[str, Tensor]` keyed like the reference modules' `state_dict()` (e.g. ``fc_net.2.weight``).
All randomness is explicit: callers pass the uniform / normal tensors the reference would draw
(`rand_like` rendering.py:77, `randn` satnerf.py:58, `rand` rendering.py:33) or a `Draws` object
that draws them from the torch global generator in the reference's order.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


# ----------------------------------------------------------------------------------------------
# random-number plumbing
# ----------------------------------------------------------------------------------------------
class Draws:
    """Source of the random tensors consumed by one `render_rays` call.

    Order of consumption in the reference (SURVEY.md §7 "RNG parity"):
      rand_like(R,S) -> randn(R,S) [-> randn(R,S) for the solar-correction pass]
      -> rand(R,N_imp) -> randn(R,S+N_imp) [-> randn again for SC].
    With `tape=None` the tensors are drawn from torch's global generator in exactly that order
    (so seeding identically to a reference run reproduces it); with a tape (list of tensors) they
    are replayed.
    """

    def __init__(self, tape: Optional[List[torch.Tensor]] = None, dtype=torch.float32, device="cpu"):
        self.tape = None if tape is None else list(tape)
        self.dtype = dtype
        self.device = device
        self.log: List[torch.Tensor] = []

    def _next(self, kind, shape):
        if self.tape is not None:
            t = self.tape.pop(0)
            assert tuple(t.shape) == tuple(shape), (kind, t.shape, shape)
        else:
            t = torch.rand(*shape, device=self.device) if kind == "u" else torch.randn(*shape, device=self.device)
        self.log.append(t)
        return t.to(self.dtype)

    def uniform(self, *shape):
        return self._next("u", shape)

    def normal(self, *shape):
        return self._next("n", shape)


# ----------------------------------------------------------------------------------------------
# samplers
# ----------------------------------------------------------------------------------------------
def stratified_depths(near, far, n_samples, u):
    """rendering.py:65-78.  near/far (R,1); u (R,S) uniform in [0,1).  perturb is hard-wired to 1."""
    steps = torch.linspace(0, 1, n_samples, device=near.device).to(near.dtype)          # :65 (device: the bench's gpu_eager leg runs this file on CUDA)
    z = near * (1 - steps) + far * steps                            # :67
    mid = 0.5 * (z[:, :-1] + z[:, 1:])                              # :72
    hi = torch.cat([mid, z[:, -1:]], -1)                            # :74
    lo = torch.cat([z[:, :1], mid], -1)                             # :75
    return lo + (hi - lo) * u                                       # :77-78


def importance_depths(bins, weights, u, eps=1e-5, return_index=False):
    """rendering.py:10-49 (`sample_pdf`, det=False).  bins (R,M+1), weights (R,M), u (R,N)."""
    m = weights.shape[1]
    w = weights + eps                                               # :23
    pdf = w / w.sum(-1, keepdim=True)                               # :24
    cdf = torch.cat([torch.zeros_like(pdf[:, :1]), pdf.cumsum(-1)], -1)  # :25-26
    u = u.contiguous()
    k = torch.searchsorted(cdf, u, right=True)                      # :36
    lo = (k - 1).clamp_min(0)                                       # :37
    hi = k.clamp_max(m)                                             # :38
    c_lo, c_hi = cdf.gather(1, lo), cdf.gather(1, hi)               # :41
    b_lo, b_hi = bins.gather(1, lo), bins.gather(1, hi)             # :42
    den = c_hi - c_lo                                               # :44
    den = torch.where(den < eps, torch.ones_like(den), den)         # :45
    z = b_lo + (u - c_lo) / den * (b_hi - b_lo)                     # :48
    if return_index:
        return z, k, cdf
    return z


# ----------------------------------------------------------------------------------------------
# fields (MLPs)
# ----------------------------------------------------------------------------------------------
def _lin(p: Params, name: str, x):
    return F.linear(x, p[name + ".weight"], p[name + ".bias"])


def freq_encode(x, n_freqs):
    """models/nerf.py:36-69 (`Mapping`): [sin(2^k x), cos(2^k x)] for k<n_freqs; x itself excluded."""
    out = []
    for k in range(n_freqs):
        f = float(2 ** k)
        out += [torch.sin(f * x), torch.cos(f * x)]
    return torch.cat(out, -1)


def _trunk(p: Params, x, n_layers, skips, siren):
    """fc_net of all three variants: models/satnerf.py:172-179 (= snerf.py, nerf.py:200-207)."""
    h = x
    for i in range(n_layers):
        if i in skips:
            h = torch.cat([x, h], -1)                               # satnerf.py:176-177
        h = _lin(p, f"fc_net.{2 * i}", h)
        if siren:
            h = torch.sin((30.0 if i == 0 else 1.0) * h)           # Siren(w0), nerf.py:33; w0=30 satnerf.py:106
        else:
            h = torch.relu(h)
    return h


def field_satnerf(p: Params, xyz, sun_d, t_emb=None, n_layers=8, skips=(4,), with_beta=True):
    """SatNeRF.forward (models/satnerf.py:156-208); with_beta=False gives ShadowNeRF.forward
    (models/snerf.py:148-196).  Returns (B,9) / (B,8): [rgb3, sigma, sun, sky3, (beta)]."""
    feat = _trunk(p, xyz, n_layers, skips, siren=True)
    sigma = F.softplus(_lin(p, "sigma_from_xyz.0", feat))           # :183
    f2 = _lin(p, "feats_from_xyz", feat)                            # :188
    rgb = torch.sigmoid(_lin(p, "rgb_from_xyzdir.2", torch.sin(_lin(p, "rgb_from_xyzdir.0", f2))))  # :193
    rgb = rgb * (1 + 2 * 0.001) - 0.001                             # :195
    s = torch.cat([f2, sun_d], -1)                                  # :199
    for j in (0, 2, 4):
        s = torch.sin(_lin(p, f"sun_v_net.{j}", s))
    s = torch.sigmoid(_lin(p, "sun_v_net.6", s))                    # :200
    sky = torch.sigmoid(_lin(p, "sky_color.2", torch.relu(_lin(p, "sky_color.0", sun_d))))  # :201
    cols = [rgb, sigma, s, sky]
    if with_beta:
        b = torch.sin(_lin(p, "beta_from_xyz.0", torch.cat([f2, t_emb], -1)))   # :204
        cols.append(F.softplus(_lin(p, "beta_from_xyz.2", b)))      # :205
    return torch.cat(cols, 1)


def field_nerf(p: Params, xyz, view_d, n_layers=8, skips=(4,), pe=(10, 4)):
    """NeRF.forward (models/nerf.py:184-227), mapping=True, siren=False.  Returns (B,4)."""
    ex = freq_encode(xyz, pe[0])                                    # :201
    feat = _trunk(p, ex, n_layers, skips, siren=False)
    sigma = F.softplus(_lin(p, "sigma_from_xyz.0", feat))           # :211
    f2 = _lin(p, "feats_from_xyz", feat)                            # :216
    hin = torch.cat([f2, freq_encode(view_d, pe[1])], -1)           # :218
    rgb = torch.sigmoid(_lin(p, "rgb_from_xyzdir.2", torch.relu(_lin(p, "rgb_from_xyzdir.0", hin))))
    rgb = rgb * (1 + 2 * 0.001) - 0.001                             # :223
    return torch.cat([rgb, sigma], 1)


# ----------------------------------------------------------------------------------------------
# compositing
# ----------------------------------------------------------------------------------------------
def composite(raw, z, noise, variant):
    """Alpha compositing of one pass: models/satnerf.py:43-78 (snerf.py:43-74, nerf.py:108-132).
    raw (R,S,C) field outputs; z (R,S); noise (R,S) already scaled by noise_std."""
    sigma = raw[..., 3]
    delta = torch.cat([z[:, 1:] - z[:, :-1], torch.full_like(z[:, :1], 1e10)], -1)     # :52-54
    alpha = 1 - torch.exp(-delta * torch.relu(sigma + noise))                             # :59
    shifted = torch.cat([torch.ones_like(alpha[:, :1]), 1 - alpha + 1e-10], -1)           # :60-61
    trans = torch.cumprod(shifted, -1)[:, :-1]                                            # :62
    w = alpha * trans                                                                     # :63
    out = {"depth": (w * z).sum(-1), "weights": w, "transparency": trans}                 # :67
    rgbs = raw[..., :3]
    if variant == "nerf":
        out["rgb"] = (w.unsqueeze(-1) * rgbs).sum(-2)                                     # nerf.py:128
        return out
    sun, sky = raw[..., 4:5], raw[..., 5:8]
    irr = sun + (1 - sun) * sky                                                           # :68
    out["rgb"] = (w.unsqueeze(-1) * rgbs * irr).sum(-2).clamp(0.0, 1.0)                   # :69-70
    out.update(albedo=rgbs, sun=sun, sky=sky)
    if variant == "sat-nerf":
        out["beta"] = raw[..., 8:9]
    return out


def _order(res, variant):
    keys = ["rgb", "depth", "weights", "transparency"]
    if variant != "nerf":
        keys += ["albedo", "sun", "sky"]
    if variant == "sat-nerf":
        keys += ["beta"]
    return {k: res[k] for k in keys}


def run_pass(p: Params, variant, origins, dirs, z, noise, sun_d=None, t_emb=None, view_d=None,
             n_layers=8, skips=(4,)):
    """`inference()` of the reference: broadcast per-ray inputs to the samples (satnerf.py:25-27),
    evaluate the field at every point (chunking at :30-40 does not change values) and composite."""
    r, s = z.shape
    xyz = (origins[:, None, :] + dirs[:, None, :] * z[:, :, None]).reshape(r * s, 3)   # rendering.py:81
    rep = lambda v: None if v is None else v.repeat_interleave(s, 0)
    if variant == "nerf":
        raw = field_nerf(p, xyz, rep(view_d), n_layers, skips)
    else:
        raw = field_satnerf(p, xyz, rep(sun_d), rep(t_emb), n_layers, skips,
                            with_beta=(variant == "sat-nerf"))
    return _order(composite(raw.reshape(r, s, -1), z, noise, variant), variant)


# ----------------------------------------------------------------------------------------------
# top level
# ----------------------------------------------------------------------------------------------
def render_rays(params: Dict[str, Params], cfg, rays, ts, draws: Optional[Draws] = None):
    """rendering.py:52-158.  `params` = {'coarse': state-dict-like, 'fine': ..., 't': (vocab,tau) tensor}.
    `cfg` needs: model, n_samples, n_importance, noise_std, sc_lambda, fc_layers (skips fixed to [4]).
    Reproduces the reference including the RNG consumption order; the two code paths the reference
    cannot execute (s-nerf + fine: NameError at :134; fine + SC: result overwritten at :138/:149,
    SURVEY.md App. B) raise NotImplementedError instead of imitating the crash."""
    draws = draws or Draws(dtype=rays.dtype, device=rays.device)
    variant, s, n_imp = cfg.model, cfg.n_samples, cfg.n_importance
    n_layers = getattr(cfg, "fc_layers", 8)
    o, d, near, far = rays[:, 0:3], rays[:, 3:6], rays[:, 6:7], rays[:, 7:8]               # :62
    z = stratified_depths(near, far, s, draws.uniform(*near.expand(-1, s).shape))          # :65-78
    sun_d = rays[:, 8:11] if variant != "nerf" else None                                    # :87/:99
    t_emb = params["t"][ts] if (variant == "sat-nerf" and ts is not None) else None         # :100

    def one_level(level, zz):
        r = zz.shape[0]
        noise = draws.normal(r, zz.shape[1]) * cfg.noise_std                                # satnerf.py:57-58
        res = run_pass(params[level], variant, o, d, zz, noise, sun_d, t_emb, view_d=d, n_layers=n_layers)
        if variant != "nerf" and cfg.sc_lambda > 0:                                          # :90-96 / :102-108
            noise2 = draws.normal(r, zz.shape[1]) * cfg.noise_std
            sc = run_pass(params[level], variant, o, sun_d, zz, noise2, sun_d, t_emb, n_layers=n_layers)
            res["weights_sc"], res["transparency_sc"], res["sun_sc"] = sc["weights"], sc["transparency"], sc["sun"]
        return {f"{k}_{level}": v for k, v in res.items()}                                   # :113-115

    out = one_level("coarse", z)
    if n_imp > 0:                                                                            # :118
        if variant == "s-nerf" or (variant == "sat-nerf" and cfg.sc_lambda > 0):
            raise NotImplementedError("reference cannot run this combination (SURVEY.md App. B)")
        mid = 0.5 * (z[:, :-1] + z[:, 1:])                                                   # :121
        zf = importance_depths(mid, out["weights_coarse"][:, 1:-1], draws.uniform(z.shape[0], n_imp)).detach()  # :122-123
        z2, _ = torch.sort(torch.cat([z, zf], -1), -1)                                       # :125
        out.update(one_level("fine", z2))
    return out


# ----------------------------------------------------------------------------------------------
# losses (metrics.py) — the consumers that decide which outputs carry gradient
# ----------------------------------------------------------------------------------------------
def _sc_terms(losses, res, level, lam):
    """metrics.py:26-34 (`solar_correction`)."""
    sun = res[f"sun_sc_{level}"].squeeze()
    t2 = ((res[f"transparency_sc_{level}"].detach() - sun) ** 2).sum(-1)
    t3 = 1 - (res[f"weights_sc_{level}"].detach() * sun).sum(-1)
    losses[f"{level}_sc_term2"] = lam / 3.0 * t2.mean()
    losses[f"{level}_sc_term3"] = lam / 3.0 * t3.mean()


def loss_snerf(res, target, lam_sc=0.0):
    """metrics.py:36-55 (`SNerfLoss`); with lam_sc=0 and only rgb this is `NerfLoss` (:8-19)."""
    losses = {}
    for level in ("coarse", "fine"):
        if f"rgb_{level}" not in res:
            continue
        losses[f"{level}_color"] = F.mse_loss(res[f"rgb_{level}"], target)
        if lam_sc > 0:
            _sc_terms(losses, res, level, lam_sc)
    return sum(losses.values()), losses


def loss_satnerf(res, target, lam_sc=0.0, beta_min=0.05):
    """metrics.py:21-25 + :57-73 (`SatNerfLoss`, coarse level)."""
    losses = {}
    beta = (res["weights_coarse"].unsqueeze(-1) * res["beta_coarse"]).sum(-2) + beta_min
    losses["coarse_color"] = ((res["rgb_coarse"] - target) ** 2 / (2 * beta ** 2)).mean()
    losses["coarse_logbeta"] = (3 + torch.log(beta).mean()) / 2
    if lam_sc > 0:
        _sc_terms(losses, res, "coarse", lam_sc)
    return sum(losses.values()), losses


def loss_depth(res, target, weights=1.0, lam_ds=1.0):
    """metrics.py:75-92 (`DepthLoss`)."""
    losses = {}
    for level in ("coarse", "fine"):
        if f"depth_{level}" in res:
            losses[f"{level}_ds"] = lam_ds / 3.0 * (weights * (res[f"depth_{level}"] - target) ** 2).mean()
    return sum(losses.values()), losses


# ----------------------------------------------------------------------------------------------
# synthetic inputs: the seeded generators live with the product's bench inputs (satnerf_b200/synth.py) so that the product
# bench never imports the checker; re-exported here for the tests.
# ----------------------------------------------------------------------------------------------
from satnerf_b200.synth import synthetic_blender_rays, synthetic_sat_rays  # noqa: E402,F401
