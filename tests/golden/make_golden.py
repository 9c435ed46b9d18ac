"""
Generates the committed golden fixtures `tests/golden/*.npz` by running the UNMODIFIED reference
(`/root/reference/rendering.py` + `/root/reference/models/`) on seeded synthetic rays.

Run in the build container only (the reference is not present on the GPU box):

    python tests/golden/make_golden.py

Every random tensor the reference draws (rand_like rendering.py:77, randn satnerf.py:58 /
snerf.py:54 / nerf.py:118, rand rendering.py:33) is recorded in call order by wrapping the three
torch factory functions for the duration of the call, so tests can replay them.
Weights: small cases store the reference's own default-initialised `state_dict()`; the h=512 case
uses weights regenerated from a numpy PCG64 stream (see `pcg_params`) so that only the seed and the
outputs need to be stored.
"""
import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

import types  # noqa: E402

import rendering as ref_rendering  # noqa: E402  (the reference)
import models as ref_models  # noqa: E402

# metrics.py imports kornia.losses.ssim (not installed; used only by the ssim metric) -> stub it so
# the reference's own loss classes can drive the gradient fixtures.
_k = types.ModuleType("kornia"); _kl = types.ModuleType("kornia.losses"); _kl.ssim = None; _k.losses = _kl
sys.modules.setdefault("kornia", _k); sys.modules.setdefault("kornia.losses", _kl)
import metrics as ref_metrics  # noqa: E402

from oracle.render_oracle import synthetic_sat_rays, synthetic_blender_rays  # noqa: E402  (input generators only)


class Recorder:
    """Records outputs of torch.rand_like / torch.randn / torch.rand while active."""

    def __enter__(self):
        self.tape = []
        self._orig = (torch.rand_like, torch.randn, torch.rand)

        def wrap(fn):
            def inner(*a, **k):
                t = fn(*a, **k)
                self.tape.append(t.detach().clone())
                return t
            return inner

        torch.rand_like, torch.randn, torch.rand = (wrap(f) for f in self._orig)
        return self

    def __exit__(self, *exc):
        torch.rand_like, torch.randn, torch.rand = self._orig


def pcg_params(shapes, seed):
    """Deterministic weights from numpy PCG64 with the scales load_model() produces: U(+-sqrt(6/fan_in)) for the
    sine_init layers (fc_net, sun_v_net; U(+-1/fan_in) for their first layers), nn.Linear's default
    U(+-1/sqrt(fan_in)) for every other weight and for all biases."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = {}
    fan = {}
    for name, shape in shapes.items():
        u = rng.random(int(np.prod(shape)), dtype=np.float32).reshape(shape) * 2 - 1
        if name.endswith(".weight"):
            fan[name[:-7]] = shape[1]
            siren = name.startswith(("fc_net.", "sun_v_net."))      # sine_init layers (satnerf.py:145-149); others keep nn.Linear's default scale
            bound = 1.0 / shape[1] if name in ("fc_net.0.weight", "sun_v_net.0.weight") else (np.sqrt(6.0 / shape[1]) if siren else 1.0 / np.sqrt(shape[1]))
        else:
            bound = 1.0 / np.sqrt(fan[name[:-5]])
        out[name] = (u * np.float32(bound)).astype(np.float32)
    return out


def trained_like(sd):
    """'Trained-like' regime: default init gives sigma ~ 0.7 (60 % of the weight on the last sample); here the
    density head is sharpened (weight x64, bias -2: rays saturate mid-way, alpha up to ~0.3 per sample) and the head layers
    carry twice their init scale, as weights grown by training do.  Deterministic transform of a state dict (numpy arrays)."""
    out = {k: v.copy() for k, v in sd.items()}
    out["sigma_from_xyz.0.weight"] = out["sigma_from_xyz.0.weight"] * np.float32(64.0)
    out["sigma_from_xyz.0.bias"] = np.full_like(out["sigma_from_xyz.0.bias"], -2.0)
    for k in ("feats_from_xyz.weight", "rgb_from_xyzdir.0.weight", "rgb_from_xyzdir.2.weight", "sun_v_net.6.weight",
              "beta_from_xyz.0.weight", "beta_from_xyz.2.weight"):
        if k in out:
            out[k] = out[k] * np.float32(2.0)
    return out


def make_args(**kw):
    base = dict(model="sat-nerf", n_samples=16, n_importance=0, noise_std=0.0, sc_lambda=0.0, chunk=5120,
                fc_layers=8, fc_units=64, t_embbeding_tau=4, t_embbeding_vocab=30)
    base.update(kw)
    return argparse.Namespace(**base)


def build_models(args, seed, pcg_seed=None, transform=None):
    torch.manual_seed(seed)
    ms = {"coarse": ref_models.load_model(args)}
    if args.n_importance > 0:
        ms["fine"] = ref_models.load_model(args)
    if args.model == "sat-nerf":
        ms["t"] = torch.nn.Embedding(args.t_embbeding_vocab, args.t_embbeding_tau)
    if pcg_seed is not None:
        for i, lvl in enumerate(k for k in ("coarse", "fine") if k in ms):
            shapes = {k: tuple(v.shape) for k, v in ms[lvl].state_dict().items()}
            raw = pcg_params(shapes, pcg_seed + i)
            if transform == "trained_v1":
                raw = trained_like(raw)
            sd = {k: torch.from_numpy(v) for k, v in raw.items()}
            ms[lvl].load_state_dict(sd)
    return ms


def run_case(name, args, n_rays, seed, with_grads=None, pcg_seed=None, store_params=True, transform=None):
    ms = build_models(args, seed, pcg_seed, transform)
    if args.model == "nerf":
        rays, ts = synthetic_blender_rays(n_rays, seed=seed + 1), None
    else:
        rays, ts = synthetic_sat_rays(n_rays, seed=seed + 1)
    torch.manual_seed(seed + 2)
    with Recorder() as rec:
        res = ref_rendering.render_rays(ms, args, rays, ts)
    blob = {"rays": rays.numpy(), "n_draws": np.int64(len(rec.tape))}
    if ts is not None:
        blob["ts"] = ts.numpy()
    for i, t in enumerate(rec.tape):
        blob[f"draw{i}"] = t.numpy()
    for k, v in res.items():
        blob[f"out.{k}"] = v.detach().numpy()
    for lvl in ("coarse", "fine"):
        if lvl in ms and store_params:
            for k, v in ms[lvl].state_dict().items():
                blob[f"param.{lvl}.{k}"] = v.numpy()
    if "t" in ms:
        blob["param.t"] = ms["t"].weight.detach().numpy()
    if pcg_seed is not None:
        blob["pcg_seed"] = np.int64(pcg_seed)
    if transform is not None:
        blob["pcg_transform"] = np.array(transform)
    cfg = {k: getattr(args, k) for k in ("model", "n_samples", "n_importance", "noise_std", "sc_lambda",
                                         "fc_layers", "fc_units", "t_embbeding_tau", "t_embbeding_vocab")}
    blob["cfg"] = np.array(repr(cfg))

    if with_grads:
        g = torch.Generator().manual_seed(seed + 3)
        target = torch.rand(n_rays, 3, generator=g)
        if with_grads == "satnerf":
            loss, _ = ref_metrics.SatNerfLoss(lambda_sc=args.sc_lambda)(res, target)
        elif with_grads == "snerf":
            loss, _ = ref_metrics.SNerfLoss(lambda_sc=args.sc_lambda)(res, target)
        elif with_grads == "depth":
            dt = 0.2 + 0.2 * torch.rand(n_rays, generator=g)
            dw = torch.rand(n_rays, generator=g)
            loss, _ = ref_metrics.DepthLoss(lambda_ds=1000.0)(res, dt, dw)
            blob["depth_target"] = dt.numpy()
            blob["depth_weights"] = dw.numpy()
        loss.backward()
        blob["target"] = target.numpy()
        blob["loss"] = loss.detach().numpy()
        blob["loss_kind"] = np.array(with_grads)
        for lvl in ("coarse", "fine"):
            if lvl in ms:
                for k, v in ms[lvl].named_parameters():
                    if v.grad is not None:
                        blob[f"grad.{lvl}.{k}"] = v.grad.numpy()
        if "t" in ms and ms["t"].weight.grad is not None:
            blob["grad.t"] = ms["t"].weight.grad.numpy()
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **blob)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} kB, draws={len(rec.tape)}, keys={sorted(res.keys())}")


def sample_pdf_case():
    """Direct fixture for rendering.py:10-49 including the int64 bin indices (recomputed with the
    same torch call the reference makes at :36 on the reference-built cdf)."""
    g = torch.Generator().manual_seed(77)
    r, m, n = 64, 62, 32
    z = torch.sort(torch.rand(r, m + 2, generator=g), -1)[0]
    bins = 0.5 * (z[:, :-1] + z[:, 1:])
    w = torch.rand(r, m, generator=g) ** 4
    w[::7] = 0.0                                    # rays with all-zero weights (uniform pdf)
    w[1::7, 10:50] = 0.0                            # empty bins -> denom<eps branch
    torch.manual_seed(5)
    with Recorder() as rec:
        out = ref_rendering.sample_pdf(bins, w, n)
    u = rec.tape[0]
    ww = w + 1e-5
    cdf = torch.cat([torch.zeros(r, 1), torch.cumsum(ww / ww.sum(-1, keepdim=True), -1)], -1)
    inds = torch.searchsorted(cdf, u.contiguous(), right=True)
    np.savez_compressed(os.path.join(HERE, "sample_pdf.npz"), z=z.numpy(), bins=bins.numpy(), weights=w.numpy(), u=u.numpy(),
                        cdf=cdf.numpy(), inds=inds.numpy(), samples=out.numpy())
    print("sample_pdf: ok")


if __name__ == "__main__":
    torch.set_num_threads(1)   # pins the reduction order of the reference's torch.sum on this host
    run_case("satnerf_h64", make_args(noise_std=0.3), 24, 10, with_grads="satnerf")
    run_case("satnerf_h64_snerfloss_depth", make_args(), 24, 11, with_grads="depth")
    run_case("satnerf_sc_h64", make_args(sc_lambda=0.05), 24, 12, with_grads="satnerf")
    run_case("satnerf_fine_h64", make_args(n_importance=8), 24, 13, with_grads="snerf")
    run_case("snerf_sc_h64", make_args(model="s-nerf", sc_lambda=0.05, noise_std=0.1), 24, 14, with_grads="snerf")
    run_case("nerf_fine_h64", make_args(model="nerf", n_importance=8), 24, 15, with_grads="snerf")
    run_case("satnerf_h512", make_args(fc_units=512, n_samples=64), 8, 16, pcg_seed=1234, store_params=False)
    run_case("satnerf_h256_s96", make_args(fc_units=256, n_samples=96), 6, 17, pcg_seed=4321, store_params=False)
    run_case("satnerf_h512_trained", make_args(fc_units=512, n_samples=64), 16, 18, pcg_seed=1234, store_params=False, transform="trained_v1")
    sample_pdf_case()
