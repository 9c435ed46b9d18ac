"""CPU, build container only: the oracle against the LIVE unmodified reference (`/root/reference/rendering.py` + `models/`)
on fresh seeded inputs — values and gradients, every variant the reference can execute.  Skipped where the reference is
absent (the GPU box); the committed fixtures (tests/golden/) carry the same pinning there."""
import argparse
import os
import sys

import pytest
import torch

from oracle import render_oracle as orc

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "rendering.py")), reason="reference tree not present")


def _ref():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import models as ref_models
    import rendering as ref_rendering
    return ref_rendering, ref_models


def _args(**kw):
    base = dict(model="sat-nerf", n_samples=32, n_importance=0, noise_std=0.0, sc_lambda=0.0, chunk=5120, fc_layers=8, fc_units=64,
                t_embbeding_tau=4, t_embbeding_vocab=30)
    base.update(kw)
    return argparse.Namespace(**base)


CASES = [dict(), dict(fc_units=256, n_samples=64), dict(sc_lambda=0.05, noise_std=0.2), dict(n_importance=16),
         dict(model="s-nerf", sc_lambda=0.05), dict(model="nerf", n_importance=8, fc_units=128)]


@pytest.mark.parametrize("kw", CASES, ids=lambda k: "-".join(f"{a}={b}" for a, b in k.items()) or "default")
def test_oracle_equals_live_reference(kw):
    rr, rm = _ref()
    args = _args(**kw)
    torch.manual_seed(3)
    ms = {"coarse": rm.load_model(args)}
    if args.n_importance > 0:
        ms["fine"] = rm.load_model(args)
    if args.model == "sat-nerf":
        ms["t"] = torch.nn.Embedding(30, 4)
    n = 40
    rays, ts = (orc.synthetic_blender_rays(n, seed=4), None) if args.model == "nerf" else orc.synthetic_sat_rays(n, seed=4)
    torch.manual_seed(5)
    want = rr.render_rays(ms, args, rays, ts)
    params = {k: ({n_: p.detach().clone().requires_grad_(True) for n_, p in m.state_dict().items()} if k != "t"
                  else m.weight.detach().clone().requires_grad_(True)) for k, m in ms.items()}
    torch.manual_seed(5)                      # same generator state: Draws() draws in the reference's order
    got = orc.render_rays(params, args, rays, ts)
    assert set(got) == set(want)
    for k in want:
        assert torch.equal(got[k], want[k]), (k, float((got[k] - want[k]).abs().max()))
    # gradients of a loss touching every differentiated output
    typ = "fine" if args.n_importance > 0 else "coarse"

    def loss(res):
        val = (res[f"rgb_{typ}"] ** 2).mean() + res[f"depth_{typ}"].mean() + (res[f"weights_{typ}"] ** 2).sum(-1).mean()
        if f"beta_{typ}" in res:
            val = val + res[f"beta_{typ}"].mean()
        if f"sun_sc_{typ}" in res:
            val = val + (res[f"sun_sc_{typ}"] ** 2).mean()
        return val

    loss(want).backward()
    loss(got).backward()
    for lvl in ("coarse", "fine"):
        if lvl in ms:
            for name, p in ms[lvl].named_parameters():
                g_ref, g = p.grad, params[lvl][name].grad
                if g_ref is None:
                    assert g is None or float(g.abs().max()) == 0.0, name
                    continue
                assert g is not None and float((g - g_ref).abs().max()) <= 1e-6 * max(1.0, float(g_ref.abs().max())), (lvl, name)
