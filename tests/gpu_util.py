"""Helpers for the -m gpu parity tests: build product models from a golden case and run the CUDA path."""
import argparse
import copy

import torch

import satnerf_b200 as sb
from golden_io import Golden


def models_from_golden(g: Golden, device="cuda"):
    args = copy.copy(g.cfg)
    ms = {}
    for lvl in ("coarse", "fine"):
        if lvl in g.params:
            m = sb.load_model(args)
            m.load_state_dict(g.params[lvl])
            ms[lvl] = m.to(device)
    if "t" in g.params:
        emb = torch.nn.Embedding(args.t_embbeding_vocab, args.t_embbeding_tau)
        with torch.no_grad():
            emb.weight.copy_(g.params["t"])
        ms["t"] = emb.to(device)
    return ms, args


def run_golden(g: Golden, precision, device="cuda"):
    ms, args = models_from_golden(g, device)
    args.precision = precision
    ts = None if g.ts is None else g.ts.to(device)
    res = sb.render_rays(ms, args, g.rays.to(device), ts, _draws=g.draws)
    return ms, args, res


def make_args(**kw):
    base = dict(model="sat-nerf", n_samples=64, n_importance=0, noise_std=0.0, sc_lambda=0.0, chunk=5120,
                fc_layers=8, fc_units=512, t_embbeding_tau=4, t_embbeding_vocab=30)
    base.update(kw)
    return argparse.Namespace(**base)
