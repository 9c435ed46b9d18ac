"""CPU: pins oracle/render_oracle.py against the golden fixtures produced by the unmodified reference."""
import numpy as np
import pytest
import torch

from golden_io import CASES, GOLDEN_DIR, TRAINED_CASES, Golden, rel_err
from oracle import render_oracle as orc


def _loss(g, res):
    if g.loss_kind == "satnerf":
        return orc.loss_satnerf(res, g.target, lam_sc=g.cfg.sc_lambda)[0]
    if g.loss_kind == "snerf":
        return orc.loss_snerf(res, g.target, lam_sc=g.cfg.sc_lambda)[0]
    return orc.loss_depth(res, g.depth_target, g.depth_weights, lam_ds=1000.0)[0]


@pytest.mark.parametrize("name", CASES + TRAINED_CASES)
def test_forward_matches_reference(name):
    g = Golden(name)
    res = orc.render_rays(g.params, g.cfg, g.rays, g.ts, orc.Draws(g.draws))
    assert set(res) == set(g.out)
    for k, ref in g.out.items():
        assert res[k].shape == ref.shape, k
        # same torch ops on the same host: agreement is at fp32 rounding level
        assert rel_err(res[k], ref) < 2e-6, (k, rel_err(res[k], ref))


@pytest.mark.parametrize("name", [c for c in CASES if "h64" in c])
def test_gradients_match_reference(name):
    g = Golden(name)
    params = {lvl: {k: v.clone().requires_grad_(True) for k, v in p.items()} if lvl != "t" else p.clone().requires_grad_(True)
              for lvl, p in g.params.items()}
    res = orc.render_rays(params, g.cfg, g.rays, g.ts, orc.Draws(g.draws))
    loss = _loss(g, res)
    assert abs(float(loss) - g.loss) <= 1e-5 * max(1.0, abs(g.loss))
    loss.backward()
    seen = 0
    for key, ref in g.grads.items():
        lvl, _, pname = key.partition(".")
        got = params["t"].grad if key == "t" else params[lvl][pname].grad
        assert got is not None, key
        assert rel_err(got, ref, floor=1e-9) < 5e-5, (key, rel_err(got, ref, floor=1e-9))
        seen += 1
    assert seen > 10


def test_sample_pdf_indices_bit_exact():
    z = np.load(f"{GOLDEN_DIR}/sample_pdf.npz")
    bins, w, u = (torch.from_numpy(z[k]) for k in ("bins", "weights", "u"))
    samples, k, cdf = orc.importance_depths(bins, w, u, return_index=True)
    assert torch.equal(k, torch.from_numpy(z["inds"]))
    assert torch.equal(cdf, torch.from_numpy(z["cdf"]))
    assert torch.equal(samples, torch.from_numpy(z["samples"]))


def test_double_precision_oracle_agrees():
    g = Golden("satnerf_h64")
    p64 = {lvl: ({k: v.double() for k, v in p.items()} if lvl != "t" else p.double()) for lvl, p in g.params.items()}
    res = orc.render_rays(p64, g.cfg, g.rays.double(), g.ts, orc.Draws(g.draws, dtype=torch.float64))
    for k, ref in g.out.items():
        assert rel_err(res[k], ref) < 1e-4, k
