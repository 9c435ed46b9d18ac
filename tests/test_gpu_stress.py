"""Sweep of the fused tensor-core pipeline over widths, sample counts and ray counts (odd group counts, partial tiles, widths
whose N-chunks are not whole K-slabs, single-CTA fallback for tiny batches), inference and training mode: every case must agree with
the fp32 CUDA-core path within the north_star tolerance (1e-3, measured like the other parity tests) and be bit-reproducible.
The layer pipeline hands tiles between the MMA issuer, the weight stream and the epilogue through ~10 mbarriers; a protocol
mistake shows up here as a trapped launch (all waits are bounded) or as a mismatch."""
import itertools

import pytest
import torch

from golden_io import rel_err
from gpu_util import make_args
from oracle import render_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("model", ["sat-nerf", "s-nerf"])
def test_pipeline_sweep(model):
    import satnerf_b200 as sb
    worst = 0.0
    for h, S in itertools.product((64, 256, 320, 384, 512), (33, 64, 96)):
        args = make_args(model=model, fc_units=h, n_samples=S, precision="tc")
        torch.manual_seed(h + S)
        ms = {"coarse": sb.load_model(args).cuda()}
        if model == "sat-nerf":
            ms["t"] = torch.nn.Embedding(30, 4).cuda()
        for R in (1, 3, 150, 297):
            rays, ts = orc.synthetic_sat_rays(R, seed=R)
            rays = rays.cuda()
            ts = ts.cuda() if model == "sat-nerf" else None
            g = torch.Generator().manual_seed(R + 1)
            draws = [torch.rand(R, S, generator=g), torch.randn(R, S, generator=g)]
            with torch.no_grad():
                args.precision = "tc"
                a = sb.render_rays(ms, args, rays, ts, _draws=draws)
                a2 = sb.render_rays(ms, args, rays, ts, _draws=draws)
                args.precision = "fp32"
                b = sb.render_rays(ms, args, rays, ts, _draws=draws)
            for k in a:
                assert torch.equal(a[k], a2[k]), (h, S, R, k, "not run-to-run deterministic")
                e = rel_err(a[k], b[k])
                worst = max(worst, e)
                assert e < 1e-3, (h, S, R, k, e)
            if R in (3, 297):                       # training mode: activation stash + tensor-core backward
                args.precision = "tc"
                res = sb.render_rays(ms, args, rays, ts, _draws=draws)
                ((res["rgb_coarse"] ** 2).mean() + (res["weights_coarse"][:, 3] ** 2).mean()).backward()
                for p in ms["coarse"].parameters():
                    assert p.grad is not None and torch.isfinite(p.grad).all()
                    p.grad = None
    torch.cuda.synchronize()
    assert worst < 1e-3
