"""Developer stress run of the fused tensor-core path: many (width, rays, samples, mode) combinations, each checked against
the fp32 CUDA-core path (1e-3) and for run-to-run bit equality; training mode also runs the backward.  A pipeline deadlock
shows up as a trapped launch (bounded waits), a race as a mismatch."""
import itertools, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import torch
import satnerf_b200 as sb
from gpu_util import make_args
from golden_io import rel_err
from oracle import render_oracle as orc

t0 = time.time(); n = 0; worst = 0.0
for model, h, S in itertools.product(("sat-nerf", "s-nerf"), (64, 128, 256, 320, 384, 512), (33, 48, 64, 96, 128)):
    args = make_args(model=model, fc_units=h, n_samples=S, precision="tc")
    torch.manual_seed(h + S)
    ms = {"coarse": sb.load_model(args).cuda()}
    if model == "sat-nerf":
        ms["t"] = torch.nn.Embedding(30, 4).cuda()
    for R in (1, 2, 3, 7, 150, 297, 1500):
        rays, ts = orc.synthetic_sat_rays(R, seed=R)
        rays = rays.cuda(); ts = ts.cuda() if model == "sat-nerf" else None
        g = torch.Generator().manual_seed(R + 1)
        draws = [torch.rand(R, S, generator=g), torch.randn(R, S, generator=g)]
        with torch.no_grad():
            args.precision = "tc"
            a = sb.render_rays(ms, args, rays, ts, _draws=draws)
            a2 = sb.render_rays(ms, args, rays, ts, _draws=draws)
            args.precision = "fp32"
            b = sb.render_rays(ms, args, rays, ts, _draws=draws)
        for k in a:
            assert torch.equal(a[k], a2[k]), (model, h, S, R, k, "not deterministic")
            e = rel_err(a[k], b[k]); worst = max(worst, e)
            assert e < 2e-3, (model, h, S, R, k, e)
        if R in (3, 150, 297):
            args.precision = "tc"
            res = sb.render_rays(ms, args, rays, ts, _draws=draws)
            loss = (res["rgb_coarse"] ** 2).mean() + (res["weights_coarse"][:, 3] ** 2).mean()
            loss.backward()
            gsum = sum(float(p.grad.abs().sum()) for p in ms["coarse"].parameters())
            assert gsum == gsum and gsum > 0, (model, h, S, R, "bad gradient")
            for p in ms["coarse"].parameters():
                p.grad = None
        n += 1
    torch.cuda.synchronize()
print(f"stress ok: {n} configurations, worst tc-vs-fp32 error {worst:.2e}, {time.time() - t0:.0f} s")
