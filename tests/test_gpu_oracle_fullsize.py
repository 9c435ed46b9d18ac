"""GPU parity at BASELINE.json's sizes, tensor-core path DIRECTLY against the CPU oracle (fp32, pinned to the unmodified
reference by tests/golden + tests/test_oracle_vs_reference.py) — not against the repo's own fp32 path.

Tolerance (written here): every element  |got - ref| <= 1e-3 * |ref| + atol(key)   (golden_io.ATOL: 2e-5 for weights, 1e-4
for everything else — outputs that reach 0 need an absolute floor; 1e-3 is north_star's relative tolerance), in addition to
the max-abs / max-ref measure used elsewhere.  Measured maxima are appended to gpurun_out/parity_report.json.

The 'trained-like' regime (density head sharpened x64 so that rays saturate mid-way) is reported separately: the fp16-operand
contractions carry ~1e-4 relative error on the trunk features, which the sharp density head multiplies — the fp32 path and
the hi+lo tensor-core path (precision 'tcx3') are gated at 1e-3 there and the fp16-operand kernel at the bound written in
`test_trained_like_regime`.
"""
import json
import os

import pytest
import torch

from golden_io import TRAINED_CASES, Golden, elementwise_excess, field_shapes, pcg_params, rel_err, trained_like
from gpu_util import make_args, run_golden
from oracle import render_oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _report(name, rows):
    try:
        d = os.path.join(ROOT, "gpurun_out")
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_report.json"), "a") as f:
            f.write(json.dumps({"test": name, "rows": rows}) + "\n")
    except OSError:
        pass


def _models(args, seed):
    import satnerf_b200 as sb
    torch.manual_seed(seed)
    ms = {"coarse": sb.load_model(args)}
    if args.n_importance > 0:
        ms["fine"] = sb.load_model(args)
    if args.model == "sat-nerf":
        ms["t"] = torch.nn.Embedding(30, 4)
    params = {k: ({n: p.detach().clone() for n, p in m.state_dict().items()} if k != "t" else m.weight.detach().clone()) for k, m in ms.items()}
    return {k: m.cuda() for k, m in ms.items()}, params


def _draws(R, S, N=0, sc=False, seed=0):
    g = torch.Generator().manual_seed(seed)
    d = [torch.rand(R, S, generator=g), torch.randn(R, S, generator=g)]
    if sc:
        d.append(torch.randn(R, S, generator=g))
    if N:
        d += [torch.rand(R, N, generator=g), torch.randn(R, S + N, generator=g)]
    return d


def _compare(name, got, want, keys=None, rtol=1e-3, gate=1.0):
    rows, worst = {}, 0.0
    for k in (keys or want):
        a, b = got[k].detach().cpu(), want[k]
        assert a.shape == b.shape and torch.isfinite(a).all(), k
        ex, re = elementwise_excess(a, b, k, rtol), rel_err(a, b)
        rows[k] = {"elementwise_excess": ex, "max_rel": re}
        worst = max(worst, ex)
    _report(name, rows)
    bad = {k: v for k, v in rows.items() if v["elementwise_excess"] > gate or v["max_rel"] > rtol * gate}
    assert not bad, (name, bad)
    return rows


def test_config2_tc_vs_oracle():
    """configs[1]: sat-nerf h=512, 4096 rays x 64 samples."""
    import satnerf_b200 as sb
    args = make_args(noise_std=0.1, precision="tc")
    ms, params = _models(args, 40)
    rays, ts = orc.synthetic_sat_rays(4096, seed=41)
    draws = _draws(4096, 64, seed=42)
    want = orc.render_rays(params, args, rays, ts, orc.Draws(draws))
    with torch.no_grad():
        got = sb.render_rays(ms, args, rays.cuda(), ts.cuda(), _draws=draws)
    assert set(got) == set(want)
    _compare("config2_tc_vs_oracle", got, want)


def test_config3_coarse_fine_tc_vs_oracle():
    """configs[2]: sat-nerf 64 coarse + 32 importance samples (96 fine), 8192 rays.
      (1) coarse level: strict (elementwise 1e-3).
      (2) fine level DECOUPLED from the coarse one: the fine field evaluated through `inference()` at the ORACLE's merged depths
          (S = 96): strict.
      (3) end to end: the fine depths are importance-sampled from the path's own coarse weights (rendering.py:121-125), 1e-4 away
          from the reference's, which shifts samples inside their bins (and, rarely, across a bin edge): per-ray fine outputs
          (rgb, depth) stay strict on ALL rays; per-sample fine outputs are compared on the rays whose 96 depths agree to 1e-4
          (reported, must be > 90 %) at 2e-3."""
    import satnerf_b200 as sb
    R = 8192
    args = make_args(n_importance=32, precision="tc")
    ms, params = _models(args, 43)
    rays, ts = orc.synthetic_sat_rays(R, seed=44)
    draws = _draws(R, 64, N=32, seed=45)
    want = orc.render_rays(params, args, rays, ts, orc.Draws(draws))
    with torch.no_grad():
        got = sb.render_rays(ms, args, rays.cuda(), ts.cuda(), _draws=draws)
    assert set(got) == set(want) and got["weights_fine"].shape == (R, 96)
    _compare("config3_coarse", got, want, keys=[k for k in want if k.endswith("_coarse")])
    # (2) the oracle's merged depths
    zc = orc.stratified_depths(rays[:, 6:7], rays[:, 7:8], 64, draws[0])
    mid = 0.5 * (zc[:, :-1] + zc[:, 1:])
    zf_ref = torch.sort(torch.cat([zc, orc.importance_depths(mid, want["weights_coarse"][:, 1:-1], draws[2])], -1), -1)[0]
    xyz = rays[:, None, 0:3] + rays[:, None, 3:6] * zf_ref[:, :, None]
    with torch.no_grad():
        dec = sb.inference(ms["fine"], args, xyz.cuda().contiguous(), zf_ref.cuda().contiguous(), sun_d=rays[:, 8:11].cuda().contiguous(),
                           rays_t=ms["t"](ts.cuda()))
    _compare("config3_fine_decoupled", {f"{k}_fine": v for k, v in dec.items()}, want, keys=[k for k in want if k.endswith("_fine")])
    # (3) end to end
    from satnerf_b200 import capi
    z = capi.stratified_depths(rays.cuda(), torch.linspace(0, 1, 64).cuda(), draws[0].cuda())
    zf_dev = capi.importance_depths(z, got["weights_coarse"].contiguous(), draws[2].cuda().contiguous()).cpu()
    same = ((zf_dev - zf_ref).abs().max(-1).values <= 1e-4)
    frac = float(same.float().mean())
    fine = [k for k in want if k.endswith("_fine")]
    rows = {"rays_with_equal_fine_depths": frac, "max_fine_depth_diff": float((zf_dev - zf_ref).abs().max())}
    for k in fine:
        a, b = got[k].cpu(), want[k]
        rows[k] = {"elementwise_excess_equal_depth_rays_rtol_2e-3": elementwise_excess(a[same], b[same], k, rtol=2e-3), "max_rel_all_rays": rel_err(a, b)}
    _report("config3_fine_end_to_end", rows)
    assert frac > 0.9, frac
    for k in fine:
        if k == "weights_fine":
            # w_i = alpha_i T_i with alpha_i ~ sigma_i * (z_{i+1} - z_i): two merged samples can lie arbitrarily close, so a 1e-5 shift of
            # one of them changes that weight by percents of ITSELF (its cumulative sum -- transparency, gated above -- is unaffected);
            # gate it on the max-abs / max-ref measure over all rays instead
            assert rows[k]["max_rel_all_rays"] < 2e-3, rows[k]
            continue
        assert rows[k]["elementwise_excess_equal_depth_rays_rtol_2e-3"] <= 1.0, (k, rows[k])
    for k in ("rgb_fine", "depth_fine"):
        assert elementwise_excess(got[k].cpu(), want[k], k) <= 1.0, (k, rows[k])


def test_config4_snerf_sc_tc_vs_oracle():
    """configs[3]: s-nerf + solar-correction pass (rendering.py:90-96), 4096 rays x 64 samples."""
    import satnerf_b200 as sb
    args = make_args(model="s-nerf", sc_lambda=0.05, noise_std=0.05, precision="tc")
    ms, params = _models(args, 46)
    rays, _ = orc.synthetic_sat_rays(4096, seed=47)
    draws = _draws(4096, 64, sc=True, seed=48)
    want = orc.render_rays(params, args, rays, None, orc.Draws(draws))
    with torch.no_grad():
        got = sb.render_rays(ms, args, rays.cuda(), None, _draws=draws)
    assert set(got) == set(want)
    _compare("config4_snerf_sc_tc_vs_oracle", got, want)


def test_config5_dsm_batch_tc_vs_oracle():
    """configs[4]: one 65 536-ray batch of the DSM extraction (create_satnerf_dsm.py:78), rays on a regular 256x256 grid."""
    import satnerf_b200 as sb
    R = 65536
    args = make_args(precision="tc")
    ms, params = _models(args, 49)
    rays, ts = orc.synthetic_sat_rays(R, n_images=1, seed=50)
    gy, gx = torch.meshgrid(torch.linspace(-1, 1, 256), torch.linspace(-1, 1, 256), indexing="ij")
    rays[:, 0], rays[:, 1] = gx.reshape(-1), gy.reshape(-1)
    draws = _draws(R, 64, seed=51)
    want = orc.render_rays(params, args, rays, ts, orc.Draws(draws))
    with torch.no_grad():
        got = sb.render_rays(ms, args, rays.cuda(), ts.cuda(), _draws=draws)
    _compare("config5_dsm_batch_tc_vs_oracle", got, want)


@pytest.mark.parametrize("name", TRAINED_CASES)
def test_trained_like_golden(name):
    """Reference fixture in the trained-like regime: fp32 path at 1e-3 elementwise; tensor-core path measured and gated at
    the documented bound (2e-2 max-abs/max-ref: the sharp density head multiplies the fp16-operand error of the trunk)."""
    g = Golden(name)
    w = g.out["weights_coarse"]
    assert float(w.argmax(-1).float().mean()) < 50 and float(w[:, -1].mean()) < 0.5          # rays do saturate mid-way
    _, _, res = run_golden(g, "fp32")
    _compare(f"{name}_fp32", res, g.out)
    _, _, res = run_golden(g, "tcx3")             # tensor cores, fp16 hi+lo operands: the same strict gate as the fp32 path
    _compare(f"{name}_tcx3", res, g.out)
    _, _, res = run_golden(g, "tc")
    rows = {k: {"elementwise_excess": elementwise_excess(res[k].cpu(), g.out[k], k), "max_rel": rel_err(res[k].cpu(), g.out[k])} for k in g.out}
    _report(f"{name}_tc", rows)
    for k, v in rows.items():
        assert v["max_rel"] < 2e-2, (k, v)


def test_trained_like_regime_fullsize():
    """4096 rays x 64 samples, h=512, trained-like weights, both device paths against the oracle (report + gates as above)."""
    import satnerf_b200 as sb
    args = make_args(precision="fp32")
    ms, _ = _models(args, 52)
    sd = {k: torch.from_numpy(v) for k, v in trained_like(pcg_params(field_shapes("sat-nerf", 512), 777)).items()}
    ms["coarse"].load_state_dict(sd)
    params = {"coarse": sd, "t": ms["t"].weight.detach().cpu().clone()}
    rays, ts = orc.synthetic_sat_rays(4096, seed=53)
    draws = _draws(4096, 64, seed=54)
    want = orc.render_rays(params, args, rays, ts, orc.Draws(draws))
    with torch.no_grad():
        got32 = sb.render_rays(ms, args, rays.cuda(), ts.cuda(), _draws=draws)
        args.precision = "tcx3"
        gotx3 = sb.render_rays(ms, args, rays.cuda(), ts.cuda(), _draws=draws)
        args.precision = "tc"
        got16 = sb.render_rays(ms, args, rays.cuda(), ts.cuda(), _draws=draws)
    _compare("trained_like_fullsize_fp32", got32, want)
    _compare("trained_like_fullsize_tcx3", gotx3, want)
    rows = {k: {"elementwise_excess": elementwise_excess(got16[k].cpu(), want[k], k), "max_rel": rel_err(got16[k].cpu(), want[k])} for k in want}
    _report("trained_like_fullsize_tc", rows)
    for k, v in rows.items():
        assert v["max_rel"] < 2e-2, (k, v)
