"""The fused tensor-core forward runs as CTA pairs (cta_group::2, the default) or as single CTAs (args.tc_cta_group = 1 ->
snb_pass_desc.flags & SNB_PASS_SINGLE_CTA).
Both modes must give bit-identical results (same MMA shapes per point, same epilogue arithmetic), odd group counts
(a pair with an idle half) included, in inference and in training mode (activation stash -> tensor-core backward)."""
import pytest
import torch

from gpu_util import make_args
from oracle import render_oracle as orc

pytestmark = pytest.mark.gpu


def _run(cg, args, ms, rays, ts, draws, target=None):
    import satnerf_b200 as sb
    args.tc_cta_group = cg
    for m in ms.values():
        m.zero_grad(set_to_none=True)
    if target is None:
        with torch.no_grad():
            return sb.render_rays(ms, args, rays, ts, _draws=draws), None
    res = sb.render_rays(ms, args, rays, ts, _draws=draws)
    orc.loss_satnerf(res, target)[0].backward()
    return res, {n: p.grad.clone() for n, p in ms["coarse"].named_parameters()}


@pytest.mark.parametrize("h,n_rays,S,train", [(512, 4096, 64, False), (512, 101, 64, False), (256, 51, 96, False), (384, 7, 48, False),
                                               (512, 101, 64, True), (128, 50, 64, True)])
def test_cta_pair_equals_single_cta(h, n_rays, S, train):
    import satnerf_b200 as sb
    args = make_args(fc_units=h, n_samples=S, precision="tc")
    torch.manual_seed(5)
    ms = {"coarse": sb.load_model(args).cuda(), "t": torch.nn.Embedding(30, 4).cuda()}
    rays, ts = orc.synthetic_sat_rays(n_rays, seed=6)
    g = torch.Generator().manual_seed(7)
    draws = [torch.rand(n_rays, S, generator=g), torch.randn(n_rays, S, generator=g)]
    target = torch.rand(n_rays, 3, generator=g).cuda() if train else None
    a, ga = _run(2, args, ms, rays.cuda(), ts.cuda(), draws, target)
    b, gb = _run(1, args, ms, rays.cuda(), ts.cuda(), draws, target)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    if train:
        for n in ga:
            assert torch.isfinite(ga[n]).all()
            assert torch.equal(ga[n], gb[n]), n


def test_packed_weight_cache_follows_parameter_updates():
    """Inference keeps the packed fp16 weight tiles between calls (snb_pass_desc.weights_packed); an in-place parameter
    update (optimizer step, load_state_dict) must invalidate them."""
    import copy

    import satnerf_b200 as sb
    args = make_args(fc_units=256, n_samples=64, precision="tc")
    torch.manual_seed(11)
    ms = {"coarse": sb.load_model(args).cuda(), "t": torch.nn.Embedding(30, 4).cuda()}
    rays, ts = orc.synthetic_sat_rays(300, seed=12)
    g = torch.Generator().manual_seed(13)
    draws = [torch.rand(300, 64, generator=g), torch.randn(300, 64, generator=g)]
    with torch.no_grad():
        a = sb.render_rays(ms, args, rays.cuda(), ts.cuda(), _draws=draws)
        b = sb.render_rays(ms, args, rays.cuda(), ts.cuda(), _draws=draws)          # served from the cached tiles
        assert all(torch.equal(a[k], b[k]) for k in a)
        for p in ms["coarse"].parameters():
            p.mul_(1.05)                                                             # in-place update: version counter moves
        c = sb.render_rays(ms, args, rays.cuda(), ts.cuda(), _draws=draws)
        fresh = {"coarse": sb.load_model(args).cuda(), "t": ms["t"]}
        fresh["coarse"].load_state_dict(copy.deepcopy(ms["coarse"].state_dict()))
        d = sb.render_rays(fresh, args, rays.cuda(), ts.cuda(), _draws=draws)
    assert not torch.equal(a["rgb_coarse"], c["rgb_coarse"])
    assert all(torch.equal(c[k], d[k]) for k in c)
