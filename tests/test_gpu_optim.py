"""Optimiser step of the training loop (snb_adam_step behind train.FlatAdam) against torch.optim.Adam -- what the reference
builds in main.py:81-94 / train_utils.py:24-53.  Tolerance: 2e-6 absolute on parameters of O(1) after 6 steps (fp32
arithmetic in a different association order)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,wd", [(1000003, 0.0), (4096, 0.01), (7, 0.0)])
def test_flat_adam_matches_torch_adam(n, wd):
    from satnerf_b200.train import FlatAdam
    g = torch.Generator().manual_seed(n)
    w0 = torch.randn(n, generator=g).cuda()
    a = torch.nn.Parameter(w0.clone()); b = torch.nn.Parameter(w0.clone())
    oa = FlatAdam([a], lr=5e-4, weight_decay=wd); ob = torch.optim.Adam([b], lr=5e-4, weight_decay=wd)
    sched = torch.optim.lr_scheduler.StepLR(oa, step_size=1, gamma=0.9)
    schedb = torch.optim.lr_scheduler.StepLR(ob, step_size=1, gamma=0.9)
    for it in range(6):
        grad = torch.randn(n, generator=g).cuda() * (10.0 ** (it - 3))
        a.grad = grad.clone(); b.grad = grad.clone()
        v0 = a._version
        oa.step(); ob.step()
        assert a._version > v0                       # caches keyed on the version counter see the update
        if it == 2:
            sched.step(); schedb.step()
    assert float((a - b).abs().max()) < 2e-6
    sa, sb_ = oa.state[a], ob.state[b]
    assert int(sa["step"]) == int(sb_["step"]) == 6
    for k in ("exp_avg", "exp_avg_sq"):           # fma vs lerp rounding: 1e-6 of the tensor's scale
        assert float((sa[k] - sb_[k]).abs().max()) <= 1e-6 * float(sb_[k].abs().max()), k


def test_flat_adam_updates_module_views():
    """The field's parameters are views of the flat buffer Adam updates: every module parameter moves, and the packed-weight
    cache of the inference path is invalidated (flat version counter)."""
    import satnerf_b200 as sb
    from satnerf_b200.train import FlatAdam
    from gpu_util import make_args
    torch.manual_seed(3)
    field = sb.load_model(make_args(fc_units=64)).cuda()
    fp = field.flat_parameter()
    before = [p.detach().clone() for p in field.parameters()]
    opt = FlatAdam([fp], lr=1e-2)
    fp.grad.fill_(1.0)
    v0 = field.flat_params()._version
    opt.step()
    assert field.flat_params()._version > v0
    for p, q in zip(field.parameters(), before):
        assert float((p - q).abs().min()) > 1e-3
