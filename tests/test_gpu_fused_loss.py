"""GPU: the losses of metrics.py evaluated inside the library.

  * `render_loss_backward` (loss gradient seeded in the compositing backward, no autograd graph): loss value and EVERY
    parameter gradient against the golden `loss` / `grad.*` recorded from the unmodified reference (fp32 path 2e-4 of each
    tensor's max |grad|, tensor-core path GRAD_TOL['tc'], both measured against the float64 oracle like test_gpu_parity).
  * the drop-in loss classes (`satnerf_b200.metrics`, thin bindings of snb_loss_forward / snb_loss_backward) on a result dict
    through autograd give the same numbers.
"""
import pytest
import torch

from golden_io import CASES, Golden, rel_err
from gpu_util import models_from_golden
from test_gpu_parity import GRAD_TOL, TOL, _grad_tol, _grads_fp64

pytestmark = pytest.mark.gpu
H64 = [c for c in CASES if "h64" in c]


def _spec(g):
    if g.loss_kind == "depth":
        return dict(depth=(g.depth_target.cuda(), g.depth_weights.cuda(), 1000.0))
    return dict(color=("beta" if g.loss_kind == "satnerf" else "mse", g.target.cuda()))


@pytest.mark.parametrize("precision", ["fp32", "tc"])
@pytest.mark.parametrize("name", H64)
def test_fused_loss_backward_matches_reference(name, precision):
    from satnerf_b200.rendering import render_loss_backward
    g = Golden(name)
    ms, args = models_from_golden(g)
    args.precision = precision
    for m in ms.values():
        if hasattr(m, "flat_grads"):
            m.flat_grads(zero=True)
    ts = None if g.ts is None else g.ts.cuda()
    loss_dict, res = render_loss_backward(ms, args, g.rays.cuda(), ts, _draws=g.draws, **_spec(g))
    loss = float(sum(loss_dict.values()))
    assert abs(loss - g.loss) <= TOL[precision] * 10 * max(1.0, abs(g.loss)), (loss, g.loss, {k: float(v) for k, v in loss_dict.items()})
    assert rel_err(res["rgb_coarse"].cpu(), g.out["rgb_coarse"]) < TOL[precision]
    exact = _grads_fp64(g)
    checked = 0
    for key, ref in g.grads.items():
        lvl, _, pname = key.partition(".")
        got = ms["t"].weight.grad if key == "t" else dict(ms[lvl].named_parameters())[pname].grad
        assert got is not None, key
        ref_noise = rel_err(ref, exact[key], floor=1e-12)
        err = rel_err(got.cpu(), exact[key], floor=1e-12)
        assert err < max(_grad_tol(precision, key), 3 * ref_noise), (key, err, ref_noise)
        checked += 1
    assert checked > 10


@pytest.mark.parametrize("name", H64)
def test_loss_classes_match_reference(name):
    """satnerf_b200.metrics.{SatNerfLoss, SNerfLoss, DepthLoss} on the autograd result dict (fp32 path)."""
    import satnerf_b200 as sb
    from satnerf_b200 import metrics
    g = Golden(name)
    ms, args = models_from_golden(g)
    args.precision = "fp32"
    ts = None if g.ts is None else g.ts.cuda()
    res = sb.render_rays(ms, args, g.rays.cuda(), ts, _draws=g.draws)
    if g.loss_kind == "satnerf":
        loss, d = metrics.SatNerfLoss(lambda_sc=args.sc_lambda)(res, g.target.cuda())
    elif g.loss_kind == "snerf":
        loss, d = metrics.SNerfLoss(lambda_sc=args.sc_lambda)(res, g.target.cuda())
    else:
        loss, d = metrics.DepthLoss(lambda_ds=1000.0)(res, g.depth_target.cuda(), g.depth_weights.cuda())
    assert abs(float(loss.detach()) - g.loss) <= 2e-4 * max(1.0, abs(g.loss)), (float(loss), g.loss)
    loss.backward()
    exact = _grads_fp64(g)
    for key, ref in g.grads.items():
        lvl, _, pname = key.partition(".")
        got = ms["t"].weight.grad if key == "t" else dict(ms[lvl].named_parameters())[pname].grad
        assert got is not None, key
        ref_noise = rel_err(ref, exact[key], floor=1e-12)
        assert rel_err(got.cpu(), exact[key], floor=1e-12) < max(GRAD_TOL["fp32"], 3 * ref_noise), key


def test_fused_step_equals_autograd_step_fullsize():
    """1024 rays x 64 samples, h=512, tensor-core path: NeRFSystem's fused step and its autograd step (loss classes on the
    result dict) produce the same loss and the same flat gradient (same kernels underneath; the compositing backward is seeded
    with upstream tensors in one case and in-kernel in the other)."""
    import argparse

    from satnerf_b200.synth import synthetic_sat_rays
    from satnerf_b200.train import NeRFSystem
    base = dict(model="sat-nerf", n_samples=64, n_importance=0, noise_std=0.0, sc_lambda=0.05, chunk=1 << 20, fc_layers=8, fc_units=512,
                t_embbeding_tau=4, t_embbeding_vocab=30, batch_size=1024, lr=5e-4, precision="tc")
    rays, ts = synthetic_sat_rays(1024, seed=3)
    batch = {"color": {"rays": rays.cuda(), "rgbs": torch.rand(1024, 3, generator=torch.Generator().manual_seed(4)).cuda(), "ts": ts.cuda().reshape(-1, 1)}}
    out = {}
    for fused in (True, False):
        torch.manual_seed(0)
        sysm = NeRFSystem(argparse.Namespace(**base, fused_loss=fused), "cuda", train_len=1024 * 10)
        sysm.train_steps = 25                      # epoch 2: SatNerfLoss (uncertainty term) + solar correction
        torch.manual_seed(1)
        if fused:
            sysm.zero_grad()
            loss, info = sysm.training_step(batch)
        else:
            loss, info = sysm.training_step(batch)
            loss.backward()
        gflat = sysm.models["coarse"].flat_grads(zero=False).clone()
        out[fused] = (float(loss.detach()), {k: float(v.detach()) for k, v in info.items()}, gflat, sysm.models["t"].weight.grad.clone())
    assert set(out[True][1]) == set(out[False][1]) == {"coarse_color", "coarse_logbeta", "coarse_sc_term2", "coarse_sc_term3", "psnr"}
    assert abs(out[True][0] - out[False][0]) <= 1e-5 * max(1.0, abs(out[False][0]))
    assert rel_err(out[True][2], out[False][2], floor=1e-12) < 1e-4
    assert rel_err(out[True][3], out[False][3], floor=1e-12) < 1e-4


def test_epoch_arithmetic_and_scheduler():
    """the loss switches to SatNerfLoss after epoch 2 (main.py:128), the epoch comes from the dataset length
    (train_utils.py:14-15), StepLR is stepped at epoch ends, and a sat-nerf system without the dataset length refuses to guess."""
    import argparse

    from satnerf_b200.data import DeviceRaySampler
    from satnerf_b200.synth import synthetic_sat_rays
    from satnerf_b200.train import NeRFSystem
    a = argparse.Namespace(model="sat-nerf", n_samples=16, n_importance=0, noise_std=0.0, sc_lambda=0.0, chunk=5120, fc_layers=8, fc_units=64,
                           t_embbeding_tau=4, t_embbeding_vocab=30, batch_size=64, lr=5e-4, precision="fp32")
    sysm = NeRFSystem(a, "cuda")
    with pytest.raises(RuntimeError):
        sysm.get_current_epoch(1)
    rays, ts = synthetic_sat_rays(200, seed=1)
    loaders = {"color": DeviceRaySampler({"rays": rays, "rgbs": torch.rand(200, 3), "ts": ts.reshape(-1, 1)}, 64)}
    sysm.set_train_loaders(loaders)
    sysm.configure_optimizers()
    assert sysm.get_current_epoch(5) == 1 and sysm.get_current_epoch(6) == 2            # 200 // 64 = 3 steps per epoch
    keys = []
    for _ in range(3):
        for batch in loaders["color"]:
            _, info = sysm.optimization_step({"color": batch})
            keys.append("coarse_logbeta" in info)
    assert keys[:5] == [False] * 5 and all(keys[5:])                                   # steps 1..5 -> epochs 0,1; from step 6 on epoch >= 2
    assert abs(sysm.optimizer.param_groups[0]["lr"] - 5e-4 * 0.9 ** 3) < 1e-12          # len(loader) = 4 batches per epoch, 12 steps
