"""Losses consumed by the training step — mirror of the reference's metrics.py:8-107 (ssim/kornia left out:
it is an evaluation metric outside the render hot path).  They define which render outputs carry gradient
(rgb, weights, beta, depth, sun_sc).  Small elementwise work on (R,3)/(R,S) tensors."""
import torch


def _mse(a, b):
    return ((a - b) ** 2).mean()


def uncertainty_aware_loss(loss_dict, inputs, gt_rgb, typ, beta_min=0.05):      # metrics.py:21-25
    beta = torch.sum(inputs[f"weights_{typ}"].unsqueeze(-1) * inputs["beta_coarse"], -2) + beta_min
    loss_dict[f"{typ}_color"] = ((inputs[f"rgb_{typ}"] - gt_rgb) ** 2 / (2 * beta ** 2)).mean()
    loss_dict[f"{typ}_logbeta"] = (3 + torch.log(beta).mean()) / 2
    return loss_dict


def solar_correction(loss_dict, inputs, typ, lambda_sc=0.05):                  # metrics.py:27-34
    sun_sc = inputs[f"sun_sc_{typ}"].squeeze()
    term2 = torch.sum(torch.square(inputs[f"transparency_sc_{typ}"].detach() - sun_sc), -1)
    term3 = 1 - torch.sum(inputs[f"weights_sc_{typ}"].detach() * sun_sc, -1)
    loss_dict[f"{typ}_sc_term2"] = lambda_sc / 3.0 * torch.mean(term2)
    loss_dict[f"{typ}_sc_term3"] = lambda_sc / 3.0 * torch.mean(term3)
    return loss_dict


class NerfLoss(torch.nn.Module):                                               # metrics.py:8-19
    def forward(self, inputs, targets):
        d = {"coarse_color": _mse(inputs["rgb_coarse"], targets)}
        if "rgb_fine" in inputs:
            d["fine_color"] = _mse(inputs["rgb_fine"], targets)
        return sum(d.values()), d


class SNerfLoss(torch.nn.Module):                                              # metrics.py:36-55
    def __init__(self, lambda_sc=0.05):
        super().__init__()
        self.lambda_sc = lambda_sc

    def forward(self, inputs, targets):
        d = {}
        for typ in ("coarse", "fine"):
            if f"rgb_{typ}" not in inputs:
                continue
            d[f"{typ}_color"] = _mse(inputs[f"rgb_{typ}"], targets)
            if self.lambda_sc > 0:
                solar_correction(d, inputs, typ, self.lambda_sc)
        return sum(d.values()), d


class SatNerfLoss(torch.nn.Module):                                            # metrics.py:57-73
    def __init__(self, lambda_sc=0.0):
        super().__init__()
        self.lambda_sc = lambda_sc

    def forward(self, inputs, targets):
        d = {}
        for typ in ("coarse", "fine"):
            if f"rgb_{typ}" not in inputs:
                continue
            uncertainty_aware_loss(d, inputs, targets, typ)
            if self.lambda_sc > 0:
                solar_correction(d, inputs, typ, self.lambda_sc)
        return sum(d.values()), d


class DepthLoss(torch.nn.Module):                                              # metrics.py:75-92
    def __init__(self, lambda_ds=1.0):
        super().__init__()
        self.lambda_ds = lambda_ds / 3.0

    def forward(self, inputs, targets, weights=1.0):
        d = {}
        for typ in ("coarse", "fine"):
            if f"depth_{typ}" in inputs:
                d[f"{typ}_ds"] = self.lambda_ds * torch.mean(weights * (inputs[f"depth_{typ}"] - targets) ** 2)
        return sum(d.values()), d


def load_loss(args):                                                           # metrics.py:94-103
    if args.model == "nerf":
        return NerfLoss()
    if args.model == "s-nerf":
        return SNerfLoss(lambda_sc=args.sc_lambda)
    if args.model == "sat-nerf":
        return SatNerfLoss(lambda_sc=args.sc_lambda)
    raise ValueError(f"model {args.model} is not valid")


def mse(image_pred, image_gt, valid_mask=None, reduction="mean"):              # metrics.py:105-111
    value = (image_pred - image_gt) ** 2
    if valid_mask is not None:
        value = value[valid_mask]
    return torch.mean(value) if reduction == "mean" else value


def psnr(image_pred, image_gt, valid_mask=None, reduction="mean"):             # metrics.py:113-114
    return -10 * torch.log10(mse(image_pred, image_gt, valid_mask, reduction))
