"""Training harness the hot path drops into — a plain-PyTorch mirror of `NeRF_pl` (main.py:26-154).
pytorch_lightning is not installed in this image, so the LightningModule's structure and
method names are kept (define_models, forward, configure_optimizers, training_step) and the Trainer loop
is replaced by explicit calls; with N>1 ranks the gradient all-reduce of satnerf_b200.dist is inserted
between backward and the optimizer step."""
from __future__ import annotations

import copy
from collections import defaultdict

import numpy as np
import torch

from . import dist as sdist
from . import metrics
from .models import load_model
from .rendering import render_rays


class NeRFSystem:
    def __init__(self, args, device="cuda"):
        self.args = copy.copy(args)
        self.device = torch.device(device)
        self.loss = metrics.load_loss(args)                                    # main.py:33
        self.depth = getattr(args, "ds_lambda", 0.0) > 0                       # main.py:34-38
        if self.depth:
            self.depth_loss = metrics.DepthLoss(lambda_ds=args.ds_lambda)
            self.ds_drop = np.round(args.ds_drop * args.max_train_steps)
        self.define_models()
        self.train_steps = 0
        self.use_ts = args.model == "sat-nerf"                                 # main.py:44-47
        if self.use_ts:
            self.loss_without_beta = metrics.SNerfLoss(lambda_sc=args.sc_lambda)
        self.steps_per_epoch = getattr(args, "steps_per_epoch", None)

    def define_models(self):                                                   # main.py:49-58
        a = self.args
        self.models = {"coarse": load_model(a).to(self.device)}
        self.nerf_coarse = self.models["coarse"]
        if a.n_importance > 0:
            self.nerf_fine = self.models["fine"] = load_model(a).to(self.device)
        if a.model == "sat-nerf":
            self.embedding_t = self.models["t"] = torch.nn.Embedding(a.t_embbeding_vocab, a.t_embbeding_tau).to(self.device)

    def state_dict(self):
        """Lightning-style keys (`nerf_coarse.*`, `nerf_fine.*`, `embedding_t.*`) read by eval_satnerf.py:23-44."""
        names = {"coarse": "nerf_coarse", "fine": "nerf_fine", "t": "embedding_t"}
        return {f"{names[k]}.{n}": v for k, m in self.models.items() for n, v in m.state_dict().items()}

    def forward(self, rays, ts):                                               # main.py:60-75
        chunk = self.args.chunk
        results = defaultdict(list)
        for i in range(0, rays.shape[0], chunk):
            out = render_rays(self.models, self.args, rays[i:i + chunk], None if ts is None else ts[i:i + chunk])
            for k, v in out.items():
                results[k].append(v)
        return {k: (v[0] if len(v) == 1 else torch.cat(v, 0)) for k, v in results.items()}

    __call__ = forward

    def configure_optimizers(self):                                            # main.py:81-94, train_utils.py:24-53
        params = [p for m in self.models.values() for p in m.parameters()]
        # same update rule as the reference (Adam, lr, no weight decay); the fused implementation is one launch on CUDA
        self.optimizer = torch.optim.Adam(params, lr=self.args.lr, weight_decay=0, fused=params[0].is_cuda)
        self.scheduler = torch.optim.lr_scheduler.StepLR(self.optimizer, step_size=1, gamma=0.9)   # stepped per epoch
        return self.optimizer

    def get_current_epoch(self, tstep):
        return 0 if not self.steps_per_epoch else tstep // self.steps_per_epoch

    def training_step(self, batch):                                            # main.py:119-154
        self.train_steps += 1
        rays, rgbs = batch["color"]["rays"], batch["color"]["rgbs"]
        ts = batch["color"]["ts"].squeeze() if self.use_ts else None
        results = self(rays, ts)
        if "beta_coarse" in results and self.get_current_epoch(self.train_steps) < 2:
            loss, loss_dict = self.loss_without_beta(results, rgbs)
        else:
            loss, loss_dict = self.loss(results, rgbs)
        self.args.noise_std *= 0.9
        if self.depth:
            tmp = self(batch["depth"]["rays"], batch["depth"]["ts"].squeeze())
            kp_depths = torch.flatten(batch["depth"]["depths"][:, 0])
            kp_weights = 1.0 if self.args.ds_noweights else torch.flatten(batch["depth"]["depths"][:, 1])
            loss_depth, tmp = self.depth_loss(tmp, kp_depths, kp_weights)
            if self.train_steps < self.ds_drop:
                loss = loss + loss_depth
            loss_dict.update(tmp)
        with torch.no_grad():
            typ = "fine" if "rgb_fine" in results else "coarse"
            loss_dict["psnr"] = metrics.psnr(results[f"rgb_{typ}"], rgbs)
        return loss, loss_dict

    def optimization_step(self, batch):
        """zero grads -> training_step -> backward -> [all-reduce] -> Adam (what Lightning's fit loop does)."""
        for m in self.models.values():          # grads are (re)assigned as slices of one flat buffer by the render backward
            for p in m.parameters():
                p.grad = None
        loss, info = self.training_step(batch)
        loss.backward()
        sdist.all_reduce_gradients(self.models)
        self.optimizer.step()
        return loss.detach(), info


def bench_training_step(args, dev, rank, world, n_rays, warm, steps, flush):
    """bench.py's training leg: n_rays per rank, fwd + bwd + gradient all-reduce + Adam per step."""
    import torch.distributed as dist
    from . import capi
    a = copy.copy(args)
    a.lr, a.chunk = 5e-4, 1 << 20
    torch.manual_seed(0)
    system = NeRFSystem(a, dev)
    system.steps_per_epoch = 10 ** 9          # stay in the first epochs (SNerfLoss branch, main.py:128-129)
    system.configure_optimizers()
    g = torch.Generator().manual_seed(200 + rank)
    u = torch.rand(n_rays, 2, generator=g) * 2 - 1
    o = torch.cat([u, torch.ones(n_rays, 1)], -1)
    d = torch.tensor([[0.3, 0.1, -0.95]]).expand(n_rays, 3) + 1e-3 * torch.randn(n_rays, 3, generator=g)
    d = d / d.norm(dim=-1, keepdim=True)
    sun = torch.tensor([[0.4, -0.5, 0.77]]).expand(n_rays, 3)
    rays = torch.cat([o, d, torch.zeros(n_rays, 1), 0.3 + 0.3 * torch.rand(n_rays, 1, generator=g), sun], -1).to(dev)
    batch = {"color": {"rays": rays, "rgbs": torch.rand(n_rays, 3, generator=g).to(dev),
                       "ts": torch.randint(0, 17, (n_rays, 1), generator=g).to(dev)}}

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warm):
        system.optimization_step(batch)
    sync()
    capi.launch_count(reset=True)
    evs = []
    for _ in range(steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); system.optimization_step(batch); e1.record()
        evs.append((e0, e1))
    sync()
    ms = sum(x.elapsed_time(y) for x, y in evs)
    launches = capi.launch_count(reset=True)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return {"value": world * n_rays * steps / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms / steps, "rays_per_gpu": n_rays,
            "what": "render_rays forward + SNerfLoss + backward + flat-gradient all-reduce + Adam", "gpu_launches": launches}
