"""Training harness the hot path drops into — a plain-PyTorch mirror of `NeRF_pl` (main.py:26-154).
pytorch_lightning is not installed in this image, so the LightningModule's structure and
method names are kept (define_models, forward, configure_optimizers, training_step) and the Trainer loop
is replaced by explicit calls; with N>1 ranks the gradient all-reduce of satnerf_b200.dist is inserted
between backward and the optimizer step."""
from __future__ import annotations

import copy
import os
from collections import defaultdict

import numpy as np
import torch

from . import capi
from . import dist as sdist
from . import metrics
from .models import load_model
from .rendering import render_loss_backward, render_rays


class FlatAdam(torch.optim.Optimizer):
    """torch.optim.Adam(lr, betas, eps, weight_decay) arithmetic (amsgrad off) with the update of each parameter tensor as ONE
    launch of the library's `snb_adam_step` -- the fields hand their whole flat buffer over as a single parameter, and torch's
    fused multi-tensor Adam runs a single large tensor on ~40 CTAs (84 us for 2.6 M parameters against ~12 us here).
    State keys are torch's (`step`, `exp_avg`, `exp_avg_sq`), so optimizer checkpoints interchange with torch.optim.Adam."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, shards=None, nvls=True):
        """shards: {id(param): dist.SymmetricBuffer} (dist.make_symmetric) switches to the data-parallel step fused with its
        collective: every rank reduces and updates its shard of each buffer through peer memory and delivers the new parameters to
        all ranks (snb_adam_step_sharded) -- no all-reduce; each rank keeps the Adam moments of its own shard only."""
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.shards = shards
        self.nvls = nvls            # reduce / deliver through the NVSwitch multicast mappings when the buffers have them

    @property
    def sharded(self) -> bool:
        return bool(self.shards)

    def _step_sharded(self):
        rank, _ = sdist.world()
        first = next(iter(self.shards.values()))
        first.barrier()                                  # every rank's gradients are final
        for group in self.param_groups:                  # ONE launch per parameter group: all its flat buffers (fields, embedding)
            b1, b2 = group["betas"]
            bufs, dev = [], None
            for p in group["params"]:
                sb = self.shards[id(p)]
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros(sb.n_pad, device=p.device, dtype=torch.float32)
                    st["exp_avg_sq"] = torch.zeros(sb.n_pad, device=p.device, dtype=torch.float32)
                st["step"] = int(st["step"]) + 1
                bufs.append(dict(param_ptrs=sb.param_ptrs, grad_ptrs=sb.grad_ptrs, mc_params=sb.mc_params if self.nvls else 0,
                                 mc_grads=sb.mc_grads if self.nvls else 0, exp_avg=st["exp_avg"], exp_avg_sq=st["exp_avg_sq"], n=sb.n_pad, step=st["step"]))
                dev = p.device
                torch.autograd.graph.increment_version(p)
            for i in range(0, len(bufs), 4):
                capi.adam_step_sharded_multi(bufs[i:i + 4], rank, float(group["lr"]), b1, b2, group["eps"], group["weight_decay"], dev)
        first.barrier()                                  # every rank's parameters have been delivered (and nobody reads our gradients any more)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if self.shards:
            self._step_sharded()
            return loss
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                st["step"] = int(st["step"]) + 1
                capi.adam_step(p.data.view(-1), p.grad.contiguous().view(-1), st["exp_avg"].view(-1), st["exp_avg_sq"].view(-1),
                               float(group["lr"]), b1, b2, group["eps"], group["weight_decay"], st["step"])
                torch.autograd.graph.increment_version(p)       # the write happened outside autograd: caches keyed on ._version must see it
        return loss


class NeRFSystem:
    def __init__(self, args, device="cuda", train_len=None):
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
        # epoch arithmetic of the reference (train_utils.py:14-15): epoch = step // (len(train_dataset) // batch_size).
        # `train_len` = number of training rays; set_train_loaders() takes it from the sampler.
        self.train_len = train_len
        self.batches_per_epoch = None          # len(loader): where Lightning steps the 'epoch'-interval scheduler
        self.fused = bool(getattr(args, "fused_loss", True))
        self.optimizer = self.scheduler = None
        sdist.broadcast_parameters(self.models)      # replicas start from rank 0's initialisation (no-op single-process)

    def define_models(self):                                                   # main.py:49-58
        a = self.args
        self.models = {"coarse": load_model(a).to(self.device)}
        self.nerf_coarse = self.models["coarse"]
        if a.n_importance > 0:
            self.nerf_fine = self.models["fine"] = load_model(a).to(self.device)
        if a.model == "sat-nerf":
            self.embedding_t = self.models["t"] = torch.nn.Embedding(a.t_embbeding_vocab, a.t_embbeding_tau).to(self.device)

    def set_train_loaders(self, loaders):
        """The dict of loaders of train_dataloader (main.py:96-110): fixes the epoch arithmetic (dataset length, batches per epoch)."""
        self.train_len = loaders["color"].n
        self.batches_per_epoch = max(len(v) for v in loaders.values())

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
        """Adam(lr, weight_decay=0) + StepLR(gamma=0.9) per epoch, as the reference.  The fields keep their parameters as views
        of one flat buffer each, so Adam runs on that buffer (one multi-tensor launch over 1-3 tensors instead of 35-70); the
        elementwise update -- and therefore every parameter value -- is identical to per-tensor Adam."""
        groups = []
        for m in self.models.values():
            if hasattr(m, "flat_parameter"):
                groups.append(m.flat_parameter())
            else:
                groups += list(m.parameters())
        if groups[0].is_cuda:
            shards = None
            if sdist.world()[1] > 1 and self.fused and getattr(self.args, "sharded_adam", True) and torch.distributed.get_backend() == "nccl":
                # data-parallel: parameters and gradients move to symmetric memory; the step reduces / updates / delivers shard-wise
                # through NVLink peer memory in one kernel per buffer instead of all-reduce + Adam
                shards = sdist.make_symmetric(self.models)
                groups = []
                for m in self.models.values():
                    groups += [m.flat_parameter()] if hasattr(m, "flat_parameter") else list(m.parameters())
            self.optimizer = FlatAdam(groups, lr=self.args.lr, weight_decay=0, shards=shards,
                                      nvls=bool(getattr(self.args, "sharded_nvls", os.environ.get("SNB_SHARDED_NVLS", "1") != "0")))      # one library launch per step
        else:
            self.optimizer = torch.optim.Adam(groups, lr=self.args.lr, weight_decay=0)
        self.scheduler = torch.optim.lr_scheduler.StepLR(self.optimizer, step_size=1, gamma=0.9)   # stepped per epoch
        return self.optimizer

    def get_current_epoch(self, tstep):                                        # main.py:230-231
        if self.train_len is None:
            if self.use_ts:
                raise RuntimeError("NeRFSystem needs the training-set length (train_len= / set_train_loaders()): sat-nerf switches from "
                                   "SNerfLoss to SatNerfLoss after epoch 2 (main.py:128) and the learning rate decays per epoch")
            return 0
        per_epoch = max(self.train_len // self.args.batch_size, 1)
        return int(tstep // per_epoch)

    def training_step(self, batch):                                            # main.py:119-154
        if self.fused:
            return self.fused_training_step(batch)
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

    def fused_training_step(self, batch):
        """training_step with the losses and their gradients evaluated inside the library (rendering.render_loss_backward):
        same loss terms, same parameter gradients (accumulated straight into the flat `.grad` buffers -- no autograd graph, no
        `loss.backward()`), same RNG consumption.  With N ranks every mean() divides by the GLOBAL batch size, so the gradient
        all-reduce is a plain sum."""
        self.train_steps += 1
        a = self.args
        world = sdist.world()[1]
        chunk = a.chunk
        kind = "mse" if (a.model != "sat-nerf" or self.get_current_epoch(self.train_steps) < 2) else "beta"
        rays, rgbs = batch["color"]["rays"], batch["color"]["rgbs"]
        ts = batch["color"]["ts"].reshape(-1) if self.use_ts else None
        n_mean = rays.shape[0] * world
        loss_dict, rgb_parts = {}, []

        def run(rays_, ts_, backward, **what):
            out = {}
            for i in range(0, rays_.shape[0], chunk):                          # main.py:60-75: ray chunks; terms of the chunks add up
                sl = slice(i, i + chunk)
                spec = {k: tuple(v_[sl] if torch.is_tensor(v_) else v_ for v_ in v) for k, v in what.items()}
                d, res = render_loss_backward(self.models, a, rays_[sl], None if ts_ is None else ts_[sl], n_rays_mean=rays_.shape[0] * world,
                                              backward=backward, **spec)
                for k, v in d.items():
                    out[k] = v if k not in out else out[k] + v
                rgb_parts.append(res)
            return out

        loss_dict.update(run(rays, ts, True, color=(kind, rgbs)))
        typ = "fine" if a.n_importance > 0 else "coarse"
        colour_mse = loss_dict.get(f"{typ}_color") if (kind == "mse" and world == 1) else None      # = mse(rgb, target) of this batch
        rgb = rgb_parts[0][f"rgb_{typ}"] if len(rgb_parts) == 1 else torch.cat([r[f"rgb_{typ}"] for r in rgb_parts], 0)
        loss = sum(loss_dict.values())
        a.noise_std *= 0.9
        if self.depth:
            dep = batch["depth"]
            kp_depths = torch.flatten(dep["depths"][:, 0])
            kp_weights = None if a.ds_noweights else torch.flatten(dep["depths"][:, 1])
            use = self.train_steps < self.ds_drop          # (afterwards the reference still renders the batch, for logging only)
            tmp = run(dep["rays"], dep["ts"].reshape(-1), use, depth=(kp_depths, kp_weights, a.ds_lambda))
            if use:
                loss = loss + sum(tmp.values())
            loss_dict.update(tmp)
        # metrics.py:113-114: psnr = -10 log10(mse(rgb, target)); with the plain colour loss that mse is the term computed above
        loss_dict["psnr"] = -10.0 * torch.log10(colour_mse) if colour_mse is not None else metrics.psnr(rgb, rgbs)
        del n_mean
        return loss, loss_dict

    def zero_grad(self):
        for m in self.models.values():
            if hasattr(m, "flat_grads"):
                m.flat_grads(zero=True)            # one memset; the parameters' .grad (and the flat parameter's) alias it
            else:
                for p in m.parameters():
                    if p.grad is not None:
                        p.grad.zero_()

    def optimization_step(self, batch):
        """zero grads -> training_step -> backward -> [all-reduce] -> Adam -> [epoch end: StepLR] (Lightning's fit loop)."""
        if self.fused:
            self.zero_grad()
            loss, info = self.training_step(batch)
            if not getattr(self.optimizer, "sharded", False):            # (the sharded step reduces the gradients itself, through peer memory)
                sdist.all_reduce_gradients(self.models, average=False)   # 1/world is already in the loss seed (n_rays_mean)
        else:
            for m in self.models.values():      # grads are (re)assigned as slices of one flat buffer by the render backward
                for p in m.parameters():
                    p.grad = None
            loss, info = self.training_step(batch)
            loss.backward()
            for m in self.models.values():
                if hasattr(m, "flat_grads"):
                    m.flat_grads(zero=False)    # (re)bind .grad to the flat buffer the backward filled
            sdist.all_reduce_gradients(self.models, average=True)
        self.optimizer.step()
        if self.batches_per_epoch and self.train_steps % self.batches_per_epoch == 0:
            self.on_epoch_end()
        return loss.detach(), info

    def on_epoch_end(self):
        """Lightning steps an interval='epoch' scheduler here (main.py:88-93): lr *= 0.9."""
        if self.scheduler is not None:
            self.scheduler.step()


def bench_training_step(args, dev, rank, world, n_rays, warm, steps, flush, depth_batch=False, n_dataset=1 << 18):
    """bench.py's training legs: n_rays per rank and step, fed by DeviceRaySampler from a GPU-resident synthetic training set
    (n_dataset rays; every rank keeps the set and takes its shard of each global batch): render forward + in-kernel loss seed
    + backward + one gradient all-reduce + Adam per step.  depth_batch adds the depth-supervision batch (main.py:134-142)."""
    import torch.distributed as dist
    from . import capi
    from .data import DeviceRaySampler, combined_loader
    from .synth import synthetic_sat_rays
    a = copy.copy(args)
    a.lr, a.chunk, a.batch_size = 5e-4, 1 << 20, n_rays * world
    torch.manual_seed(0)
    rays, ts = synthetic_sat_rays(n_dataset, seed=200)
    g = torch.Generator().manual_seed(201)
    data = {"rays": rays, "rgbs": torch.rand(n_dataset, 3, generator=g), "ts": ts.reshape(-1, 1)}
    loaders = {"color": DeviceRaySampler(data, a.batch_size, device=dev, generator=torch.Generator().manual_seed(7), rank=rank, world=world)}
    if depth_batch:
        nd = n_dataset // 8
        dd = {"rays": rays[:nd], "ts": ts[:nd].reshape(-1, 1),
              "depths": torch.stack([0.1 + 0.2 * torch.rand(nd, generator=g), torch.rand(nd, generator=g)], -1)}
        loaders["depth"] = DeviceRaySampler(dd, a.batch_size, device=dev, generator=torch.Generator().manual_seed(8), rank=rank, world=world)
    system = NeRFSystem(a, dev)
    system.set_train_loaders(loaders)
    system.configure_optimizers()
    batches = combined_loader(loaders)

    def next_batch():
        nonlocal batches
        try:
            return next(batches)
        except StopIteration:
            batches = combined_loader(loaders)
            return next(batches)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warm):
        system.optimization_step(next_batch())
    sync()
    capi.launch_count(reset=True)
    evs = []
    for _ in range(steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); system.optimization_step(next_batch()); e1.record()      # the sampler's gather is inside the timed region
        evs.append((e0, e1))
    sync()
    ms = sum(x.elapsed_time(y) for x, y in evs)
    launches = capi.launch_count(reset=True)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    what = "DeviceRaySampler batch + render forward + loss seeded in the compositing backward + backward + " + (
        ("gradient reduction, Adam and parameter delivery fused in one kernel over symmetric memory (snb_adam_step_sharded_multi: " +
         ("in-switch reduction + multicast delivery, multimem.ld_reduce / multimem.st)" if (system.optimizer.nvls and all(sb.mc_grads and sb.mc_params for sb in system.optimizer.shards.values()))
          else "NVLink peer loads / stores)")) if getattr(system.optimizer, "sharded", False)
        else "one flat-gradient all-reduce + Adam")
    if depth_batch:
        what += " (colour batch + depth-supervision batch, both coarse + fine)"
    return {"value": world * n_rays * steps / (ms * 1e-3), "unit": "rays/s", "ms_per_step": ms / steps, "rays_per_gpu": n_rays,
            "model": a.model, "what": what, "gpu_launches": launches / steps}
