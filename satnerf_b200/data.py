"""GPU-resident ray sampler — the training-side neighbour of the render path.

The reference feeds `training_step` from a per-ray `DataLoader(train_dataset, shuffle=True, batch_size=B,
num_workers=4, pin_memory=True)` (main.py:96-110) whose dataset answers `__getitem__(idx)` with
`{"rays": all_rays[idx], "rgbs": all_rgbs[idx], "ts": all_ids[idx].long()}` (datasets/satellite.py:347-350;
depth set: `"depths": all_depths[idx]`, datasets/satellite_depth.py:138-141).  That is B Python `__getitem__` calls plus a
collate per step in worker processes — 1e4-1e5 rays/s, far below the fused render kernel.  `DeviceRaySampler` keeps the
same tensors resident on the GPU and yields the same batches — a fresh random permutation every epoch, consecutive slices
of B rays, last batch short (`drop_last=False`) — assembling each batch on the device with one gather launch over all per-ray
tables (`snb_gather_rows`).
"""
from __future__ import annotations

from typing import Dict, Iterator, Optional

import torch

from . import capi
from .dist import shard_bounds


class DeviceRaySampler:
    """Iterable over `{"rays", "ts", "rgbs" | "depths"}` batches with DataLoader(shuffle=True) semantics.

    tensors: the dataset's per-ray arrays (`all_rays (N,11|8)`, `all_ids (N,1)`, `all_rgbs (N,3)` / `all_depths (N,2)`),
    given under the keys the dataset's `__getitem__` uses; `ts` is cast to int64 like `.long()` there."""

    def __init__(self, tensors: Dict[str, torch.Tensor], batch_size: int, device="cuda", shuffle: bool = True,
                 generator: Optional[torch.Generator] = None, rank: int = 0, world: int = 1):
        n = {int(v.shape[0]) for v in tensors.values()}
        if len(n) != 1:
            raise ValueError(f"per-ray tensors disagree on the number of rays: {sorted(n)}")
        self.n = n.pop()
        self.batch_size, self.shuffle, self.generator = int(batch_size), shuffle, generator
        self.device = torch.device(device)
        self.rank, self.world = rank, world
        self.data = {k: (v.long() if k == "ts" else v.to(torch.float32)).to(self.device).contiguous() for k, v in tensors.items()}

    def __len__(self) -> int:                       # batches per epoch, as len(DataLoader)
        return (self.n + self.batch_size - 1) // self.batch_size

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        if self.shuffle:                            # the permutation is drawn on the host generator (like RandomSampler) so that every
            perm = torch.randperm(self.n, generator=self.generator).to(self.device)      # rank of a job draws the same one
        else:
            perm = torch.arange(self.n, device=self.device)
        for b in range(len(self)):
            idx = perm[b * self.batch_size:(b + 1) * self.batch_size]
            if self.world > 1:                      # contiguous ray shard of the global batch per rank: the split of dist.shard_bounds
                lo, hi = shard_bounds(idx.numel(), self.rank, self.world)
                idx = idx[lo:hi]
            if self.device.type == "cuda" and len(self.data) <= 4:
                yield capi.gather_rows(self.data, idx)       # one launch for all per-ray tables (snb_gather_rows)
            else:
                yield {k: v[idx] for k, v in self.data.items()}


def combined_loader(loaders: Dict[str, DeviceRaySampler]) -> Iterator[Dict[str, Dict[str, torch.Tensor]]]:
    """The dict of loaders of `train_dataloader` (main.py:96-110) as Lightning 1.x consumes it: dict batches
    `{"color": ..., "depth": ...}`, one epoch = the LONGEST loader, shorter ones restart (`max_size_cycle`)."""
    n = max(len(v) for v in loaders.values())
    its = {k: iter(v) for k, v in loaders.items()}
    for _ in range(n):
        batch = {}
        for k in loaders:
            try:
                batch[k] = next(its[k])
            except StopIteration:
                its[k] = iter(loaders[k])
                batch[k] = next(its[k])
        yield batch
