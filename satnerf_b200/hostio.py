"""Host-buffer front end of `render_rays`: ray batches come from (pinned) host memory and the result dict goes back to pinned
host memory, with the copy-out of batch i running on a second stream under the render pass of batch i + 1.

This is what the reference's evaluation scripts do serially (`batched_inference` + `.cpu()` on every value,
eval_satnerf.py:46-66, create_satnerf_dsm.py:78-110): at 4096 rays per batch the full result dict is 10.5 MB, whose
device->host copy costs a fifth of the render pass when it is not overlapped.

    pipe = HostPipeline(models, args, device)
    for rays_h, ts_h in batches:            # pinned CPU tensors
        slot = pipe.submit(rays_h, ts_h)    # returns at once; results of `slot` are valid after pipe.wait(slot)
    pipe.wait()                             # all outstanding batches
    pipe.host[slot]["rgb_coarse"] ...       # pinned host tensors (one set per slot, reused round-robin)
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from .rendering import render_rays


class HostPipeline:
    def __init__(self, models, args, device, depth: int = 2, keys: Optional[Sequence[str]] = None):
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.models, self.args, self.device, self.depth = models, args, torch.device(device), depth
        self.keys = None if keys is None else tuple(keys)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.host: List[Optional[Dict[str, torch.Tensor]]] = [None] * depth       # pinned result buffers per slot
        self._live: List[Optional[tuple]] = [None] * depth                        # device tensors the slot's copies still read
        self._done: List[Optional[torch.cuda.Event]] = [None] * depth
        self._next = 0
        self.h2d_bytes = self.d2h_bytes = 0                                       # of the most recent batch

    def submit(self, rays_h: torch.Tensor, ts_h: Optional[torch.Tensor]) -> int:
        slot = self._next
        self._next = (slot + 1) % self.depth
        self.wait(slot)                                   # the slot's previous copy-out has finished: host buffers and device outputs are free
        compute = torch.cuda.current_stream(self.device)
        rays = rays_h.to(self.device, non_blocking=True)
        ts = None if ts_h is None else ts_h.to(self.device, non_blocking=True)
        with torch.no_grad():
            out = render_rays(self.models, self.args, rays, ts)
        keys = self.keys or tuple(out)
        bufs = self.host[slot]
        if bufs is None or any(k not in bufs or bufs[k].shape != out[k].shape for k in keys):
            bufs = {k: torch.empty(out[k].shape, dtype=out[k].dtype).pin_memory() for k in keys}
            self.host[slot] = bufs
        ready = torch.cuda.Event()
        ready.record(compute)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(ready)
            for k in keys:
                bufs[k].copy_(out[k], non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.copy_stream)
        self._live[slot] = (out, rays, ts)               # kept until `done`: the caching allocator must not hand them out earlier
        self._done[slot] = done
        self.h2d_bytes = rays_h.numel() * rays_h.element_size() + (0 if ts_h is None else ts_h.numel() * ts_h.element_size())
        self.d2h_bytes = sum(bufs[k].numel() * bufs[k].element_size() for k in keys)
        return slot

    def wait(self, slot: Optional[int] = None) -> None:
        for s in (range(self.depth) if slot is None else (slot,)):
            if self._done[s] is not None:
                self._done[s].synchronize()
                self._done[s] = None
                self._live[s] = None

    def join(self) -> None:
        """Makes the current stream wait for every outstanding copy-out (device-side; for timing with events)."""
        torch.cuda.current_stream(self.device).wait_stream(self.copy_stream)
