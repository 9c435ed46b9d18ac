"""Multi-GPU plumbing of the render path: one process per GPU, rays sharded by rank (rays are independent),
ONE sum-all-reduce per training step over the flat gradient buffer of each field (+ the tiny embedding
gradient).  The reference is single-GPU; this is new work."""
from typing import Dict, Tuple

import torch
import torch.distributed as dist


def world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n_rays: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous split of n_rays; the first (n_rays % world) ranks take one extra ray."""
    base, rem = divmod(n_rays, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_rays(rays, ts=None, rank=None, world_size=None):
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    lo, hi = shard_bounds(rays.shape[0], rank, world_size)
    return rays[lo:hi], (None if ts is None else ts[lo:hi])


def gather_rays(local: Dict[str, torch.Tensor], n_rays: int) -> Dict[str, torch.Tensor]:
    """Inference (config 5): concatenates the per-rank result dicts along the ray axis on every rank."""
    r, w = world()
    if w == 1:
        return local
    out = {}
    for k, v in local.items():
        sizes = [shard_bounds(n_rays, i, w)[1] - shard_bounds(n_rays, i, w)[0] for i in range(w)]
        bufs = [torch.empty((s, *v.shape[1:]), dtype=v.dtype, device=v.device) for s in sizes]
        dist.all_gather(bufs, v.contiguous())
        out[k] = torch.cat(bufs, 0)
    return out


def all_reduce_gradients(models: Dict[str, torch.nn.Module], average: bool = True) -> int:
    """Sums (averages) gradients across ranks: one collective per field over its flat gradient buffer, one for
    the embedding.  Returns the number of collectives issued."""
    r, w = world()
    if w == 1:
        return 0
    n = 0
    for name, m in models.items():
        if hasattr(m, "flat_grads"):
            g = m.flat_grads(zero=False)
        else:
            ps = [p for p in m.parameters() if p.grad is not None]
            if not ps:
                continue
            g = torch.cat([p.grad.reshape(-1) for p in ps])
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        if average:
            g.div_(w)
        if not hasattr(m, "flat_grads"):
            off = 0
            for p in ps:
                p.grad.copy_(g[off:off + p.numel()].view_as(p.grad)); off += p.numel()
        n += 1
    return n
