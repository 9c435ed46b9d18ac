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


def broadcast_parameters(models: Dict[str, torch.nn.Module], src: int = 0) -> int:
    """Start-up: every rank takes rank `src`'s parameters (one broadcast per field over its flat buffer, one for the embedding), so
    replicas agree whatever their local seeds were.  Writes go through `.data`, which autograd's version counters do not see:
    the cached fp16 weight tiles are invalidated explicitly.  Returns the number of collectives."""
    r, w = world()
    if w == 1:
        return 0
    n = 0
    for m in models.values():
        if hasattr(m, "flat_params"):
            dist.broadcast(m.flat_params(), src)
            m.invalidate_packed()
        else:
            for p in m.parameters():
                dist.broadcast(p.data, src)
        n += 1
    return n


def all_reduce_gradients(models: Dict[str, torch.nn.Module], average: bool = True) -> int:
    """Sums gradients across ranks with ONE collective launch per step: when every gradient already lives in one bucket (a single
    field and no embedding) the flat buffer is reduced in place; with several buffers NCCL reduces them in place inside one group
    launch (other backends: packed into one bucket, reduced, unpacked).  `average=False` when 1/world is already folded into
    the loss seed (render_loss_backward(n_rays_mean = global batch)).  Returns the number of collectives issued (0 or 1)."""
    r, w = world()
    if w == 1:
        return 0
    parts = []
    for m in models.values():
        if hasattr(m, "flat_grads"):
            parts.append(m.flat_grads(zero=False))
        else:
            for p in m.parameters():
                if p.grad is not None:
                    parts.append(p.grad)
    if not parts:
        return 0
    if len(parts) > 1 and parts[0].is_cuda and dist.get_backend() == "nccl":
        # several buffers (field gradient + embedding, or coarse + fine): one NCCL group launch reduces them in place -- no pack /
        # unpack copies of the 10 MB buffers (the bucket path below cost ~50 us per step on top of the collective)
        with dist._coalescing_manager(device=parts[0].device):
            for g in parts:
                dist.all_reduce(g, op=dist.ReduceOp.SUM)
        if average:
            torch._foreach_div_(parts, float(w))
        return 1
    if len(parts) == 1:
        bucket = parts[0].view(-1)
    else:
        bucket = torch.cat([g.reshape(-1) for g in parts])
    dist.all_reduce(bucket, op=dist.ReduceOp.SUM)
    if average:
        bucket.div_(w)
    if len(parts) > 1:
        off = 0
        for g in parts:
            g.copy_(bucket[off:off + g.numel()].view_as(g))
            off += g.numel()
    return 1


class SymmetricBuffer:
    """One flat parameter buffer and its gradient buffer living in symmetric memory: every rank can address every rank's copy
    (NVLink peer loads / stores), which is what the fused reduce-scatter + Adam + all-gather step (snb_adam_step_sharded) runs on."""

    def __init__(self, n: int, device):
        import torch.distributed._symmetric_memory as symm
        self.n = n
        self.n_pad = (n + 3) // 4 * 4                       # the kernel works on float4s; pad elements stay zero
        group = dist.group.WORLD.group_name
        self.params = symm.empty(self.n_pad, dtype=torch.float32, device=device)
        self.grads = symm.empty(self.n_pad, dtype=torch.float32, device=device)
        self.params.zero_(); self.grads.zero_()
        self.hp = symm.rendezvous(self.params, group)
        self.hg = symm.rendezvous(self.grads, group)
        self.param_ptrs = [int(x) for x in self.hp.buffer_ptrs]
        self.grad_ptrs = [int(x) for x in self.hg.buffer_ptrs]
        # multicast (NVLS) addresses of the same allocations: 0 when the fabric / driver offers none
        self.mc_params = int(getattr(self.hp, "multicast_ptr", 0) or 0)
        self.mc_grads = int(getattr(self.hg, "multicast_ptr", 0) or 0)

    def barrier(self):
        """Device-side barrier among the ranks on the current stream (signal pads of the symmetric allocation)."""
        self.hp.barrier()


def make_symmetric(models: Dict[str, torch.nn.Module]) -> Dict[int, "SymmetricBuffer"]:
    """Re-homes the flat parameter / gradient buffers of the fields (and the embedding table) in symmetric memory.  Returns
    {id(parameter the optimizer steps): SymmetricBuffer}; collective: every rank must call it with the same models."""
    out = {}
    for m in models.values():
        if hasattr(m, "adopt_flat_storage"):
            n = m.flat_params().numel()
            sb = SymmetricBuffer(n, m.flat_params().device)
            m.adopt_flat_storage(sb.params[:n], sb.grads[:n])
            out[id(m.flat_parameter())] = sb
        else:
            for p in m.parameters():
                sb = SymmetricBuffer(p.numel(), p.device)
                with torch.no_grad():
                    sb.params[:p.numel()].copy_(p.data.reshape(-1))
                    p.data = sb.params[:p.numel()].view(p.shape)
                    p.grad = sb.grads[:p.numel()].view(p.shape)
                out[id(p)] = sb
    return out
