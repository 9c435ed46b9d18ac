"""Probe: torch symmetric memory on this box (2+ GPUs): rendezvous, peer pointers, device barrier, and a timing of NCCL all-reduce of the
10.5 MB flat gradient against torch's symmetric-memory two-shot all-reduce.  torchrun --nproc-per-node N profiles/dev/symm_probe.py"""
import os, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
n = 2635785 + 120
n = (n + 1023) // 1024 * 1024
t = symm.empty(n, dtype=torch.float32, device=dev)
hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
print(rank, "rendezvous ok", "world", hdl.world_size, "ptrs", [hex(p) for p in hdl.buffer_ptrs][:2], "signal pads", len(hdl.signal_pad_ptrs), flush=True)
t.fill_(rank + 1.0)
hdl.barrier()
peer = hdl.get_buffer((rank + 1) % world, (n,), torch.float32)
print(rank, "peer value", float(peer[0]), flush=True)
g = torch.full((n,), float(rank + 1), device=dev)
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
print(rank, "nccl all_reduce us", timeit(lambda: dist.all_reduce(g)), flush=True)
for name in ("two_shot_all_reduce_", "one_shot_all_reduce", "multimem_all_reduce_"):
    try:
        op = getattr(torch.ops.symm_mem, name)
        us = timeit(lambda: op(t, "sum", dist.group.WORLD.group_name))
        print(rank, name, "us", us, flush=True)
    except Exception as e:
        print(rank, name, "failed:", str(e)[:120], flush=True)
dist.destroy_process_group()
