"""Probe (N GPUs, torchrun): cost of the pieces of the fused data-parallel step: symmetric-memory barrier, snb_adam_step_sharded on the
10.5 MB buffer, NCCL all-reduce + snb_adam_step for comparison."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from satnerf_b200 import capi
from satnerf_b200 import dist as sdist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
n = 2635785
sb = sdist.SymmetricBuffer(n, dev)
sb.params.normal_(); sb.grads.normal_()
m = torch.zeros(sb.n_pad, device=dev); v = torch.zeros(sb.n_pad, device=dev)
g = torch.randn(n, device=dev); p = torch.randn(n, device=dev); m2 = torch.zeros(n, device=dev); v2 = torch.zeros(n, device=dev)
def timeit(fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
step = [0]
def sharded():
    step[0] += 1
    sb.barrier()
    capi.adam_step_sharded(sb.param_ptrs, sb.grad_ptrs, rank, m, v, sb.n_pad, 5e-4, 0.9, 0.999, 1e-8, 0.0, step[0], dev)
    sb.barrier()
def kernel_only():
    step[0] += 1
    capi.adam_step_sharded(sb.param_ptrs, sb.grad_ptrs, rank, m, v, sb.n_pad, 5e-4, 0.9, 0.999, 1e-8, 0.0, step[0], dev)
def nccl():
    dist.all_reduce(g); capi.adam_step(p, g, m2, v2, 5e-4, 0.9, 0.999, 1e-8, 0.0, 1)
r = {"barrier_us": timeit(sb.barrier), "sharded_kernel_us": timeit(kernel_only), "sharded_step_us": timeit(sharded), "nccl_allreduce_plus_adam_us": timeit(nccl)}
if rank == 0: print(world, "GPUs:", {k: round(x, 1) for k, x in r.items()}, flush=True)
dist.destroy_process_group()
