#!/usr/bin/env python
"""Benchmark of the render_rays hot path (BASELINE.json metric: rays/s at 64 samples/ray).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision tc|fp32]

One "step" = one render_rays() call (forward, no_grad — the DSM-extraction / evaluation use of the path)
over a batch of 4096 synthetic sat-nerf rays x 64 samples with the default h=512 field (BASELINE.json
configs[1]); each of the N ranks renders its own 4096 rays (weak scaling, no data-path collective).
The same JSON line also reports, under "train", the training step of the path (forward + backward +
one NCCL all-reduce of the flat gradient buffer + Adam) on 1024 rays per rank.

Timing: W>=3 warm-up steps, then K steps each bracketed by CUDA events on the launching stream; a 256 MiB
buffer is written between steps to flush L2 (outside the events); max over ranks.
`--impl reference` times the CPU restatement of the reference (oracle/, torch CPU, all host threads) on a
bounded sample of the same workload (the reference itself is Python and is not present on the GPU box).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RAYS_PER_GPU = 4096
N_SAMPLES = 64
WIDTH = 512
TRAIN_RAYS = 1024
MACS_PER_POINT = 2_629_632          # sat-nerf h=512: 10h^2 + (14 + tau/2)h
FLOP_PER_RAY = 2 * MACS_PER_POINT * N_SAMPLES


def field_args(**kw):
    base = dict(model="sat-nerf", n_samples=N_SAMPLES, n_importance=0, noise_std=0.0, sc_lambda=0.0, chunk=1 << 20,
                fc_layers=8, fc_units=WIDTH, t_embbeding_tau=4, t_embbeding_vocab=30)
    base.update(kw)
    return argparse.Namespace(**base)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops"]), float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "measured"
    except Exception:
        return 1590.0, 1400.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.path = gpu_index, None, None

    def __enter__(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.path or not os.path.exists(self.path):
            return out
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        os.unlink(self.path)
        return out


def cpu_reference_rays_per_s(n_rays, reps, warm, threads=None):
    """Times the oracle (CPU restatement of rendering.py + models/satnerf.py, same torch ops as the reference)."""
    import torch
    from oracle import render_oracle as orc
    if threads:
        torch.set_num_threads(threads)
    args = field_args()
    import satnerf_b200 as sb
    torch.manual_seed(0)
    field = sb.load_model(args)
    params = {"coarse": {k: v.detach().clone() for k, v in field.state_dict().items()}, "t": torch.randn(30, 4)}
    rays, ts = orc.synthetic_sat_rays(n_rays, seed=1)
    times = []
    with torch.no_grad():
        for i in range(warm + reps):
            t0 = time.perf_counter()
            orc.render_rays(params, args, rays, ts)
            dt = time.perf_counter() - t0
            if i >= warm:
                times.append(dt)
    return n_rays / statistics.median(times), torch.get_num_threads(), times


def workload_config(world, precision="tc"):
    """The `config` object of the JSON line: identical for both arms (the reference arm times a bounded sample of it)."""
    return {"workload": "configs[1]: sat-nerf synthetic RPC rays, 4096 rays x 64 samples per GPU, h=512, 8 layers, "
                        "render_rays forward (no_grad)", "rays_per_gpu": RAYS_PER_GPU, "n_samples": N_SAMPLES,
            "fc_units": WIDTH, "precision": precision, "parallelism": f"ray-sharded x{world}, no collective in forward",
            "l2": "256 MiB buffer written between steps (outside the timed events)"}


def run_reference(opt, rank, world):
    if rank != 0:
        return
    n = 1024
    steps = max(1, opt.steps)
    rps, cores, times = cpu_reference_rays_per_s(n, steps, max(1, min(opt.warmup, 3)), threads=os.cpu_count())
    ms = 1e3 * statistics.median(times)
    sample = f"{n} rays x {N_SAMPLES} samples per step, sat-nerf h={WIDTH}, forward no_grad, torch CPU fp32, {cores} threads"
    line = {"impl": "reference", "metric": "rays/sec (64 samples/ray)", "value": rps, "unit": "rays/s", "n_gpus": opt.gpus,
            "steps": steps, "warmup": opt.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(opt.gpus, opt.precision),
            "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="tc", choices=["tc", "fp32"])
    ap.add_argument("--no-train", action="store_true", help="skip the training-step leg")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    opt = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if opt.impl == "reference":
        run_reference(opt, rank, world)
        return

    import torch
    import torch.distributed as dist
    import satnerf_b200 as sb
    from satnerf_b200 import capi, rendering
    from oracle import render_oracle as orc          # input generators + cpu_baseline leg only

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warm = max(3, opt.warmup)
    K = max(1, opt.steps)

    args = field_args(precision=opt.precision)
    torch.manual_seed(0)
    models = {"coarse": sb.load_model(args).to(dev), "t": torch.nn.Embedding(30, 4).to(dev)}
    rays_h, ts_h = orc.synthetic_sat_rays(RAYS_PER_GPU, seed=100 + rank)
    rays_h, ts_h = rays_h.pin_memory(), ts_h.pin_memory()
    rays, ts = rays_h.to(dev), ts_h.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    # time the dominant kernel (the fused render pass) with its own pair of events
    kern_ms = []
    orig_fwd = capi.render_forward
    timing = {"on": False}

    def timed_forward(*a, **k):
        if not timing["on"]:
            return orig_fwd(*a, **k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); orig_fwd(*a, **k); e1.record()
        kern_ms.append((e0, e1))

    capi.render_forward = timed_forward

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        with torch.no_grad():
            return rendering.render_rays(models, args, rays, ts)

    def timed_steps(fn, n):
        evs = []
        for _ in range(n):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs)

    for _ in range(warm):
        step()
    barrier()
    capi.launch_count(reset=True)
    timing["on"] = True
    with ClockSampler(local) as clk:
        total_ms = timed_steps(step, K)
        barrier()
        timing["on"] = False
        launches = capi.launch_count(reset=True)

        # end-to-end through the public API with host buffers
        rgb_h = torch.empty(RAYS_PER_GPU, 3).pin_memory()
        depth_h = torch.empty(RAYS_PER_GPU).pin_memory()

        def e2e_step():
            r = rays_h.to(dev, non_blocking=True)
            t = ts_h.to(dev, non_blocking=True)
            with torch.no_grad():
                out = rendering.render_rays(models, args, r, t)
            rgb_h.copy_(out["rgb_coarse"], non_blocking=True)
            depth_h.copy_(out["depth_coarse"], non_blocking=True)

        for _ in range(3):
            e2e_step()
        barrier()
        e2e_ms = timed_steps(e2e_step, K)
        barrier()

        train = None
        if not opt.no_train:
            from satnerf_b200 import train as trn
            train = trn.bench_training_step(args, dev, rank, world, TRAIN_RAYS, warm, K, flush)
        clocks = clk.summary() if rank == 0 else None
    kernel_ms = sum(a.elapsed_time(b) for a, b in kern_ms) / max(1, len(kern_ms))

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    total_ms = max_over_ranks(total_ms)
    e2e_ms = max_over_ranks(e2e_ms)
    kernel_ms = max_over_ranks(kernel_ms)
    if rank == 0:
        burst, sustained, src = peaks()
        rays_total = world * RAYS_PER_GPU * K
        value = rays_total / (total_ms * 1e-3)
        achieved = RAYS_PER_GPU * FLOP_PER_RAY / (kernel_ms * 1e-3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get("render_forward_dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": "rays/sec (64 samples/ray)", "value": value, "unit": "rays/s", "n_gpus": world, "steps": K, "warmup": warm,
            "ms_per_step": total_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f16 operands / f32 accumulate (first layer, heads, compositing f32)" if opt.precision == "tc" else "f32",
            "data": "synthetic",
            "config": workload_config(world, opt.precision),
            "e2e": {"value": rays_total / (e2e_ms * 1e-3), "unit": "rays/s",
                    "h2d_bytes_per_step": RAYS_PER_GPU * (11 * 4 + 8), "d2h_bytes_per_step": RAYS_PER_GPU * 16},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": burst, "unit": "TFLOP/s", "frac": achieved / burst,
                         "traffic": traffic, "peak_source": f"{src} cuBLAS bf16 burst ({burst}); sustained {sustained}",
                         "kernel_ms": kernel_ms, "flop_per_launch": RAYS_PER_GPU * FLOP_PER_RAY},
            "clocks": clocks,
        }
        if train is not None:
            line["train"] = train
        if world == 1 and not opt.no_cpu:
            rps, cores, times = cpu_reference_rays_per_s(1024, 5, 2)
            line["cpu_baseline"] = {"value": rps, "unit": "rays/s", "cores": cores, "kind": "port",
                                    "sample": f"1024 rays x 64 samples, sat-nerf h=512, forward no_grad, torch CPU fp32, median of 5 "
                                              f"({1e3 * statistics.median(times):.0f} ms each)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
