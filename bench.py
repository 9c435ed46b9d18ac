#!/usr/bin/env python
"""Benchmark of the render_rays hot path (BASELINE.json metric: rays/s at 64 samples/ray).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision tc|fp32] [--only LEGS]

Headline (`value`, `e2e`, `roofline`): BASELINE.json configs[1] -- one step = one render_rays() call (forward, no_grad) over
4096 synthetic sat-nerf rays x 64 samples per GPU with the default h=512 field; N ranks render their own rays (weak scaling,
no data-path collective).  The same JSON line carries sub-records (each with its own `roofline`):

    train     fwd + in-kernel loss seed + bwd + ONE gradient all-reduce + Adam, 1024 rays per rank fed by DeviceRaySampler
    config3   sat-nerf coarse+fine (64 + 32 importance = 96 fine samples; 160 field evaluations per ray), 8192 rays / rank,
              forward, and the depth-supervised training step (colour batch + depth batch, main.py:127-142)
    config4   s-nerf + solar-correction pass (128 evaluations per ray), 4096 rays / rank: forward and the training step with
              the gradient all-reduce (BASELINE configs[3] quotes it on 4 GPUs: run with --gpus 4)
    config5   create_satnerf_dsm: one 512x512 tile (262 144 rays) in 65 536-ray batches through batched_inference, the tile
              split across the N ranks (STRONG scaling), repeated back to back for >= 2 s so the sustained clock applies
    precise   the headline workload at fp32-level accuracy: precision 'tcx3' (tensor cores, fp16 hi+lo operands) and 'fp32' (FFMA)
    gpu_eager the reference's algorithm (oracle port, stock torch eager fp32) on the same B200 -- the honest GPU comparator
    cpu_baseline  the same on the host cores (bounded sample)

Timing: W >= 3 warm-up steps, then K steps each bracketed by CUDA events on the launching stream; a 256 MiB buffer is written
between steps to flush L2 (outside the events); max over ranks.  `--impl reference` times the CPU restatement of the reference
(oracle/, torch CPU, all host threads) on a bounded sample of the headline workload (the reference itself is Python and is not
present on the GPU box).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RAYS_PER_GPU = 4096
N_SAMPLES = 64
WIDTH = 512
TRAIN_RAYS = 1024
REF_SAMPLE_RAYS = 1024
MACS = {"sat-nerf": 2_629_632, "s-nerf": 2_497_280, "sat-nerf-nobeta": 2_497_280,      # per point, h=512
        "density-only": 512 * 3 + 6 * 512 * 512 + 512 * 515 + 512}                        # trunk (skip layer 515 wide) + sigma head
FLOP_PER_RAY = 2 * MACS["sat-nerf"] * N_SAMPLES                                         # 336.6 MFLOP


def field_args(**kw):
    base = dict(model="sat-nerf", n_samples=N_SAMPLES, n_importance=0, noise_std=0.0, sc_lambda=0.0, chunk=1 << 20,
                fc_layers=8, fc_units=WIDTH, t_embbeding_tau=4, t_embbeding_vocab=30, batch_size=TRAIN_RAYS, lr=5e-4)
    base.update(kw)
    return argparse.Namespace(**base)


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops"]), float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "measured"
    except Exception:
        return 1590.0, 1400.0, "fallback"


def hbm_peak():
    """Measured HBM copy bandwidth (GB/s) of MEASURED_PEAKS.json, else the profiling recipe's fallback."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6650.0


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


def reference_setup(args, n_rays, device="cpu"):
    """Inputs of the reference-algorithm legs (oracle port): default-init parameters as plain tensors + seeded rays."""
    import torch
    import satnerf_b200 as sb
    from satnerf_b200.synth import synthetic_sat_rays
    torch.manual_seed(0)
    field = sb.load_model(args)
    params = {"coarse": {k: v.detach().clone().to(device) for k, v in field.state_dict().items()}, "t": torch.randn(30, 4).to(device)}
    rays, ts = synthetic_sat_rays(n_rays, seed=1)
    return params, rays.to(device), ts.to(device)


def cpu_reference_rays_per_s(n_rays, reps, warm, threads=None):
    """Times the oracle (CPU restatement of rendering.py + models/satnerf.py, same torch ops as the reference)."""
    import torch
    from oracle import render_oracle as orc
    if threads:
        torch.set_num_threads(threads)
    args = field_args()
    params, rays, ts = reference_setup(args, n_rays)
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
    """The `config` object of the JSON line: identical for both arms; it states what the reference arm really times."""
    return {"workload": "configs[1]: sat-nerf synthetic RPC rays, 4096 rays x 64 samples per GPU, h=512, 8 layers, "
                        "render_rays forward (no_grad)", "rays_per_gpu": RAYS_PER_GPU, "n_samples": N_SAMPLES,
            "fc_units": WIDTH, "precision": precision, "parallelism": f"ray-sharded x{world}, no collective in forward",
            "l2": "256 MiB buffer written between steps (outside the timed events)",
            "reference_arm_sample": f"--impl reference times {REF_SAMPLE_RAYS} rays x {N_SAMPLES} samples of this workload per step "
                                    "(bounded CPU sample; rays/s is size-independent: the reference chunks points by 5120)"}


def run_reference(opt, rank, world):
    if rank != 0:
        return
    n = REF_SAMPLE_RAYS
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
    ap.add_argument("--precision", default="tc", choices=["tc", "tcx3", "fp32"])
    ap.add_argument("--only", default="", help="comma list of extra legs to run (train,config3,config4,config5,geometry,precise,gpu_eager,cpu); default: all")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step legs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    opt = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if opt.impl == "reference":
        run_reference(opt, rank, world)
        return
    legs = set(x for x in opt.only.split(",") if x) or {"train", "config3", "config4", "config5", "geometry", "precise", "gpu_eager", "cpu"}
    if opt.no_train:
        legs.discard("train")
    if opt.no_cpu:
        legs.discard("cpu")

    import torch
    import torch.distributed as dist
    import satnerf_b200 as sb
    from satnerf_b200 import capi, rendering
    from satnerf_b200 import dist as sdist
    from satnerf_b200 import train as trn
    from satnerf_b200.synth import synthetic_sat_rays

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=240))      # a desynchronised collective fails in minutes, not after NCCL's default 10
    warm = max(3, opt.warmup)
    K = max(1, opt.steps)
    burst, sustained, peak_src = peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_steps(fn, n, do_flush=True):
        evs = []
        for _ in range(n):
            if do_flush:
                flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs)

    def measure(fn, n=K, w=warm):
        """warm-up, barrier, n timed steps (events, L2 flushed between), max over ranks -> (ms per step, launches per step)."""
        for _ in range(w):
            fn()
        barrier()
        capi.launch_count(reset=True)
        ms = timed_steps(fn, n)
        launches = capi.launch_count(reset=True)
        barrier()
        return max_over_ranks(ms) / n, launches / n

    def roof(rays_per_s, flop_per_ray, peak, peak_name, n_gpus=world):
        ach = rays_per_s * flop_per_ray / 1e12 / n_gpus
        return {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak, "traffic": None,
                "peak_source": f"{peak_src} cuBLAS bf16 {peak_name}", "flop_per_ray": flop_per_ray, "per_gpu": True}

    def models_for(a, seed=0):
        torch.manual_seed(seed)
        ms = {"coarse": sb.load_model(a).to(dev)}
        if a.n_importance > 0:
            ms["fine"] = sb.load_model(a).to(dev)
        if a.model == "sat-nerf":
            ms["t"] = torch.nn.Embedding(30, 4).to(dev)
        return ms

    # ------------------------------------------------------------------ headline: configs[1]
    args = field_args(precision=opt.precision)
    models = models_for(args)
    rays_h, ts_h = synthetic_sat_rays(RAYS_PER_GPU, seed=100 + rank)
    rays_h, ts_h = rays_h.pin_memory(), ts_h.pin_memory()
    rays, ts = rays_h.to(dev), ts_h.to(dev)

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

    def step():
        with torch.no_grad():
            return rendering.render_rays(models, args, rays, ts)

    extra = {}
    with ClockSampler(local) as clk:
        for _ in range(warm):
            step()
        barrier()
        capi.launch_count(reset=True)
        timing["on"] = True
        total_ms = timed_steps(step, K)
        barrier()
        timing["on"] = False
        launches = capi.launch_count(reset=True)
        capi.render_forward = orig_fwd

        # end-to-end through the public API with host buffers: rays in from pinned memory, the WHOLE result dict back to pinned memory
        out0 = step()
        host_out = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in out0.items()}
        d2h_bytes = sum(v.numel() * v.element_size() for v in out0.values())
        del out0

        def e2e_step(keys=None):
            r = rays_h.to(dev, non_blocking=True)
            t = ts_h.to(dev, non_blocking=True)
            with torch.no_grad():
                out = rendering.render_rays(models, args, r, t)
            for k in (keys or out):
                host_out[k].copy_(out[k], non_blocking=True)

        e2e_serial_ms, _ = measure(e2e_step, K, 3)
        # the same through satnerf_b200.hostio.HostPipeline: the copy-out of step i (second stream) runs under the render pass of
        # step i + 1.  ONE pair of events around the K steps, L2 flush INSIDE the timed region, the last copy-out joined before the
        # closing event.
        from satnerf_b200.hostio import HostPipeline
        pipe = HostPipeline(models, args, dev, depth=2)
        for _ in range(3):
            pipe.submit(rays_h, ts_h)
        pipe.wait()
        barrier()
        ep0, ep1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ep0.record()
        for _ in range(K):
            flush.fill_(1)
            pipe.submit(rays_h, ts_h)
        pipe.join()
        ep1.record()
        torch.cuda.synchronize()
        pipe.wait()
        e2e_ms = ep0.elapsed_time(ep1)
        assert pipe.d2h_bytes == d2h_bytes
        barrier()
        e2e_small_ms, _ = measure(lambda: e2e_step(("rgb_coarse", "depth_coarse")), K, 3)

        # ------------------------------------------------------------------ sub-records
        if "train" in legs:
            extra["train"] = trn.bench_training_step(args, dev, rank, world, TRAIN_RAYS, warm, K, flush)
            tr = extra["train"]
            tr["roofline"] = roof(tr["value"], 3 * FLOP_PER_RAY, burst, "burst")

        if "config3" in legs:
            a3 = field_args(precision=opt.precision, n_importance=32)
            m3 = models_for(a3, 1)
            R3 = 8192
            r3, t3 = synthetic_sat_rays(R3, seed=300 + rank)
            r3, t3 = r3.to(dev), t3.to(dev)

            def f3():
                with torch.no_grad():
                    rendering.render_rays(m3, a3, r3, t3)
            ms3, l3 = measure(f3, max(5, K // 2))
            flop3 = 2 * MACS["sat-nerf"] * 160
            rps3 = world * R3 / (ms3 * 1e-3)
            extra["config3"] = {"workload": "configs[2]: sat-nerf coarse (64) + fine (64+32 importance samples), 8192 rays per GPU, render_rays forward",
                                "value": rps3, "unit": "rays/s", "ms_per_step": ms3, "gpu_launches": l3, "roofline": roof(rps3, flop3, burst, "burst")}
            if not opt.no_train:
                extra["config3"]["train"] = trn.bench_training_step(
                    field_args(precision=opt.precision, n_importance=32, ds_lambda=1000.0, ds_drop=0.25, max_train_steps=300000, ds_noweights=False),
                    dev, rank, world, 2048, 3, max(3, K // 4), flush, depth_batch=True)
                t3r = extra["config3"]["train"]
                t3r["roofline"] = roof(t3r["value"], 3 * 2 * flop3, burst, "burst")      # colour batch + depth batch, 160 evaluations per ray each
            del m3

        if "config4" in legs:
            a4 = field_args(precision=opt.precision, model="s-nerf", sc_lambda=0.05)
            m4 = models_for(a4, 2)
            r4, _ = synthetic_sat_rays(RAYS_PER_GPU, seed=400 + rank)
            r4 = r4.to(dev)

            def f4():
                with torch.no_grad():
                    rendering.render_rays(m4, a4, r4, None)
            ms4, l4 = measure(f4, K)
            flop4 = 2 * MACS["s-nerf"] * 128
            rps4 = world * RAYS_PER_GPU / (ms4 * 1e-3)
            extra["config4"] = {"workload": "configs[3]: s-nerf + solar-correction pass (2 field passes, 128 evaluations per ray), 4096 rays per GPU "
                                            f"(BASELINE quotes it on 4 GPUs; this run: {world})",
                                "value": rps4, "unit": "rays/s", "ms_per_step": ms4, "gpu_launches": l4, "roofline": roof(rps4, flop4, burst, "burst")}
            if not opt.no_train:
                extra["config4"]["train"] = trn.bench_training_step(a4, dev, rank, world, RAYS_PER_GPU, 3, max(3, K // 2), flush)
                t4r = extra["config4"]["train"]
                t4r["roofline"] = roof(t4r["value"], 3 * flop4, burst, "burst")
            del m4

        if "config5" in legs:
            # create_satnerf_dsm.py:72-78: all rays of one 512x512 view through batched_inference in 65 536-ray chunks
            a5 = field_args(precision=opt.precision, chunk=65536)
            tile = 512 * 512
            lo, hi = sdist.shard_bounds(tile, rank, world)
            r5, t5 = synthetic_sat_rays(tile, n_images=1, seed=500)
            gy, gx = torch.meshgrid(torch.linspace(-1, 1, 512), torch.linspace(-1, 1, 512), indexing="ij")
            r5[:, 0], r5[:, 1] = gx.reshape(-1), gy.reshape(-1)
            r5, t5 = r5[lo:hi].to(dev), t5[lo:hi].to(dev)
            rec5 = {"workload": "configs[4]: create_satnerf_dsm 512x512 tile (262 144 rays) in 65 536-ray batches through batched_inference, "
                                f"tile split across {world} GPU(s) (strong scaling), repeated back to back for >= 2 s (sustained clock)",
                    "scaling": "strong", "unit": "rays/s"}
            for mode, macs in (("depth", MACS["sat-nerf-nobeta"]), ("full", MACS["sat-nerf"]), ("depth_only", MACS["density-only"])):
                a5.render_outputs = mode

                def f5():
                    return rendering.batched_inference(models, r5, t5, a5)
                for _ in range(2):
                    f5()
                barrier()
                t_one = timed_steps(f5, 2) / 2
                reps = max(3, int(2200.0 / max(t_one, 1e-3)) + 1)
                barrier()
                capi.launch_count(reset=True)
                ms5 = max_over_ranks(timed_steps(f5, reps)) / reps
                l5 = capi.launch_count(reset=True) / reps
                barrier()
                rps5 = tile / (ms5 * 1e-3)
                rec5[mode] = {"value": rps5, "ms_per_tile": ms5, "tiles_timed": reps, "gpu_launches_per_tile": l5,
                              "outputs": {"depth": "rgb + depth per ray (uncertainty head skipped)", "full": "full reference result dict",
                                          "depth_only": "depth per ray: density trunk + sigma head only (the fields' sigma_only=True) -- all create_satnerf_dsm.py:78 "
                                                        "consumes; FLOPs counted are the executed ones"}[mode],
                              "roofline": roof(rps5, 2 * macs * N_SAMPLES, sustained, "sustained (>= 2 s back to back)")}
            rec5["value"] = rec5["depth"]["value"]
            # end to end: host rays in, depth + rgb back (what create_satnerf_dsm.py:78-110 consumes)
            a5.render_outputs = "depth"
            r5h, t5h = r5.cpu().pin_memory(), t5.cpu().pin_memory()
            d5h, c5h = torch.empty(hi - lo).pin_memory(), torch.empty(hi - lo, 3).pin_memory()

            def e5():
                out = rendering.batched_inference(models, r5h.to(dev, non_blocking=True), t5h.to(dev, non_blocking=True), a5)
                d5h.copy_(out["depth_coarse"], non_blocking=True); c5h.copy_(out["rgb_coarse"], non_blocking=True)
            ms5e, _ = measure(e5, 5, 2)
            rec5["e2e"] = {"value": tile / (ms5e * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": (hi - lo) * 52, "d2h_bytes_per_step": (hi - lo) * 16}
            extra["config5"] = rec5

        if "geometry" in legs and rank == 0:
            # SURVEY 8 f3 / f4, the neighbours of the render path in create_satnerf_dsm.py: rays of one 512x512 view from its RPC model
            # (datasets/satellite.py:18-65, :185-227) and the DSM of the rendered depths (:246-338).  float64 element-wise kernels; the
            # numpy oracle (the reference's own arithmetic + restated rpcm / pyproj / plyflatten) is timed beside them on a bounded sample.
            from satnerf_b200 import geo
            from oracle import geo_oracle as gor
            import numpy as np
            rpc = gor.synthetic_rpc(seed=1)
            center, rng = [799046.0, -5451605.0, 3202158.0], 400.0
            sg = geo.SatelliteGeometry(center, rng, device=dev)
            tile = 512 * 512

            def fr():
                return sg.rays_for_image(rpc, 512, 512, -10.0, 60.0, 50.0, 140.0)
            def local_ms(fn, n=10, w=3):          # rank-0-only leg: no barrier / all-reduce in here (the other ranks are not in this block)
                for _ in range(w):
                    fn()
                torch.cuda.synchronize()
                capi.launch_count(reset=True)
                ms = timed_steps(fn, n) / n
                return ms, capi.launch_count(reset=True) / n

            rays_g = fr()
            depth_g = rays_g[:, 7] * 0.5
            ms_r, l_r = local_ms(fr)

            def fd():
                return sg.get_dsm_from_nerf_prediction(rays_g, depth_g)
            ms_d, l_d = local_ms(fd)
            n_s = 16384
            cols, rows = np.meshgrid(np.arange(128), np.arange(128))
            t0 = time.perf_counter(); ref_r = gor.normalize_rays(gor.get_rays(cols.ravel(), rows.ravel(), rpc, -10.0, 60.0), center, rng); t_r = time.perf_counter() - t0
            t0 = time.perf_counter()
            la, lo_, al = gor.latlonalt_from_prediction(ref_r, 0.5 * ref_r[:, 7], center, rng); e_, n_ = gor.utm_forward(la, lo_, 17)
            t_d = time.perf_counter() - t0          # (points + projection; the pure-Python raster loop of the oracle is not a baseline)
            extra["geometry"] = {
                "workload": "one 512x512 view (262 144 pixels): RPC ray generation; rays + depths -> UTM point cloud -> 0.5 m DSM raster",
                "rays": {"value": tile / (ms_r * 1e-3), "unit": "pixels/s", "ms_per_tile": ms_r, "gpu_launches": l_r,
                         "roofline": {"bound": "hbm", "achieved": tile * 44 / (ms_r * 1e-3) / 1e9, "peak": hbm_peak(), "unit": "GB/s",
                                      "frac": tile * 44 / (ms_r * 1e-3) / 1e9 / hbm_peak(), "traffic": None,
                                      "note": "44 B written per pixel; the kernel is float64-arithmetic bound (two iterative inverse-RPC solves per pixel), not HBM bound"},
                         "cpu_baseline": {"value": n_s / t_r, "unit": "pixels/s", "cores": 1, "kind": "port", "sample": f"{n_s} pixels, numpy float64"}},
                "dsm": {"value": tile / (ms_d * 1e-3), "unit": "points/s", "ms_per_tile": ms_d, "gpu_launches": l_d,
                        "cpu_baseline": {"value": n_s / t_d, "unit": "points/s", "cores": 1, "kind": "port",
                                         "sample": f"{n_s} points, numpy float64, ECEF -> geodetic -> UTM only (no raster)"}}}

        if "precise" in legs and rank == 0 and opt.precision == "tc":
            # the fp32-level modes on the headline workload: 'tcx3' (tensor cores, fp16 hi+lo operands, layer by layer) and the FFMA path
            import copy
            pr = {}
            for prec, n in (("tcx3", 5), ("fp32", 3)):
                ap_ = copy.copy(args); ap_.precision = prec

                def fp():
                    with torch.no_grad():
                        rendering.render_rays(models, ap_, rays, ts)
                fp(); fp()
                torch.cuda.synchronize()
                capi.launch_count(reset=True)
                msp = timed_steps(fp, n) / n
                pr[prec] = {"value": RAYS_PER_GPU / (msp * 1e-3), "unit": "rays/s", "ms_per_step": msp, "gpu_launches": capi.launch_count(reset=True) / n}
            pr["what"] = ("same workload as `value` at fp32-level accuracy (the trained-like fixture passes 1e-3 elementwise on both; the fp16-operand kernel "
                          "reaches 2e-2 there): tcx3 = 3 tcgen05 MMAs per K-step on fp16 hi+lo operands, fp32 = FFMA CUDA-core path")
            extra["precise"] = pr

        if "gpu_eager" in legs and rank == 0:
            # the reference's algorithm with stock torch eager on this GPU: same ops as rendering.py + models/satnerf.py
            from oracle import render_oracle as orc
            pg, rg, tg = reference_setup(args, RAYS_PER_GPU, dev)
            prev = torch.backends.cuda.matmul.allow_tf32
            torch.backends.cuda.matmul.allow_tf32 = False

            def fe():
                with torch.no_grad():
                    orc.render_rays(pg, args, rg, tg)
            for _ in range(3):
                fe()
            torch.cuda.synchronize()
            mse = timed_steps(fe, 5) / 5
            torch.backends.cuda.matmul.allow_tf32 = prev
            extra["gpu_eager"] = {"value": RAYS_PER_GPU / (mse * 1e-3), "unit": "rays/s", "ms_per_step": mse, "n_gpus": 1,
                                  "what": "oracle port of rendering.py + models/satnerf.py, stock torch eager fp32 (TF32 off) on cuda:0, same workload as `value`"}
            del pg
        if world > 1:
            dist.barrier()
        clocks = clk.summary() if rank == 0 else None
    kernel_ms = sum(a.elapsed_time(b) for a, b in kern_ms) / max(1, len(kern_ms))

    total_ms = max_over_ranks(total_ms)
    e2e_ms = max_over_ranks(e2e_ms)
    e2e_serial_ms = max_over_ranks(e2e_serial_ms)
    kernel_ms = max_over_ranks(kernel_ms)
    if rank == 0:
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
                    "h2d_bytes_per_step": RAYS_PER_GPU * (11 * 4 + 8), "d2h_bytes_per_step": d2h_bytes,
                    "what": "satnerf_b200.hostio.HostPipeline: pinned host rays -> render_rays -> the WHOLE result dict (rgb, depth, weights, transparency, "
                            "albedo, sun, sky, beta) to pinned host every step; the copy-out of step i runs on a second stream under the render pass "
                            "of step i+1; one event pair around the K steps, the 256 MiB L2 flush of every step INSIDE the timed region"},
            "e2e_serial": {"value": rays_total / (e2e_serial_ms * K * 1e-3), "unit": "rays/s", "d2h_bytes_per_step": d2h_bytes,
                           "what": "same copies issued serially on one stream (rays in, render_rays, whole dict out), per-step events, flush outside"},
            "e2e_ray_outputs": {"value": world * RAYS_PER_GPU / (e2e_small_ms * 1e-3), "unit": "rays/s", "d2h_bytes_per_step": RAYS_PER_GPU * 16,
                                "what": "same, reading back rgb + depth only"},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": burst, "unit": "TFLOP/s", "frac": achieved / burst,
                         "traffic": traffic, "peak_source": f"{peak_src} cuBLAS bf16 burst ({burst}); sustained {sustained}",
                         "kernel_ms": kernel_ms, "flop_per_launch": RAYS_PER_GPU * FLOP_PER_RAY},
            "clocks": clocks,
        }
        line.update(extra)
        if world == 1 and "cpu" in legs:
            rps, cores, times = cpu_reference_rays_per_s(REF_SAMPLE_RAYS, 5, 2)
            line["cpu_baseline"] = {"value": rps, "unit": "rays/s", "cores": cores, "kind": "port",
                                    "sample": f"{REF_SAMPLE_RAYS} rays x 64 samples, sat-nerf h=512, forward no_grad, torch CPU fp32, median of 5 "
                                              f"({1e3 * statistics.median(times):.0f} ms each)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
