"""Two-GPU parity (NCCL, one process per GPU; skipped on a single-GPU box) -- SURVEY.md §4 / §8e:
  * the N-way ray-sharded render, gathered, equals the 1-GPU render bit for bit (rays are independent: no collective in the
    forward, the same kernel on the same rows);
  * all-reduced gradients of the shards (loss seed normalised by the GLOBAL batch, plain sum all-reduce) equal the gradients of
    the concatenated batch on one GPU: tolerance 2e-5 of each tensor's max |grad| (fp32 sums in a different association
    order: per-shard split-K partials, then the ring's), embedding gradient included;
  * broadcast_parameters makes replicas with different local seeds agree.
"""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import satnerf_b200 as sb
        from satnerf_b200 import dist as sdist
        from satnerf_b200.rendering import render_loss_backward
        from satnerf_b200.synth import synthetic_sat_rays
        from gpu_util import make_args
        a = make_args(fc_units=128, precision="tc")
        torch.manual_seed(100 + rank)                      # different local seeds: the broadcast must fix that
        models = {"coarse": sb.load_model(a).to(dev), "t": torch.nn.Embedding(30, 4).to(dev)}
        n_coll = sdist.broadcast_parameters(models)
        R = 510
        rays, ts = synthetic_sat_rays(R, seed=7)
        g = torch.Generator().manual_seed(8)
        u, nz, tgt = torch.rand(R, 64, generator=g), torch.randn(R, 64, generator=g), torch.rand(R, 3, generator=g)
        lo, hi = sdist.shard_bounds(R, rank, world)
        # ---- forward: shard, render, gather
        with torch.no_grad():
            mine = sb.render_rays(models, a, rays[lo:hi].to(dev), ts[lo:hi].to(dev), _draws=[u[lo:hi], nz[lo:hi]])
            gathered = sdist.gather_rays(mine, R)
        # ---- training: loss seed with the global mean, sum all-reduce
        for m in models.values():
            m.zero_grad(set_to_none=True)
        models["coarse"].flat_grads(zero=True)
        models["t"].weight.grad = torch.zeros_like(models["t"].weight)
        render_loss_backward(models, a, rays[lo:hi].to(dev), ts[lo:hi].to(dev), color=("beta", tgt[lo:hi].to(dev)), n_rays_mean=R,
                             _draws=[u[lo:hi], nz[lo:hi]])
        n_red = sdist.all_reduce_gradients(models, average=False)
        ok, msg = True, ""
        if rank == 0:
            with torch.no_grad():
                full = sb.render_rays(models, a, rays.to(dev), ts.to(dev), _draws=[u, nz])
            for k, v in full.items():
                if not torch.equal(v, gathered[k]):
                    ok, msg = False, f"forward {k}: max diff {float((v - gathered[k]).abs().max())}"
            g_shard = models["coarse"].flat_grads(zero=False).clone(); gt_shard = models["t"].weight.grad.clone()
            models["coarse"].flat_grads(zero=True); models["t"].weight.grad.zero_()
            render_loss_backward(models, a, rays.to(dev), ts.to(dev), color=("beta", tgt.to(dev)), n_rays_mean=R, _draws=[u, nz])
            g_full = models["coarse"].flat_grads(zero=False); gt_full = models["t"].weight.grad
            off = 0
            for name, p in models["coarse"].named_parameters():
                n = p.numel()
                a_, b_ = g_shard[off:off + n], g_full[off:off + n]
                err = float((a_ - b_).abs().max() / b_.abs().max().clamp_min(1e-20))
                if err > 2e-5:
                    ok, msg = False, f"grad {name}: {err}"
                off += n
            err_t = float((gt_shard - gt_full).abs().max() / gt_full.abs().max().clamp_min(1e-20))
            if err_t > 2e-5:
                ok, msg = False, f"grad embedding: {err_t}"
            if not (n_coll == 2 and n_red == 1):
                ok, msg = False, f"collectives: broadcast {n_coll}, all-reduce {n_red}"
        # replicas agree after the broadcast
        flat = models["coarse"].flat_params()
        ref = flat.clone(); dist.broadcast(ref, 0)
        if not torch.equal(ref, flat):
            ok, msg = False, "parameters differ across ranks after broadcast_parameters"
        q.put((rank, ok, msg))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_gpu_shards_equal_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(r[:2] for r in res) == [(0, True), (1, True)], res
