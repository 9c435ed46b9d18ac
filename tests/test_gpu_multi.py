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


def _worker_sharded(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from satnerf_b200 import train as trn
        from satnerf_b200 import dist as sdist
        from satnerf_b200.synth import synthetic_sat_rays
        from gpu_util import make_args
        R = 256                                             # global batch; every rank takes its shard
        rays, ts = synthetic_sat_rays(3 * R, seed=17)
        g = torch.Generator().manual_seed(18)
        rgbs = torch.rand(3 * R, 3, generator=g)
        finals, has_mc = [], None
        for sharded, nvls in ((True, True), (True, False), (False, False)):
            a = make_args(fc_units=128, precision="tc", lr=5e-4, batch_size=R, sharded_adam=sharded, sharded_nvls=nvls)
            torch.manual_seed(5)                            # same initialisation and noise on both runs
            system = trn.NeRFSystem(a, dev, train_len=3 * R)
            system.configure_optimizers()
            assert system.optimizer.sharded == sharded
            if sharded and nvls:
                has_mc = all(sb.mc_params and sb.mc_grads for sb in system.optimizer.shards.values())
            for it in range(3):
                lo, hi = sdist.shard_bounds(R, rank, world)
                sl = slice(it * R + lo, it * R + hi)
                torch.manual_seed(100 + it)                 # the render draws its jitter from the global generator
                batch = {"color": {"rays": rays[sl].to(dev), "rgbs": rgbs[sl].to(dev), "ts": ts[sl].reshape(-1, 1).to(dev)}}
                system.optimization_step(batch)
            torch.cuda.synchronize()
            finals.append((system.models["coarse"].flat_params().clone(), system.models["t"].weight.detach().clone()))
        ok, msg = True, f"multicast mappings: {has_mc}"
        pb, tb = finals[2]
        for name, (pa, ta) in (("multimem", finals[0]), ("peer", finals[1])):
            # fused reduce + Adam + deliver step == all-reduce + Adam: the same sum of two ranks' gradients, the same update arithmetic
            e1 = float((pa - pb).abs().max()); e2 = float((ta - tb).abs().max())
            if e1 > 1e-7 or e2 > 1e-7:
                ok, msg = False, f"sharded ({name}) vs all-reduce: params {e1}, embedding {e2}"
            ref = pa.clone(); dist.broadcast(ref, 0)
            if not torch.equal(ref, pa):
                ok, msg = False, f"replicas differ after the sharded step ({name})"
        q.put((rank, ok, msg))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_adam_step_equals_all_reduce_plus_adam():
    """The data-parallel step fused with its collective (snb_adam_step_sharded_multi over symmetric memory, one launch for the field
    and the embedding) against all-reduce + snb_adam_step over three steps, in both variants -- in-switch reduction / multicast
    delivery (`multimem.ld_reduce` / `multimem.st`, when the allocations have multicast mappings) and peer loads / stores:
    <= 1e-7 on parameters of O(0.1) (two ranks: the same two-term sums), replicas bit-identical."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker_sharded, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    print(res)
    assert sorted(r[:2] for r in res) == [(0, True), (1, True)], res


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
