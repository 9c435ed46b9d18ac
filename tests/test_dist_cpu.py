"""CPU tests (gloo, world_size 2) of the multi-GPU host logic: ray sharding and the single flat-gradient all-reduce."""
import argparse
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from satnerf_b200 import dist as sdist


def test_shard_bounds_cover_all_rays():
    for n in (0, 1, 7, 4096, 65537):
        for w in (1, 2, 3, 8):
            spans = [sdist.shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import satnerf_b200 as sb
        a = argparse.Namespace(model="sat-nerf", fc_layers=8, fc_units=32, t_embbeding_tau=4)
        torch.manual_seed(0)
        models = {"coarse": sb.load_model(a), "t": torch.nn.Embedding(30, 4)}
        g = models["coarse"].flat_grads()
        g.fill_(float(rank + 1))
        models["t"].weight.grad = torch.full_like(models["t"].weight, float(10 * (rank + 1)))
        n = sdist.all_reduce_gradients(models)
        ok = n == 1 and torch.allclose(g, torch.full_like(g, 1.5)) and torch.allclose(models["t"].weight.grad, torch.full((30, 4), 15.0))
        ok = ok and all(torch.allclose(p.grad, torch.full_like(p, 1.5)) for p in models["coarse"].parameters())
        rays = torch.arange(10.0).reshape(10, 1).repeat(1, 11)
        mine, _ = sdist.shard_rays(rays)
        gathered = sdist.gather_rays({"x": mine}, 10)
        ok = ok and torch.equal(gathered["x"], rays)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_flat_gradient_all_reduce_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, True), (1, True)]
