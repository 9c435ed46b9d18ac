"""HostPipeline (satnerf_b200/hostio.py): host rays in, result dict out through pinned buffers with the copy-out on a second stream.
Every batch must equal the plain render_rays call on the same rays with the same random draws, bit for bit, whatever the overlap."""
import pytest
import torch

from gpu_util import make_args
from oracle import render_oracle as orc

pytestmark = pytest.mark.gpu


def test_host_pipeline_matches_render_rays():
    import satnerf_b200 as sb
    from satnerf_b200.hostio import HostPipeline
    args = make_args(fc_units=128, precision="tc")
    torch.manual_seed(3)
    ms = {"coarse": sb.load_model(args).cuda(), "t": torch.nn.Embedding(30, 4).cuda()}
    batches = []
    for i in range(5):
        rays, ts = orc.synthetic_sat_rays(700 + 13 * i, seed=60 + i)      # ragged batch sizes: the slot buffers are re-sized
        batches.append((rays.pin_memory(), ts.pin_memory()))
    want = []
    for i, (r, t) in enumerate(batches):
        torch.manual_seed(100 + i)
        with torch.no_grad():
            want.append({k: v.cpu() for k, v in sb.render_rays(ms, args, r.cuda(), t.cuda()).items()})
    pipe = HostPipeline(ms, args, "cuda", depth=2)
    got = []
    for i, (r, t) in enumerate(batches):
        torch.manual_seed(100 + i)
        slot = pipe.submit(r, t)
        if i >= 1:                                 # read batch i-1 while batch i is in flight
            pipe.wait(prev)
            got.append({k: v.clone() for k, v in pipe.host[prev].items()})
        prev = slot
    pipe.wait()
    got.append({k: v.clone() for k, v in pipe.host[prev].items()})
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert set(g) == set(w)
        for k in w:
            assert g[k].shape == w[k].shape and torch.equal(g[k], w[k]), k
    # a subset of the keys
    pipe2 = HostPipeline(ms, args, "cuda", depth=1, keys=("rgb_coarse", "depth_coarse"))
    torch.manual_seed(100)
    s = pipe2.submit(*batches[0])
    pipe2.wait()
    assert set(pipe2.host[s]) == {"rgb_coarse", "depth_coarse"} and torch.equal(pipe2.host[s]["depth_coarse"], want[0]["depth_coarse"])
    assert pipe2.d2h_bytes == batches[0][0].shape[0] * 16
