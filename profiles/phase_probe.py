"""Developer probe: per-GEMM phase breakdown (cycles) of the fused tcgen05 kernel, from in-kernel clock64 marks."""
import argparse, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import satnerf_b200 as sb
from satnerf_b200 import capi
from oracle import render_oracle as orc

a = argparse.Namespace(model="sat-nerf", n_samples=64, n_importance=0, noise_std=0.0, sc_lambda=0.0, chunk=1 << 20, fc_layers=8,
                       fc_units=int(sys.argv[1]) if len(sys.argv) > 1 else 512, t_embbeding_tau=4, t_embbeding_vocab=30, precision="tc")
torch.manual_seed(0)
ms = {"coarse": sb.load_model(a).cuda(), "t": torch.nn.Embedding(30, 4).cuda()}
train = len(sys.argv) > 2 and sys.argv[2] == "train"      # training-mode forward (activation stash) on 1024 rays
rays, ts = orc.synthetic_sat_rays(1024 if train else 4096, seed=1)
rays, ts = rays.cuda(), ts.cuda()
with torch.enable_grad() if train else torch.no_grad():
    for _ in range(3):
        sb.render_rays(ms, a, rays, ts)
torch.cuda.synchronize()
from satnerf_b200 import capi_dev
t = capi_dev.debug_timestamps()
print("layer0:", t[60, 1] - t[60, 0])
tot_wait = tot_epi = 0
for g in range(12):
    wait, epi = t[g, 1] - t[g, 0], t[g, 2] - t[g, 1]
    tot_wait += wait; tot_epi += epi
    print(f"gemm {g:2d}: wait-for-acc {wait:7d}  epilogue {epi:7d}")
print("sum wait", tot_wait, "sum epi", tot_epi, "composite", t[61, 1] - t[61, 0], "tile total", t[11, 2] - t[60, 0])

