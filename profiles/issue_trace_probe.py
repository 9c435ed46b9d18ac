"""Developer probe (needs the -DSNB_TC_PROBE build: SNB_LIBRARY_PATH=profiles/dev/variants/probe.so): timeline of the MMA
issuer for GEMMs 1 and 2 of block 0's second tile: per issue region, cycles spent waiting for the weight stages / ready
signals and cycles spent issuing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import argparse, torch
import satnerf_b200 as sb
from satnerf_b200 import capi
from oracle import render_oracle as orc
a = argparse.Namespace(model="sat-nerf", n_samples=64, n_importance=0, noise_std=0.0, sc_lambda=0.0, chunk=1 << 20, fc_layers=8,
                       fc_units=512, t_embbeding_tau=4, t_embbeding_vocab=30, precision="tc")
torch.manual_seed(0)
ms = {"coarse": sb.load_model(a).cuda(), "t": torch.nn.Embedding(30, 4).cuda()}
rays, ts = orc.synthetic_sat_rays(4096, seed=1)
rays, ts = rays.cuda(), ts.cuda()
with torch.no_grad():
    for _ in range(3):
        sb.render_rays(ms, a, rays, ts)
torch.cuda.synchronize()
from satnerf_b200 import capi_dev
t = capi_dev.debug_timestamps()
t0 = t[20, 0]
prev_end = t0
for n in range(40):
    b, w, e, tag = (int(x) for x in t[20 + n])
    if e == 0:
        break
    print(f"gemm {tag >> 16} stage {(tag >> 8) & 255:2d} pair {tag & 1}: start {b - t0:6d}  gap {b - prev_end:5d}  wait {w - b:5d}  issue {e - w:5d}")
    prev_end = e
for g in (0, 1, 2, 3):
    print(f"gemm {g}: epilogue starts waiting {int(t[g,0]-t0):6d}  chunk0 ready {int(t[g,1]-t0):6d}  chunk1 ready {int(t[g,3]-t0):6d}  epilogue done {int(t[g,2]-t0):6d}")
print("gemm 1 final chunk, thread 0: acc barrier passed", int(t[53,0]-t0), " named barrier 2 passed", int(t[53,1]-t0), " block loop done", int(t[52,0]-t0),
      " tcgen05 fence", int(t[52,1]-t0), " proxy fence", int(t[52,2]-t0), " named barrier 1 passed", int(t[1,2]-t0))
