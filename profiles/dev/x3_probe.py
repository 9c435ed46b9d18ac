"""Developer probe: <Field>.forward on 8192 points (h=512) at points_precision fp32 / tcx3 -- timing and an ncu target.
    python profiles/dev/x3_probe.py [n_points]"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import satnerf_b200 as sb

B = int(sys.argv[1]) if len(sys.argv) > 1 else 9472
args = argparse.Namespace(model="sat-nerf", fc_layers=8, fc_units=512, t_embbeding_tau=4, t_embbeding_vocab=30)
torch.manual_seed(0)
m = sb.load_model(args).cuda()
xyz, sun, t = torch.rand(B, 3, device="cuda") * 2 - 1, torch.rand(B, 3, device="cuda"), torch.randn(B, 4, device="cuda")
with torch.no_grad():
    for prec in ("fp32", "tcx3"):
        m.points_precision = prec
        for _ in range(3):
            m(xyz, input_sun_dir=sun, input_t=t)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            m(xyz, input_sun_dir=sun, input_t=t)
        e1.record()
        torch.cuda.synchronize()
        print(prec, "ms per call", e0.elapsed_time(e1) / 10, "points", B)
