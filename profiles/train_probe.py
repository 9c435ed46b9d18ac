"""Developer probe: a few training steps (1024 rays x 64 samples, sat-nerf h=512) for `ncu --metrics gpu__time_duration.sum`."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from satnerf_b200 import train as trn

a = argparse.Namespace(model="sat-nerf", n_samples=64, n_importance=0, noise_std=0.0, sc_lambda=0.0, chunk=1 << 20, fc_layers=8,
                       fc_units=512, t_embbeding_tau=4, t_embbeding_vocab=30, precision="tc", lr=5e-4)
flush = torch.empty(1 << 20, dtype=torch.uint8, device="cuda")
print(trn.bench_training_step(a, torch.device("cuda", 0), 0, 1, 1024, 2, int(sys.argv[1]) if len(sys.argv) > 1 else 2, flush))
