"""Developer probe: CTA-pair (cta_group::2) weight ring in isolation: cycles per M256 x N x K16 tcgen05.mma.
usage: mma_ring2_probe.py FLAGS [N]   (one configuration per process: a trap poisons the context)"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from satnerf_b200 import capi
torch.zeros(1, device="cuda")
from satnerf_b200 import capi_dev
lib = capi_dev.lib()          # microbenchmarks live in libsatnerf_b200_dev.so
out = (C.c_longlong * 2)()
flags = int(sys.argv[1]); N = int(sys.argv[2]) if len(sys.argv) > 2 else 256
row = []
for depth in (4, 6):
    rc = lib.snb_debug_mma_ring2(N, depth, 1024, flags, 148, out)
    if rc:
        print("flags", flags, "error:", lib.snb_last_error().decode()); break
    row.append(f"{out[0] / out[1]:6.1f}")
print(f"N={N} flags={flags:4d} depth 4,6: " + " ".join(row), flush=True)
