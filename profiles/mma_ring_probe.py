"""Developer probe: the weight ring in isolation (cycles per M128 x N x K16 tcgen05.mma as a function of stage
granularity, ring depth, and whether the stages are refilled by real bulk copies)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from satnerf_b200 import capi
torch.zeros(1, device="cuda")
from satnerf_b200 import capi_dev
lib = capi_dev.lib()          # microbenchmarks live in libsatnerf_b200_dev.so
out = (C.c_longlong * 2)()
def run(N, kstage, depth, flags, blocks=148, total_mma=4096):
    groups = total_mma * 16 // kstage
    rc = lib.snb_debug_mma_ring(N, kstage, depth, groups, flags, blocks, out)
    if rc or out[1] == 0:
        return None
    return out[0] / out[1]
print("flags: 1=random data 2=real copies 4=serialised (commit->wait every stage)")
for N in (256, 128):
    for flags in (1, 5, 3, 2):
        for kstage in (64, 32, 16):
            row = []
            for depth in (1, 2, 3, 4, 6, 8):
                if 131072 + depth * N * kstage * 2 + N * (128 - 2 * kstage) > 227 * 1024 - 2048:
                    row.append("   -  "); continue
                c = run(N, kstage, depth, flags)
                row.append("  err " if c is None else f"{c:6.1f}")
            print(f"N={N:3d} flags={flags} kstage={kstage:2d} | depth 1,2,3,4,6,8: " + " ".join(row), flush=True)
