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
    return None if rc or out[1] == 0 else out[0] / out[1]
for N in (256, 128, 64):
    for kstage in (64, 16):
        for flags, name in ((1, "base"), (1+8+16+32+64, "nothing"), (1+8+16+32+64+128, "nothing, no syncwarp"), (1+8+16+32+64+128+256, "nothing, no syncwarp, no lane0 block"),
                            (1+8+16+64+128+256, "commit each, no sync/lane0"), (1+512, "base, acc0 first")):
            for blocks in (1, 148):
                print(f"N={N} kstage={kstage} depth=2 blocks={blocks} {name:40s}: {run(N, kstage, 2, flags, blocks):.1f}", flush=True)
