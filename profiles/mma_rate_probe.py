"""Developer probe: raw tcgen05.mma issue rate (cycles per instruction, MAC/clk/SM) for several shapes."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from satnerf_b200 import capi
torch.zeros(1, device="cuda")
from satnerf_b200 import capi_dev
lib = capi_dev.lib()          # microbenchmarks live in libsatnerf_b200_dev.so
out = (C.c_longlong * 2)()
names = {0: "SS cg1 K-major", 1: "TS cg1 (A in TMEM)", 2: "SS cg2 M=256", 3: "SS cg1 MN-major"}
for blocks, mode, N, it in [(b, m, n, i) for b in (1, 148) for m in (0, 1, 2, 3) for n in (64, 128, 256) for i in (0, 16)]:
    if True:
        if True:
            rc = lib.snb_debug_mma_rate(mode + it, N, 2048, blocks, out)
            names[mode] = names[mode].split(" |")[0] + (" | random data" if it else " | ones       ")
            if rc or out[0] == 0:
                print('mode', mode, 'N', N, 'rc', rc, lib.snb_last_error().decode()); continue
            cyc = out[0] / 2048
            M = 256 if mode == 2 else 128
            macs = M * N * 16 / cyc / (2 if mode == 2 else 1)
            print(f"blocks={blocks:3d} {names[mode]:36s} N={N:3d}: {cyc:7.1f} cyc/instr  {macs:7.0f} MAC/clk/SM")
