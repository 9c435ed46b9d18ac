"""Opcode histogram per kernel of the in-tree library (cuobjdump -sass): evidence that the hot kernels are tcgen05 / TMEM / bulk-copy
code.  usage: python profiles/dev/sass_hist.py > profiles/r2_sass_opcodes.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
lib = os.path.join(ROOT, "satnerf_b200", "libsatnerf_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
KEY = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UBLKPF", "UTMALDG", "UTMASTG", "SYNCS", "MUFU", "HMMA", "ATOMG", "RED", "LDG", "STG", "LDS", "STS", "LDL", "STL", "BAR")
kern, hist = None, collections.OrderedDict()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip().split("(")[0]
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m and kern:
        hist[kern][m.group(1)] += 1
print("# opcode counts per kernel (static SASS of satnerf_b200/libsatnerf_b200.so); only the mnemonic families below are listed")
print("# " + " ".join(KEY))
for k, c in hist.items():
    tot = sum(c.values())
    fam = collections.Counter()
    for op, n in c.items():
        for key in KEY:
            if op.startswith(key):
                fam[key if key != "UTCHMMA" else op] += n
                break
    if not fam:
        continue
    print(f"\n{k}   [{tot} instructions]")
    print("   " + "  ".join(f"{op}={n}" for op, n in sorted(fam.items(), key=lambda x: -x[1])))
