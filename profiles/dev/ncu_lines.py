"""Joins an ncu SASS source page (ncu -i X --page source --csv --kernel-name regex:K) with `nvdisasm -g` of the cubin the
kernel came from, and prints warp-stall samples per CUDA source line.

  cuobjdump -xelf all lib.so; nvdisasm -g tc_bwd.sm_100a.cubin > dis.txt
  python ncu_lines.py page.csv dis.txt <mangled-kernel-substring> [top_n]
"""
import collections, csv, re, sys
page, dis, kname = sys.argv[1:4]
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
rows = list(csv.reader(open(page, errors="replace")))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]; body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
col = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
lines, cur, on = [], ("?", 0), False
for ln in open(dis, errors="replace"):
    if ln.startswith(".text."):
        on = kname in ln
        continue
    if not on:
        continue
    if ln.startswith("//-----"):
        on = False
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
print("sass instructions: ncu", len(body), "nvdisasm", len(lines))
n = min(len(body), len(lines))
agg = collections.defaultdict(lambda: collections.Counter())
inst = collections.Counter()
for i in range(n):
    r = body[i]
    for s in stalls:
        v = int(r[col[s]] or 0)
        if v:
            agg[lines[i]][s] += v
    inst[lines[i]] += int(r[col["Instructions Executed"]] or 0)
tot = sum(sum(c.values()) for c in agg.values())
print("samples", tot)
for key, c in sorted(agg.items(), key=lambda x: -sum(x[1].values()))[:top_n]:
    t = sum(c.values())
    why = ", ".join(f"{k[6:]}={v}" for k, v in c.most_common(3))
    print(f"{t:7d} {100.0 * t / tot:5.1f}%  {key[0]}:{key[1]:<5d} inst={inst[key]:>9d}  {why}")
