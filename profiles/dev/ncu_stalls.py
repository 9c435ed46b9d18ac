"""Aggregates the SASS source page of an ncu report (ncu -i X --page source --csv --kernel-name regex:K > f.csv):
total samples per stall reason and the top-N instructions by samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]; body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
col = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
tot = {n: sum(int(r[col[n]] or 0) for r in body) for n in stalls}
allsamp = sum(int(r[col["# Samples"]] or 0) for r in body)
print("total samples", allsamp, " instructions executed (warp)", sum(int(r[col["Instructions Executed"]] or 0) for r in body))
for n, v in sorted(tot.items(), key=lambda x: -x[1])[:10]:
    print(f"  {n:28s} {v:8d} {100.0 * v / max(allsamp, 1):5.1f}%")
idx = sorted(range(len(body)), key=lambda i: -int(body[i][col["# Samples"]] or 0))[:top_n]
for i in idx:
    r = body[i]
    why = sorted(((int(r[col[n]] or 0), n) for n in stalls), reverse=True)[:2]
    print(f"{int(r[col['# Samples']]):7d}  line {i:5d}  {r[col['Source']].strip()[:90]:90s} {why[0][1]}={why[0][0]} {why[1][1]}={why[1][0]}")
