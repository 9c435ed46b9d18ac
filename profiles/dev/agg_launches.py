"""Aggregates an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr, agg = None, collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(d["Metric Value"].replace(",", ""))
    if d["Metric Unit"] in ("us", "usecond"):
        v *= 1e3
    elif d["Metric Unit"] in ("ms", "msecond"):
        v *= 1e6
    a = agg.setdefault(d["Kernel Name"][:90], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(t for _, t in agg.values())
print(f"{'n':>5} {'total us':>10} {'avg us':>9} {'share':>6}  kernel")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]:
    print(f"{n:5d} {t / 1e3:10.1f} {t / 1e3 / n:9.1f} {100 * t / tot:5.1f}%  {k}")
