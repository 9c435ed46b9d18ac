import json, sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
for line in sys.stdin:
    line = line.strip()
    if line.startswith("{"):
        d = json.loads(line)
        r = d.get("roofline", {})
        print(tag, "rays/s %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"], "kernel_ms %.4f" % r.get("kernel_ms", 0), "frac %.4f" % r.get("frac", 0),
              "train", (d.get("train") or {}).get("ms_per_step"), d.get("clocks"))
