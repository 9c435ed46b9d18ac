#!/bin/bash
# One GPU session: parity tests, bench, launch list.  Usage: gpu_session.sh [pytest args...]
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.json
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=15 "$@" > gpurun_out/pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"
tail -c 3000 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench.json").read().strip().splitlines()[-1])
    keep = {k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}
    keep["e2e"] = d["e2e"]["value"]; keep["frac"] = d["roofline"]["frac"]; keep["kernel_ms"] = d["roofline"]["kernel_ms"]
    for k in ("train", "config3", "config4", "config5", "gpu_eager", "cpu_baseline"):
        v = d.get(k)
        if v: keep[k] = {kk: vv for kk, vv in v.items() if kk in ("value", "ms_per_step", "gpu_launches", "train", "depth", "full", "e2e")}
    print(json.dumps(keep, indent=1)[:6000])
except Exception as e:
    print("bench parse failed", e)
PY
