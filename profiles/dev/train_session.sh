#!/bin/bash
# Training-path GPU session: backward parity tests, the 1024-ray training probe (timed), then its launch list under ncu.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "gradient or fused_loss or backward or cta_modes or stash" > gpurun_out/train_pytest.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/train_pytest.log
timeout 300 python profiles/train_probe.py 20 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/train_launches.csv python profiles/train_probe.py 3 > gpurun_out/train_under_ncu.log 2>&1
python profiles/dev/agg_launches.py gpurun_out/train_launches.csv 14
