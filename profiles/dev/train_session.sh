#!/bin/bash
# Training-path GPU session: backward parity tests, the 1024-ray training probe (timed), its launch list under ncu, and
# (with NCU=regex) one `ncu --set full` capture of the matching training kernels -> gpurun_out/${NAME}_train.ncu-rep
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
NAME=${NAME:-r2}
timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "gradient or fused_loss or backward or cta_modes or stash or adam" > gpurun_out/train_pytest.log 2>&1
echo "pytest exit $?"; tail -15 gpurun_out/train_pytest.log
timeout 300 python profiles/train_probe.py 20 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/train_launches.csv python profiles/train_probe.py 3 > gpurun_out/train_under_ncu.log 2>&1
python profiles/dev/agg_launches.py gpurun_out/train_launches.csv 14
if [ -n "$NCU" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$NCU" --launch-skip ${SKIP:-2} --launch-count ${COUNT:-1} \
      -f -o gpurun_out/${NAME}_train python profiles/train_probe.py 2 > gpurun_out/${NAME}_train_ncu.log 2>&1
  ls -la gpurun_out/${NAME}_train.ncu-rep
fi
