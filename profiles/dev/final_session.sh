#!/bin/bash
# Round-end GPU session: full parity suite, default bench, launch list of one bench pass, ncu --set full of the hot kernels.
# usage (under gpurun): bash profiles/dev/final_session.sh NAME
cd "$(dirname "$0")/../.."
NAME=${1:-r2z}
mkdir -p gpurun_out
rm -f gpurun_out/parity_report.json
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/${NAME}_pytest.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/${NAME}_pytest.log
timeout 400 python bench.py > gpurun_out/${NAME}_bench.json 2> gpurun_out/${NAME}_bench.err
echo "bench exit $?"
# launch list: headline steps + training steps (the kernel's share of the step)
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${NAME}_launches.csv \
    python bench.py --steps 3 --warmup 3 --only train --no-cpu > gpurun_out/${NAME}_bench_under_ncu.log 2>&1
# ncu --set full: inference kernel (config 2), the three training kernels, the hi+lo layer kernel
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'tc_render_kernel' --launch-skip 3 --launch-count 1 \
    -f -o gpurun_out/${NAME}_fwd python bench.py --steps 2 --warmup 3 --only none --no-train --no-cpu > gpurun_out/${NAME}_fwd_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'tc_render_kernel|tc_chain_kernel|tc_dw_kernel' --launch-skip 6 --launch-count 3 \
    -f -o gpurun_out/${NAME}_train python profiles/train_probe.py 2 > gpurun_out/${NAME}_train_ncu.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'linear_fwd_x3' --launch-skip 26 --launch-count 2 \
    -f -o gpurun_out/${NAME}_x3 python profiles/dev/x3_probe.py > gpurun_out/${NAME}_x3_ncu.log 2>&1
ls -la gpurun_out/${NAME}_*
