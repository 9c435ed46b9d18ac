#!/bin/bash
# ncu --set full of the three training kernels (one launch each, after the warm-up steps) -> gpurun_out/$1_train.ncu-rep
cd "$(dirname "$0")/../.."
NAME=${1:-r2}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'tc_render_kernel|tc_chain_kernel|tc_dw_kernel' --launch-skip 6 --launch-count 3 \
    -f -o gpurun_out/${NAME}_train python profiles/train_probe.py 2 > gpurun_out/${NAME}_train_ncu.log 2>&1
tail -3 gpurun_out/${NAME}_train_ncu.log
ls -la gpurun_out/${NAME}_train.ncu-rep
