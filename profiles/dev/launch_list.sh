#!/bin/bash
# usage: launch_list.sh NAME <bench args...>  -> gpurun_out/NAME_launches.csv (+ aggregated summary on stdout)
cd "$(dirname "$0")/../.."
NAME=$1; shift
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${NAME}_launches.csv python bench.py "$@" > gpurun_out/${NAME}_bench_under_ncu.log 2>&1
python profiles/dev/agg_launches.py gpurun_out/${NAME}_launches.csv
