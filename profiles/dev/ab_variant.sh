#!/bin/bash
# A/B of a developer variant (profiles/dev/variants/$1.so) against the dev library: headline forward + training step.
cd "$(dirname "$0")/../.."
for lib in satnerf_b200/libsatnerf_b200_dev.so profiles/dev/variants/$1.so; do
  export SNB_LIBRARY_PATH=$PWD/$lib SNB_DEV_LIBRARY_PATH=$PWD/$lib
  echo "== $lib"
  timeout 300 python bench.py --only train --no-cpu 2>/dev/null | python profiles/dev/bench_line.py
done
