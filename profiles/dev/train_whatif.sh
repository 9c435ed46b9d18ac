#!/bin/bash
# What-if timings of the 1024-ray training step with parts of the stash traffic switched off (dev library knobs).
cd "$(dirname "$0")/../.."
export SNB_LIBRARY_PATH=$PWD/satnerf_b200/libsatnerf_b200_dev.so
for dbg in 0 128 256 384 512 1024 1536 1920; do
  echo -n "SNB_TC_DBG=$dbg  "
  SNB_TC_DBG=$dbg timeout 200 python profiles/train_probe.py 10 2>&1 | tail -1 | grep -o "'ms_per_step': [0-9.]*"
done
