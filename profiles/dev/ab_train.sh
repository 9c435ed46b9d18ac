#!/bin/bash
# A/B of developer variants on the 1024-ray training step: ab_train.sh NAME...   (profiles/dev/variants/NAME.so)
cd "$(dirname "$0")/../.."
for v in dev "$@"; do
  lib=profiles/dev/variants/$v.so; [ "$v" = dev ] && lib=satnerf_b200/libsatnerf_b200_dev.so
  export SNB_LIBRARY_PATH=$PWD/$lib SNB_DEV_LIBRARY_PATH=$PWD/$lib
  echo -n "$v  "; timeout 200 python profiles/train_probe.py 10 2>&1 | tail -1 | grep -o "'ms_per_step': [0-9.]*"
done
