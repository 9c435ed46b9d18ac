#!/bin/bash
# usage: build_variant.sh NAME "-DFLAG ..."   -> profiles/dev/variants/NAME.so (developer A/B builds of the DEV library;
# use with SNB_LIBRARY_PATH=... SNB_DEV_LIBRARY_PATH=... so that both bindings load the variant)
set -e
cd "$(dirname "$0")/../.."
mkdir -p profiles/dev/variants /tmp/snbv_$1
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -DSNB_DEV_BUILD -I include -I satnerf_b200/csrc $2"
for s in layout sampling composite simt_field tc_field tc_backward tc_bwd geo optim mma_rate capi; do
  /usr/local/cuda/bin/nvcc $F -c satnerf_b200/csrc/$s.cu -o /tmp/snbv_$1/$s.o &
done
wait
/usr/local/cuda/bin/nvcc -shared -o profiles/dev/variants/$1.so /tmp/snbv_$1/*.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC
echo built profiles/dev/variants/$1.so
