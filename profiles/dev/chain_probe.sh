cd /root/repo
export SNB_LIBRARY_PATH=$PWD/satnerf_b200/libsatnerf_b200_dev.so
SNB_TC_DBG=4096 timeout 200 python profiles/train_probe.py 1 2>&1 | grep -A14 "chain tile probe" | head -16
