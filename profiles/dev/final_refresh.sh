cd /root/repo
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2z_pytest.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/r2z_pytest.log
timeout 400 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err; echo "bench exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'linear_fwd_x3' --launch-skip 26 --launch-count 2 -f -o gpurun_out/r2z_x3 python profiles/dev/x3_probe.py > gpurun_out/r2z_x3_ncu.log 2>&1
tail -c 400 gpurun_out/r2z_bench.json
