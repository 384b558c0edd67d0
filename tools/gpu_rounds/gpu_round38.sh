set -x
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:k_ienks_pre' -c 1 -f -o gpurun_out/r38_ienks_pre python tools/bench_ketkf.py --only cfg2 --steps 1 --warmup 0 > gpurun_out/r38_ncu.log 2>&1
tail -2 gpurun_out/r38_ncu.log
