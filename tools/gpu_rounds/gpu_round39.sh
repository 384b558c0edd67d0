set -x
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r39_launches_widened.csv python tools/bench_ketkf.py --only cfg2 --steps 1 --warmup 0 > gpurun_out/r39.log 2>&1
tail -2 gpurun_out/r39.log
