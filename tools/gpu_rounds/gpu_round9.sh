set -x
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r9_launches_etkf.csv python tools/bench_etkf.py --n-grid 1000000 --steps 1 --warmup 1 > gpurun_out/r9_etkf.log 2>&1
