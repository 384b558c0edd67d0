set -x
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_apply_global -c 1 -o gpurun_out/r23_apply python tools/bench_etkf.py --steps 1 --warmup 0 --n-grid 4000000 --n-obs 400000 > gpurun_out/r23_ncu_apply.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_letkf_gram -c 1 -o gpurun_out/r23_gram python tools/run_once.py --workload cfg3 --blocks 0:2000 > gpurun_out/r23_ncu_gram.log 2>&1
ls -la gpurun_out/*.ncu-rep
