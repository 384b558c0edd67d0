set -x
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tc_gram -c 1 -o gpurun_out/r13_tc python tools/run_once.py --workload cfg3 --dtype f32 --blocks 0:150 > gpurun_out/r13_ncu_tc.log 2>&1
