set -x
ncu --set full --clock-control none --import-source on -k regex:k_tc_gram -c 1 -o gpurun_out/r4_tc python tools/run_once.py --workload small --dtype f32 > gpurun_out/r4_ncu_tc.log 2>&1
