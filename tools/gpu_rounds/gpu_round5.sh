set -x
timeout 300 python -m pytest tests/test_gpu_gram.py tests/test_gpu_fp32.py -m gpu -q -x 2>&1 | tail -30 > gpurun_out/r5_pytest.log
timeout 200 python tools/run_once.py --workload cfg3 --dtype f32 --repeat 2 > gpurun_out/r5_cfg3_f32.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_tc_gram -c 1 -o gpurun_out/r5_tc python tools/run_once.py --workload cfg3 --dtype f32 --blocks 0:150 > gpurun_out/r5_ncu_tc.log 2>&1
