set -x
timeout 300 python -m pytest tests/test_gpu_gram.py tests/test_gpu_fp32.py -m gpu -q 2>&1 | tail -6 > gpurun_out/r12_pytest_closed.log
B200DA_TC_WTABLE=1 timeout 300 python -m pytest tests/test_gpu_gram.py tests/test_gpu_fp32.py -m gpu -q 2>&1 | tail -6 > gpurun_out/r12_pytest_table.log
timeout 200 python tools/run_once.py --workload cfg3 --dtype f32 --repeat 2 > gpurun_out/r12_cfg3_f32_closed.log 2>&1
B200DA_TC_WTABLE=1 timeout 200 python tools/run_once.py --workload cfg3 --dtype f32 --repeat 2 > gpurun_out/r12_cfg3_f32_table.log 2>&1
