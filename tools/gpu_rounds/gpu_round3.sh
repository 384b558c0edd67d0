set -x
timeout 300 python -m pytest tests/test_gpu_gram.py -m gpu -q -x 2>&1 | tail -40 > gpurun_out/r3_pytest_gram.log
timeout 300 python -m pytest tests/test_gpu_fp32.py -m gpu -q 2>&1 | tail -40 > gpurun_out/r3_pytest_fp32.log
timeout 120 python tools/run_once.py --workload small --dtype f32 --repeat 2 > gpurun_out/r3_small_f32.log 2>&1
timeout 200 python tools/run_once.py --workload cfg3 --dtype f32 --repeat 2 > gpurun_out/r3_cfg3_f32.log 2>&1
