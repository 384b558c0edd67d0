set -x
timeout 300 python -m pytest tests/test_gpu_gram.py tests/test_gpu_fp32.py tests/test_gpu_interface.py -m gpu -q 2>&1 | tail -12 > gpurun_out/r11_pytest.log
timeout 200 python tools/run_once.py --workload cfg3 --dtype f32 --repeat 2 > gpurun_out/r11_cfg3_f32.log 2>&1
