set -x
timeout 600 python -m pytest tests/test_gpu_fp32.py tests/test_gpu_interface.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_pytest_fp32.log
timeout 300 python bench.py --dtype f32 --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r2_bench_cfg3_f32.json 2> gpurun_out/r2_bench_cfg3_f32.err
