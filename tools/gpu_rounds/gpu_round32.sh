set -x
mkdir -p gpurun_out
( timeout 420 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_interface.py -m gpu -q -x 2>&1 | grep -E "^E  |FAILED|passed|failed|Error|error" | head -60 ) > gpurun_out/r32_pytest_new.log 2>&1
cat gpurun_out/r32_pytest_new.log
( timeout 200 python tools/bench_ketkf.py --steps 5 --warmup 3 > gpurun_out/r32_bench_ketkf.jsonl ) 2> gpurun_out/r32_bench_ketkf.err
cat gpurun_out/r32_bench_ketkf.jsonl; tail -5 gpurun_out/r32_bench_ketkf.err
