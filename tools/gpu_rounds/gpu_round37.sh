set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_ienks.py -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed|Error|error" | head -80 ) > gpurun_out/r37_pytest_new.log 2>&1
cat gpurun_out/r37_pytest_new.log
( timeout 300 python tools/bench_ketkf.py --steps 5 --warmup 3 > gpurun_out/r37_bench_widened.jsonl ) 2> gpurun_out/r37_bench_widened.err
cat gpurun_out/r37_bench_widened.jsonl; tail -5 gpurun_out/r37_bench_widened.err
