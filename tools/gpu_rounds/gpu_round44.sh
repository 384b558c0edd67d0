set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_ienks.py tests/test_gpu_kernels.py -m gpu -q -k "six_chained or validation" 2>&1 | grep -E "^E  |FAILED|passed|failed|Error|error|Mismatch|Max" | head -80 ) > gpurun_out/r44_pytest_new.log 2>&1
cat gpurun_out/r44_pytest_new.log
