set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_core_modules.py tests/test_gpu_interface.py -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed|Error|error" | head -80 ) > gpurun_out/r40_pytest_new.log 2>&1
cat gpurun_out/r40_pytest_new.log
