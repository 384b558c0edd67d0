set -x
mkdir -p gpurun_out
( timeout 100 python -m pytest tests/test_gpu_ienks.py -m gpu -q -k "multi_chunk" --durations=1 2>&1 | grep -E "^E  |FAILED|passed|failed|Error|error|call " | head -40 ) > gpurun_out/r47_pytest_new.log 2>&1
cat gpurun_out/r47_pytest_new.log
