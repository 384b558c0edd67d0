mkdir -p gpurun_out
( timeout 45 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed|Error" | head -20 ) > gpurun_out/r49_pytest.log 2>&1
cat gpurun_out/r49_pytest.log
