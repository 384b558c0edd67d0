set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed|Error" | head -60 ) > gpurun_out/r46_pytest.log 2>&1
cat gpurun_out/r46_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
