set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed|Error" | head -40 ) > gpurun_out/r30_pytest.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_letkf_solve_ns -c 1 -o gpurun_out/r30_solve python tools/run_once.py --workload cfg3 --blocks 0:2000 > gpurun_out/r30_ncu_solve.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_etkf_gram -c 1 -o gpurun_out/r30_egram python tools/bench_etkf.py --steps 1 --warmup 0 > gpurun_out/r30_ncu_egram.log 2>&1
cat gpurun_out/r30_pytest.log
