set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_core_modules.py -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed|Error|error" | head -80 ) > gpurun_out/r45_pytest_new.log 2>&1
cat gpurun_out/r45_pytest_new.log
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:k_kernelise' -c 1 -f -o gpurun_out/r45_kernelise python tools/bench_ketkf.py --only cfg2 --steps 3 --warmup 2 > gpurun_out/r45_bench.log 2>&1
grep lketkf gpurun_out/r45_bench.log | cut -c1-1200
