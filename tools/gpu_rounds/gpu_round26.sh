set -x
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stiff or sharded_etkf or cfg4" > gpurun_out/r26_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/r26_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "stiff_ensemble or sharded_etkf" > gpurun_out/r26_racecheck.log 2>&1; echo "rc=$?" >> gpurun_out/r26_racecheck.log
timeout 600 python tools/bench_interface.py > gpurun_out/r26_interface_f64.json 2> gpurun_out/r26_interface_f64.err
tail -4 gpurun_out/r26_memcheck.log; tail -4 gpurun_out/r26_racecheck.log; cat gpurun_out/r26_interface_f64.json; tail -3 gpurun_out/r26_interface_f64.err
