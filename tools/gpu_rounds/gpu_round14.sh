set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r14_pytest.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fixture or cfg1 or kat or no_observation" > gpurun_out/r14_memcheck_f64.log 2>&1; echo "rc=$?" >> gpurun_out/r14_memcheck_f64.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests/test_gpu_fp32.py -m gpu -q -x -k "edge or cfg1" > gpurun_out/r14_memcheck_f32.log 2>&1; echo "rc=$?" >> gpurun_out/r14_memcheck_f32.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r14_smoke.log 2>&1
python bench.py > gpurun_out/r14_bench_cfg3_f64.json 2> gpurun_out/r14_bench_cfg3_f64.err
python bench.py --dtype f32 > gpurun_out/r14_bench_cfg3_f32.json 2> gpurun_out/r14_bench_cfg3_f32.err
python bench.py --workload cfg2 > gpurun_out/r14_bench_cfg2_f64.json 2> gpurun_out/r14_bench_cfg2_f64.err
