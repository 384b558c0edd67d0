set -x
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 ) > gpurun_out/r15_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r15_smoke.log 2>&1
python tools/bench_etkf.py > gpurun_out/r15_etkf_f64.json 2> gpurun_out/r15_etkf_f64.err
python tools/bench_etkf.py --dtype f32 > gpurun_out/r15_etkf_f32.json 2> gpurun_out/r15_etkf_f32.err
python bench.py --workload cfg2 > gpurun_out/r15_bench_cfg2_f64.json 2> gpurun_out/r15_bench_cfg2_f64.err
( time python bench.py > gpurun_out/r15_bench_cfg3_f64.json 2> gpurun_out/r15_bench_cfg3_f64.err ) 2> gpurun_out/r15_bench_time.log
tail -3 gpurun_out/r15_pytest.log; cat gpurun_out/r15_smoke.log | tail -2; cat gpurun_out/r15_etkf_f64.json
