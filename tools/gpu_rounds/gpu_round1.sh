set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1_pytest.log
python bench.py > gpurun_out/r1_bench_cfg3.json 2> gpurun_out/r1_bench_cfg3.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_cfg3.csv python bench.py --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline > gpurun_out/r1_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_letkf_gram -c 1 -o gpurun_out/r1_gram python tools/run_once.py --workload cfg3 --blocks 0:2000 > gpurun_out/r1_ncu_gram.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_letkf_solve_ns -c 1 -o gpurun_out/r1_ns python tools/run_once.py --workload cfg3 --blocks 0:2000 > gpurun_out/r1_ncu_ns.log 2>&1
python bench.py --workload cfg2 > gpurun_out/r1_bench_cfg2.json 2> gpurun_out/r1_bench_cfg2.err
