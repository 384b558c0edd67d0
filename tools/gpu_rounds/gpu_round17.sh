set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 ) > gpurun_out/r17_pytest.log 2>&1
python bench.py --steps 3 --warmup 3 > gpurun_out/r17_bench_cfg3_f64.json 2> gpurun_out/r17_bench_cfg3_f64.err
python bench.py --workload cfg2 > gpurun_out/r17_bench_cfg2_f64.json 2> gpurun_out/r17_bench_cfg2_f64.err
python tools/sweep_cfg5.py --help > gpurun_out/r17_sweep_help.txt 2>&1
tail -3 gpurun_out/r17_pytest.log
python - <<'PY'
import json
for f in ("r17_bench_cfg3_f64","r17_bench_cfg2_f64"):
    d=json.load(open("gpurun_out/%s.json"%f)); print(f, d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"])
PY
