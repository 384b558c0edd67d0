set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed|Error" | head -60 ) > gpurun_out/r31_pytest.log 2>&1
python bench.py --steps 2 --warmup 3 > gpurun_out/r31_bench_cfg3_f64.json 2> gpurun_out/r31_bench_cfg3_f64.err
cat gpurun_out/r31_pytest.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/r31_bench_cfg3_f64.json")); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"])
PY
