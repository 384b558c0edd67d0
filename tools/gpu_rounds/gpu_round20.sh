set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed|Error" | head -40 ) > gpurun_out/r20_pytest.log 2>&1
python bench.py --workload cfg2 > gpurun_out/r20_bench_cfg2_f64.json 2> gpurun_out/r20_bench_cfg2_f64.err
python bench.py --dtype f32 > gpurun_out/r20_bench_cfg3_f32.json 2> gpurun_out/r20_bench_cfg3_f32.err
python tools/bench_etkf.py --dtype f32 > gpurun_out/r20_etkf_f32.json 2> gpurun_out/r20_etkf_f32.err
cat gpurun_out/r20_pytest.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/r20_bench_cfg2_f64.json")); print(d["value"], d["ms_per_step"], d["roofline"]["solve_kernel"]["kernel_ms"])
d=json.load(open("gpurun_out/r20_bench_cfg3_f32.json")); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["solve_kernel"]["kernel_ms"])
d=json.load(open("gpurun_out/r20_etkf_f32.json")); print(d["ms_per_step"], d["weights"]["ms"], d["update"]["ms"])
PY
