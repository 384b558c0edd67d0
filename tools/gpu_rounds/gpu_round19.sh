set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed|Error" | head -60 ) > gpurun_out/r19_pytest.log 2>&1
python tools/diag_stiff.py > gpurun_out/r19_diag.log 2>&1
python tools/diag_etkf.py 1000000 >> gpurun_out/r19_diag.log 2>&1
python tools/bench_etkf.py > gpurun_out/r19_etkf_f64.json 2> gpurun_out/r19_etkf_f64.err
python bench.py --workload cfg2 > gpurun_out/r19_bench_cfg2_f64.json 2> gpurun_out/r19_bench_cfg2_f64.err
cat gpurun_out/r19_pytest.log; cat gpurun_out/r19_diag.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/r19_etkf_f64.json")); print(d["ms_per_step"], d["weights"]["ms"], d["update"]["ms"])
d=json.load(open("gpurun_out/r19_bench_cfg2_f64.json")); print(d["value"], d["ms_per_step"], d["roofline"]["solve_kernel"]["kernel_ms"])
PY
