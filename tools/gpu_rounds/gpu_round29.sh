set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed|Error" | head -40 ) > gpurun_out/r29_pytest.log 2>&1
python bench.py --workload cfg2 > gpurun_out/r29_bench_cfg2_f64.json 2> gpurun_out/r29_bench_cfg2_f64.err
timeout 900 python tools/sweep_cfg5.py > gpurun_out/r29_sweep_cfg5.jsonl 2> gpurun_out/r29_sweep_cfg5.err
cat gpurun_out/r29_pytest.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/r29_bench_cfg2_f64.json")); print(d["value"], d["ms_per_step"], d["roofline"]["solve_kernel"]["kernel_ms"])
for l in open("gpurun_out/r29_sweep_cfg5.jsonl"):
    if l.startswith("{"):
        d=json.loads(l); print({k:d[k] for k in d if k in ("k","dtype","value","gridpoints_per_s","ms_per_step","gram_ms","solve_ms","fraction")})
PY
