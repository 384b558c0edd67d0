set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed|Error" | head -60 ) > gpurun_out/r42_pytest.log 2>&1
cat gpurun_out/r42_pytest.log
python bench.py --workload cfg1 --steps 20 --warmup 5 > gpurun_out/r42_bench_cfg1_f64.json 2> gpurun_out/r42_bench_cfg1.err
python bench.py --workload cfg1 --impl reference --steps 3 --warmup 1 > gpurun_out/r42_bench_cfg1_reference.json 2>> gpurun_out/r42_bench_cfg1.err
python bench.py --steps 3 --warmup 3 > gpurun_out/r42_bench_cfg3_f64.json 2> gpurun_out/r42_bench_cfg3_f64.err
python - <<'PY'
import json
for f in ("r42_bench_cfg1_f64","r42_bench_cfg1_reference","r42_bench_cfg3_f64"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, d["value"], d["ms_per_step"], d.get("e2e",{}).get("value"), d.get("gpu_launches"), d.get("clocks"))
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/r42_bench_cfg1.err
