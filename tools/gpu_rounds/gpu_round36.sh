set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed|Error" | head -60 ) > gpurun_out/r36_pytest.log 2>&1
cat gpurun_out/r36_pytest.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r36_bench_cfg3_f64.json 2> gpurun_out/r36_bench_cfg3_f64.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/r36_bench_cfg3_f64.json")); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"], d["clocks"])
PY
timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:k_kernelise' -c 1 -f -o gpurun_out/r36_kernelise python tools/bench_ketkf.py --only cfg2 --steps 1 --warmup 0 > gpurun_out/r36_ncu.log 2>&1
tail -2 gpurun_out/r36_ncu.log
