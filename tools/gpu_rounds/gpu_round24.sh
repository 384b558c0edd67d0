set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q 2>&1 | grep -E "^E  |FAILED|passed|failed|Error" | head -40 ) > gpurun_out/r24_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r24_smoke.log 2>&1
python tools/bench_etkf.py > gpurun_out/r24_etkf_f64.json 2> gpurun_out/r24_etkf_f64.err
python tools/bench_etkf.py --dtype f32 > gpurun_out/r24_etkf_f32.json 2> gpurun_out/r24_etkf_f32.err
python bench.py > gpurun_out/r24_bench_cfg3_f64.json 2> gpurun_out/r24_bench_cfg3_f64.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r24_launches_cfg3.csv python bench.py --steps 1 --warmup 1 > gpurun_out/r24_ncu_bench.log 2>&1
cat gpurun_out/r24_pytest.log; tail -1 gpurun_out/r24_smoke.log
python - <<'PY'
import json
for f in ("r24_etkf_f64","r24_etkf_f32"):
    d=json.load(open("gpurun_out/%s.json"%f)); print(f, d["ms_per_step"], d["weights"]["ms"], d["update"]["ms"])
d=json.load(open("gpurun_out/r24_bench_cfg3_f64.json")); print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"])
PY
