set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -k "etkf or interface or fixture" 2>&1 | grep -E "^E  |FAILED|passed|failed|Error" | head -40 ) > gpurun_out/r21_pytest.log 2>&1
python tools/bench_etkf.py > gpurun_out/r21_etkf_f64.json 2> gpurun_out/r21_etkf_f64.err
python tools/bench_etkf.py --dtype f32 > gpurun_out/r21_etkf_f32.json 2> gpurun_out/r21_etkf_f32.err
cat gpurun_out/r21_pytest.log
python - <<'PY'
import json
for f in ("r21_etkf_f64","r21_etkf_f32"):
    d=json.load(open("gpurun_out/%s.json"%f)); print(f, d["ms_per_step"], d["weights"]["ms"], d["update"]["ms"])
PY
