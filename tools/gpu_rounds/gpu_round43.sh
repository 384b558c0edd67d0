set -x
mkdir -p gpurun_out
python bench.py --workload cfg1 --steps 20 --warmup 5 > gpurun_out/r43_bench_cfg1_f64.json 2> gpurun_out/r43_bench_cfg1.err
python bench.py --workload cfg1 --impl reference --steps 3 --warmup 1 > gpurun_out/r43_bench_cfg1_reference.json 2>> gpurun_out/r43_bench_cfg1.err
timeout 120 python tools/bench_interface.py --nlat 300 --nlon 300 --n-obs 225000 --steps 2 > gpurun_out/r43_bench_interface_small.json 2> gpurun_out/r43_bench_interface.err
python - <<'PY'
import json
for f in ("r43_bench_cfg1_f64","r43_bench_cfg1_reference"):
    d=json.load(open("gpurun_out/%s.json"%f)); print(f, d["value"], d["ms_per_step"], d["cpu_baseline"]["value"], d["cpu_baseline"]["sample"][:60])
PY
cat gpurun_out/r43_bench_interface_small.json | cut -c1-600; tail -3 gpurun_out/r43_bench_interface.err
