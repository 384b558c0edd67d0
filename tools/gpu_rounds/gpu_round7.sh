set -x
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cfg4" 2>&1 | tail -15 > gpurun_out/r7_pytest_cfg4.log
timeout 300 python tools/bench_etkf.py --dtype f64 > gpurun_out/r7_etkf_f64.json 2> gpurun_out/r7_etkf_f64.err
timeout 300 python tools/bench_etkf.py --dtype f32 > gpurun_out/r7_etkf_f32.json 2> gpurun_out/r7_etkf_f32.err
timeout 900 python tools/sweep_cfg5.py > gpurun_out/r7_sweep_cfg5.jsonl 2> gpurun_out/r7_sweep_cfg5.err
