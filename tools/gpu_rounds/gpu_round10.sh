set -x
timeout 600 python -m pytest tests -m gpu -q -k "etkf or core or fixture" 2>&1 | tail -8 > gpurun_out/r10_pytest.log
timeout 300 python tools/bench_etkf.py --dtype f64 > gpurun_out/r10_etkf_f64.json 2> gpurun_out/r10_etkf_f64.err
timeout 300 python tools/bench_etkf.py --dtype f32 > gpurun_out/r10_etkf_f32.json 2> gpurun_out/r10_etkf_f32.err
