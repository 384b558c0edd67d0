set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r8_pytest.log
timeout 900 python tools/sweep_cfg5.py --ks 64,128 > gpurun_out/r8_sweep_cfg5b.jsonl 2> gpurun_out/r8_sweep_cfg5b.err
