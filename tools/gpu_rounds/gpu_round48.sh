mkdir -p gpurun_out
timeout 80 python tools/bench_ketkf.py --only cfg3b --steps 2 --warmup 1 > gpurun_out/r48_bench_widened_cfg3b.jsonl 2> gpurun_out/r48.err
cat gpurun_out/r48_bench_widened_cfg3b.jsonl | cut -c1-1500; tail -2 gpurun_out/r48.err
