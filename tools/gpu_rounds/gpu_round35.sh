set -x
mkdir -p gpurun_out
( timeout 300 python tools/bench_ketkf.py --steps 5 --warmup 3 > gpurun_out/r35_bench_widened.jsonl ) 2> gpurun_out/r35_bench_widened.err
cat gpurun_out/r35_bench_widened.jsonl; tail -5 gpurun_out/r35_bench_widened.err
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:k_kernelise|k_ienks_pre|k_ienks_keep' -c 4 -f -o gpurun_out/r35_widened python tools/bench_ketkf.py --only cfg2 --steps 1 --warmup 0 > gpurun_out/r35_ncu.log 2>&1
tail -3 gpurun_out/r35_ncu.log; ls -la gpurun_out/
