set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
timeout 500 $TR bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r25_bench_cfg3_f64_n8.json 2> gpurun_out/r25_bench_cfg3_f64_n8.err
timeout 400 $TR bench.py --gpus 8 --steps 3 --warmup 3 --dtype f32 > gpurun_out/r25_bench_cfg3_f32_n8.json 2> gpurun_out/r25_bench_cfg3_f32_n8.err
timeout 300 $TR tools/bench_etkf.py > gpurun_out/r25_etkf_f64_n8.json 2> gpurun_out/r25_etkf_f64_n8.err
head -c 400 gpurun_out/r25_bench_cfg3_f64_n8.json; echo; head -c 400 gpurun_out/r25_bench_cfg3_f32_n8.json; echo; cat gpurun_out/r25_etkf_f64_n8.json; tail -3 gpurun_out/r25_bench_cfg3_f64_n8.err
