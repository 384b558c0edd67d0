set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r16_bench_cfg3_f64_n4.json 2> gpurun_out/r16_bench_cfg3_f64_n4.err
timeout 600 $TR bench.py --gpus 4 --steps 3 --warmup 3 --dtype f32 > gpurun_out/r16_bench_cfg3_f32_n4.json 2> gpurun_out/r16_bench_cfg3_f32_n4.err
timeout 300 $TR tools/bench_etkf.py > gpurun_out/r16_etkf_f64_n4.json 2> gpurun_out/r16_etkf_f64_n4.err
timeout 300 $TR tools/bench_etkf.py --dtype f32 > gpurun_out/r16_etkf_f32_n4.json 2> gpurun_out/r16_etkf_f32_n4.err
cat gpurun_out/r16_etkf_f64_n4.json; tail -3 gpurun_out/r16_etkf_f64_n4.err
