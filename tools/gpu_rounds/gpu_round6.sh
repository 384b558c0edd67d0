set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --e2e-steps 1 > gpurun_out/r6_bench_cfg3_f64_n2.json 2> gpurun_out/r6_bench_cfg3_f64_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --e2e-steps 2 --dtype f32 > gpurun_out/r6_bench_cfg3_f32_n2.json 2> gpurun_out/r6_bench_cfg3_f32_n2.err
python bench.py --dtype f32 > gpurun_out/r6_bench_cfg3_f32_n1.json 2> gpurun_out/r6_bench_cfg3_f32_n1.err
