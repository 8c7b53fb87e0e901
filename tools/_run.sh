T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus 8 --workload config4 --steps 30 --warmup 3 --no-cpu > gpurun_out/r01_g4_bench_config4_8gpu.json 2> gpurun_out/r01_g4_bench_config4_8gpu.err
tail -c 2500 gpurun_out/r01_g4_bench_config4_8gpu.json
$T bench.py --gpus 8 --steps 100 --warmup 5 --no-cpu > gpurun_out/r01_g4_bench_config2_8gpu.json 2> gpurun_out/r01_g4_bench_config2_8gpu.err
tail -c 2500 gpurun_out/r01_g4_bench_config2_8gpu.json
