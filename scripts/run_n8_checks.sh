TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519"
timeout 900 $TR scripts/north_star_200k.py --mode exact > gpurun_out/r2_north_star_200k_8gpu_exact.json 2> gpurun_out/r2_north_star_200k_8gpu_exact.err; tail -c 1800 gpurun_out/r2_north_star_200k_8gpu_exact.json; tail -3 gpurun_out/r2_north_star_200k_8gpu_exact.err
timeout 600 $TR bench.py --gpus 8 --steps 1 --warmup 1 > gpurun_out/r2_bench_n8_gather.json 2> gpurun_out/r2_bench_n8_gather.err; tail -c 600 gpurun_out/r2_bench_n8_gather.json; tail -2 gpurun_out/r2_bench_n8_gather.err
