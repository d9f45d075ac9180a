TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521"
timeout 800 $TR scripts/north_star_200k.py --cells 100000 --mode exact --compare-single > gpurun_out/r2_n2_exact_100k.json 2> gpurun_out/r2_n2_exact_100k.err; tail -c 1700 gpurun_out/r2_n2_exact_100k.json; grep -v "OMP_NUM\|^\*\*\*" gpurun_out/r2_n2_exact_100k.err | tail -5
