# BASELINE config 4 (NJ 200k x 40, exact mode) as a bench.py line on N GPUs of one node: bash scripts/run_c4_bench.sh N
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --config c4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_c4_n$N.json 2> gpurun_out/r2_bench_c4_n$N.err
grep -v "OMP_NUM\|^\*\*\*\|^W1" gpurun_out/r2_bench_c4_n$N.err | tail -5
tail -c 2500 gpurun_out/r2_bench_c4_n$N.json
