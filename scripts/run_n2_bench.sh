# bench.py under torchrun on 2 GPUs: default policy (map rows on 2 GPUs, loop on one), the forced sharded loop, the
# fast-mode sharded loop (tensor-core shards for the prior-weighted map), and the reference arm (rank 0 only)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517"
mkdir -p gpurun_out
show() { grep "^{" $1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d.get('roofline',{}); print('$2', d.get('value'), d.get('e2e',{}).get('value'), d.get('loop'), r.get('map_kernel'), r.get('map_and_gather_ms'), d.get('parity',{}).get('join_hash'), d.get('impl'))"; }
$TR bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2_bench_n2.json 2> gpurun_out/n2a.err; show gpurun_out/r2_bench_n2.json gather
$TR bench.py --gpus 2 --steps 1 --warmup 1 --sharded-loop --no-cpu-baseline > gpurun_out/r2_bench_n2_sharded_loop.json 2> gpurun_out/n2b.err; show gpurun_out/r2_bench_n2_sharded_loop.json sharded
$TR bench.py --gpus 2 --steps 1 --warmup 1 --sharded-loop --mode fast --states 50 --no-cpu-baseline > gpurun_out/r2_bench_n2_sharded_fast.json 2> gpurun_out/n2c.err; show gpurun_out/r2_bench_n2_sharded_fast.json sharded-fast
tail -2 gpurun_out/n2c.err
