# 2 GPUs: unweighted NJ 100k x 40, exact mode, shards built by the tcgen05 kernel in shard mode vs the CUDA-core kernel
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
mkdir -p gpurun_out
for MAP in tensor exact; do
  timeout 600 $TR scripts/north_star_200k.py --cells 100000 --mode exact --no-priors --map $MAP $([ $MAP = tensor ] && echo --compare-single) \
    > gpurun_out/r2_n2_plain_100k_$MAP.json 2> gpurun_out/r2_n2_plain_100k_$MAP.err
  grep -v "OMP_NUM\|^\*\*\*\|^W1" gpurun_out/r2_n2_plain_100k_$MAP.err | tail -3
  grep "^{" gpurun_out/r2_n2_plain_100k_$MAP.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$MAP', {k:d[k] for k in ('map_seconds','map_gpairs_per_s','loop_seconds','join_hash','sampled_map_rows_bit_exact_vs_cpu_oracle','map_kernel')}, d.get('single_gpu'))"
done
