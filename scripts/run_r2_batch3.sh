# after making the tie-aware pruned scan the UPGMA default: full GPU suite, UPGMA timings at 50k, C2 bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 900 python scripts/upgma_keyed_time.py --cells 50000 > gpurun_out/r2_upgma_keyed_50k.md 2>&1; cat gpurun_out/r2_upgma_keyed_50k.md
timeout 600 python bench.py --config c2 --no-cpu-baseline > gpurun_out/r2_bench_c2_upgma10k.json 2> gpurun_out/c2.err; tail -2 gpurun_out/c2.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c2_upgma10k.json')); print('c2', d['value'], d['e2e']['value'], d['roofline']['kernel'], d['roofline'].get('latency_model',{}).get('us_per_iteration'), d['parity']['join_hash'])"
