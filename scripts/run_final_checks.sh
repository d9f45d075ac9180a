# Round-2 final validation on one B200: GPU test suite, smoke, the default bench line (C3, exact mode), fast mode, C1, C2.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches','mode')}); print(d['e2e']['value'], d['e2e']['solver_api']['value']); print(d['roofline']['latency_model']['us_per_iteration'], d['roofline']['map_ms'], d['cpu_baseline']['value'], d['parity']['join_hash'])"; tail -2 gpurun_out/r2_bench_n1.err
python bench.py --mode fast --no-cpu-baseline > gpurun_out/r2_bench_n1_fast.json 2> /dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1_fast.json')); print('fast', {k:d[k] for k in ('value','mode')}, d['e2e']['value'], d['parity']['join_hash'])"
python bench.py --config c1 --no-cpu-baseline > gpurun_out/r2_bench_c1_nj1024.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c1_nj1024.json')); print('c1', d['value'], d['e2e']['value'])"
python bench.py --config c2 --no-cpu-baseline > gpurun_out/r2_bench_c2_upgma10k.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c2_upgma10k.json')); print('c2', d['value'], d['e2e']['value'])"
python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-400
