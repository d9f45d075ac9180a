python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches','mode')}); print(d['e2e']['value'], d['e2e']['solver_api']['value']); print(d['roofline']['latency_model']['us_per_iteration'], d['roofline']['map_ms'], d['cpu_baseline']['value'])"; tail -2 gpurun_out/r2_bench_n1.err
python bench.py --mode fast --no-cpu-baseline > gpurun_out/r2_bench_n1_fast.json 2> gpurun_out/r2_bench_n1_fast.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1_fast.json')); print('fast', {k:d[k] for k in ('value','mode')}, d['parity']['join_hash'])"
python bench.py --config c2 --no-cpu-baseline > gpurun_out/r2_bench_c2.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c2.json')); print('c2', d['value'], d['roofline']['frac'])"
python scripts/exact_time.py --cells 20000 50000 100000 --strict-max 20000 > gpurun_out/r2_exact_time.md 2>&1; tail -9 gpurun_out/r2_exact_time.md
python scripts/exact_time.py --cells 100000 --chars 40 --states 50 --missing 0.1 --strict-max 0 > gpurun_out/r2_exact_time_c4like.md 2>&1; tail -3 gpurun_out/r2_exact_time_c4like.md
