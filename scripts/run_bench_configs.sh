python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches','mode')}); print(d['e2e']['value'], d['e2e']['solver_api']); print(d['roofline']['latency_model']); print(d['cpu_baseline']['value'])"; tail -2 gpurun_out/r2_bench_n1.err
python bench.py --config c2 > gpurun_out/r2_bench_c2.json 2> gpurun_out/r2_bench_c2.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c2.json')); print({k:d[k] for k in ('metric','value','ms_per_step','gpu_launches','mode')}); r=d['roofline']; print({k:r[k] for k in ('kernel','achieved','frac','map_ms','loop_ms')}); print(d['e2e']['value'], d['parity']); print(d['cpu_baseline']['value'])"; tail -2 gpurun_out/r2_bench_c2.err
python bench.py --config c1 > gpurun_out/r2_bench_c1.json 2> gpurun_out/r2_bench_c1.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c1.json')); print({k:d[k] for k in ('metric','value','ms_per_step','gpu_launches','mode')}); print(d['e2e']['value'], d['e2e']['solver_api'], d['cpu_baseline']['value'])"; tail -2 gpurun_out/r2_bench_c1.err
