# ncu evidence for the tie-aware UPGMA loop + C2 bench line after the e2e warm-up call
mkdir -p gpurun_out
timeout 600 python bench.py --config c2 --no-cpu-baseline > gpurun_out/r2_bench_c2_upgma10k.json 2> gpurun_out/c2.err; tail -2 gpurun_out/c2.err; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_c2_upgma10k.json')); print('c2', d['value'], d['e2e']['value'], d['e2e']['device_ms'], d['e2e']['solver_api']['value'])"
cat > /tmp/upgma_prof.py <<'PY'
import sys, os
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
os.environ["PRUNE_CHECK_DEFAULT_MIN_K"] = "1"
from prune_check import build_map
from cassiopeia_b200 import engine
import torch
D, nn = build_map(50000, 40, 50, "f64", 0.1, 0.1, weighted=False)
engine.agglomerate(D, nn, "upgma", "strict", max_iters=5040)
torch.cuda.synchronize()
PY
timeout 300 python /tmp/upgma_prof.py && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 15020 --launch-count 60 --csv --log-file gpurun_out/r2_launches_upgma_keyed_k45k.csv python /tmp/upgma_prof.py > gpurun_out/ncu_upgma.log 2>&1; tail -3 gpurun_out/ncu_upgma.log; head -5 gpurun_out/r2_launches_upgma_keyed_k45k.csv
