"""Pruned-scan diagnostics over windows of a solve: rows listed per iteration, drift, elements read."""
import argparse, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from cassiopeia_b200 import engine
ap = argparse.ArgumentParser(); ap.add_argument("--cells", type=int, default=100000)
ap.add_argument("--fracs", type=str, default="0.2,0.4,0.6,0.7,0.8,0.9"); a = ap.parse_args()
args = argparse.Namespace(cells=a.cells, chars=60, states=100, missing=0.2, mutation_rate=0.1, seed=0)
codes, missing_code, table, n_states = bench.make_workload(args)
n = codes.shape[0]
prev = {"rows_listed": 0, "scanned_elements": 0, "it": 0, "ms": 0.0}
for frac in [float(x) for x in a.fracs.split(",")]:
    iters = int(frac * n)
    D = engine.whd_map(codes, missing_code, table, n_states, dtype="f32", method="exact")
    torch.cuda.synchronize(); t0 = time.perf_counter()
    engine.agglomerate(D, n, "nj", "fast", max_iters=iters)
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) * 1e3
    st = dict(engine.last_stats)
    di = iters - prev["it"]
    print(f"iterations {prev['it']}..{iters} (k {n - prev['it']}..{n - iters}): rows listed / iteration "
          f"{(st['rows_listed'] - prev['rows_listed']) / di:.0f}, elements / iteration "
          f"{(st['scanned_elements'] - prev['scanned_elements']) / di:.0f}, us / iteration {(ms - prev['ms']) * 1e3 / di:.1f}, "
          f"drift C {st['drift']:.4f}, near ties {st['near_ties']}", flush=True)
    prev = {"rows_listed": st["rows_listed"], "scanned_elements": st["scanned_elements"], "it": iters, "ms": ms}
    del D
