"""Short run of the hot path for ncu: map + the first iterations of the NJ loop at workload size."""
import argparse, sys
import numpy as np
import torch
sys.path.insert(0, ".")
import bench
from cassiopeia_b200 import engine

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=50000)
ap.add_argument("--iters", type=int, default=24)
ap.add_argument("--mode", default="fast")
ap.add_argument("--algo", default="nj")
ap.add_argument("--method", default="exact")
a = ap.parse_args()
args = argparse.Namespace(cells=a.cells, chars=60, states=100, missing=0.2, mutation_rate=0.1, seed=0)
codes, missing_code, table, n_states = bench.make_workload(args)
n = codes.shape[0]
dt = "f32" if a.mode == "fast" else "f64"
D = engine.whd_map(codes, missing_code, table if a.algo == "nj" else None, n_states, dtype=dt, method=a.method)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
joins, last2 = engine.agglomerate(D, n, a.algo, a.mode, max_iters=a.iters)
e1.record(); torch.cuda.synchronize()
print(f"n={n} {a.algo} {a.mode}: {a.iters} iterations in {e0.elapsed_time(e1):.2f} ms; first joins {joins[:3].tolist()}")
