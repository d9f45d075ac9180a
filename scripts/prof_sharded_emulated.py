"""ncu target: the sharded exact-mode loop with 2 virtual ranks on ONE GPU (cas_sharded_solve_emulated_p2p), first
`--iters` iterations at 50k cells -- per-kernel durations of the sharded loop's six launches per iteration."""
import argparse, sys
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from cassiopeia_b200 import sharded

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=50000)
ap.add_argument("--iters", type=int, default=3040)
ap.add_argument("--world", type=int, default=2)
a = ap.parse_args()
args = argparse.Namespace(cells=a.cells, chars=60, states=100, missing=0.2, mutation_rate=0.1, seed=0, algo="nj", priors=True)
codes, missing_code, table, n_states = bench.make_workload(args)
sharded.emulated_solve(codes, missing_code, table, n_states, "nj", a.world, dtype="f64", max_iters=a.iters, p2p=True,
                       mode="exact")
torch.cuda.synchronize()
print("done")
