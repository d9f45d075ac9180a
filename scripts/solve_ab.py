"""A/B of NeighborJoiningSolver.solve() wall time at 50k cells with and without the tail worker thread."""
import os, sys, time
import numpy as np, pandas as pd, torch
sys.path.insert(0, ".")
import argparse, bench
import cassiopeia_b200 as cb
from cassiopeia_b200.synthetic import simulate_character_matrix
n, m, S = 50000, 60, 100
cm, priors, _ = simulate_character_matrix(n, m, S, seed=0, mutation_rate=0.1, missing_fraction=0.2)
names = [f"c{i}" for i in range(n)]
for rep in range(3):
    for flag in ("1", "0"):
        os.environ["CAS_TAIL_THREAD"] = flag
        tree = cb.CassiopeiaTree(character_matrix=pd.DataFrame(cm, index=names), priors=priors)
        solver = cb.NeighborJoiningSolver(add_root=True)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        solver.solve(tree, collapse_mutationless_edges=True)
        dt = time.perf_counter() - t0
        print(f"rep {rep} worker thread {'on ' if flag == '1' else 'off'}: {dt:.3f} s", flush=True)
