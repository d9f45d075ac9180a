"""Wall time of the public API: NeighborJoiningSolver.solve(CassiopeiaTree) on synthetic data."""
import argparse, sys, time
import numpy as np, pandas as pd
sys.path.insert(0, ".")
import cassiopeia_b200 as cb
from cassiopeia_b200.synthetic import simulate_character_matrix

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=20000)
ap.add_argument("--algo", default="nj")
ap.add_argument("--fast", action="store_true")
a = ap.parse_args()
cm, priors, _ = simulate_character_matrix(a.cells, 60, 100, seed=0, mutation_rate=0.1, missing_fraction=0.2)
tree = cb.CassiopeiaTree(character_matrix=pd.DataFrame(cm, index=[f"c{i}" for i in range(a.cells)]), priors=priors)
solver = (cb.NeighborJoiningSolver(add_root=True, fast=a.fast) if a.algo == "nj" else cb.UPGMASolver(fast=a.fast))
import torch; torch.zeros(1).cuda()
t0 = time.perf_counter()
solver.solve(tree, collapse_mutationless_edges=True)
t1 = time.perf_counter()
print(f"{a.algo} solve({a.cells} cells, fast={a.fast}) wall {t1 - t0:.2f} s; nodes {len(tree.nodes)}, leaves {len(tree.leaves)}")
t0 = time.perf_counter(); s = tree.get_newick(); print(f"newick {len(s)} chars in {time.perf_counter() - t0:.2f} s")
