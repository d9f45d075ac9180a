"""Short run of the tensor-core map for ncu (int8 unweighted and bf16 weighted, 20k x 60 x 100)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from cassiopeia_b200 import engine
from cassiopeia_b200.synthetic import simulate_character_matrix
n, m, S = 20000, 60, 100
cm, priors, table = simulate_character_matrix(n, m, S, seed=1, mutation_rate=0.7, missing_fraction=0.2)
with np.errstate(invalid="ignore"): w = np.nan_to_num(-np.log(table))
for ww in (None, w):
    for _ in range(2):
        D = engine.whd_map(cm, -1, ww, S, dtype="f32", method="tensor")
    torch.cuda.synchronize()
print("ok")
