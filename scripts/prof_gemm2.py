"""Short run of the SM-pair tensor-core map for ncu: int8 20k x 60 x 100, int8 30k x 200 x 100, bf16 20k x 60 x 100."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from cassiopeia_b200 import engine
from cassiopeia_b200.synthetic import simulate_character_matrix
for n, m, S, weighted in ((20000, 60, 100, False), (30000, 200, 100, False), (20000, 60, 100, True)):
    cm, priors, table = simulate_character_matrix(n, m, S, seed=1, mutation_rate=0.7, missing_fraction=0.2)
    with np.errstate(invalid="ignore"): w = np.nan_to_num(-np.log(table))
    for _ in range(2):
        D = engine.whd_map(cm, -1, w if weighted else None, S, dtype="f32", method="tensor")
    torch.cuda.synchronize()
    del D
print("ok")
