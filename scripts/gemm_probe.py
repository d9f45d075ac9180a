import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from cassiopeia_b200 import engine
from cassiopeia_b200.synthetic import simulate_character_matrix
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = fn()
    e1.record(); torch.cuda.synchronize()
    return r, e0.elapsed_time(e1) / reps
for n, m, S in ((300, 8, 4), (4000, 40, 50), (20000, 40, 50), (20000, 60, 100)):
    cm, priors, table = simulate_character_matrix(n, m, S, seed=1, mutation_rate=0.7, missing_fraction=0.1)
    with np.errstate(invalid="ignore"): w = np.nan_to_num(-np.log(table))
    for weighted in (False, True):
        ww = w if weighted else None
        ref, te = timeit(lambda: engine.whd_map(cm, -1, ww, S, dtype="f32", method="exact"))
        D, tt = timeit(lambda: engine.whd_map(cm, -1, ww, S, dtype="f32", method="tensor"))
        a, b = D[:, :n].double(), ref[:, :n].double()
        err = ((a - b).abs() / b.abs().clamp(min=1.0)).max().item()
        pairs = n * (n - 1) / 2
        flop = pairs * 2 * m * (S + 3)
        print(f"n={n} m={m} S={S} weighted={weighted}: exact {te:.2f} ms ({pairs/te/1e6:.1f} Gpairs/s) tensor {tt:.2f} ms "
              f"({pairs/tt/1e6:.1f} Gpairs/s, {flop/tt/1e9:.1f} alg TFLOP/s) max rel err {err:.2e}", flush=True)
