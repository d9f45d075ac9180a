"""NJ solve seconds @ n cells on one B200 (BASELINE.json metric (i)): map + complete loop, fast mode (float32 map),
default scan policy (pruned above 2048 live nodes), with the dense scan beside it where that takes seconds.
60 characters, 100 states, priors, 20 % missing, implicit root.  Prints a markdown table."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from cassiopeia_b200 import engine
from cassiopeia_b200.synthetic import simulate_character_matrix


def solve(cm, w, S, n, dense):
    if dense:
        os.environ["CAS_PRUNE"] = "0"
    else:
        os.environ.pop("CAS_PRUNE", None)
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize()
    e[0].record()
    D = engine.whd_map(cm, -1, w, S, dtype="f32", method="exact")
    e[1].record()
    joins, last2 = engine.agglomerate(D, n, "nj", "fast")
    e[2].record()
    torch.cuda.synchronize()
    os.environ.pop("CAS_PRUNE", None)
    return joins, e[0].elapsed_time(e[1]) / 1e3, e[1].elapsed_time(e[2]) / 1e3, dict(engine.last_stats)


print("| cells | map s | loop s (default) | solve s (default) | us / iteration | map elements read | loop s (dense scan) | same joins |")
print("|---:|---:|---:|---:|---:|---:|---:|:---:|")
for cells, rate in ((1000, 0.1), (5000, 0.1), (10000, 0.1), (20000, 0.1), (50000, 0.1), (50000, 0.7), (100000, 0.1)):
    cm, priors, table = simulate_character_matrix(cells, 60, 100, seed=0, mutation_rate=rate, missing_fraction=0.2)
    cm = np.vstack([cm, np.zeros((1, 60), dtype=cm.dtype)])
    n = cells + 1
    with np.errstate(invalid="ignore", divide="ignore"):
        w = np.nan_to_num(-np.log(table), nan=0.0, posinf=0.0)
    cm_dev = torch.from_numpy(cm).cuda()
    solve(cm_dev, w, 100, n, False) if cells <= 20000 else None  # warm-up
    j, tm, tl, st = solve(cm_dev, w, 100, n, False)
    dense_total = sum(k * (k - 1) // 2 for k in range(3, n + 1))
    if cells <= 20000:
        jd, _, tld, _ = solve(cm_dev, w, 100, n, True)
        dense_s, same = f"{tld:.2f}", str(bool(np.array_equal(j, jd)))
    else:
        dense_s, same = "-", "-"
    print(f"| {cells} (rate {rate}) | {tm:.3f} | {tl:.2f} | {tm + tl:.2f} | {tl / (n - 2) * 1e6:.1f} | "
          f"{st['scanned_elements'] / dense_total:.4%} | {dense_s} | {same} |", flush=True)
