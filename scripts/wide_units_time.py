"""Split work list / wide units of the pruned scan (CAS_PRUNE_WIDE=1, default) against short units for every
listed row (CAS_PRUNE_WIDE=0): identical joins, loop time.  NJ at the bench size, UPGMA on a tie-heavy input."""
import os, sys
import numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
os.environ["PRUNE_CHECK_DEFAULT_MIN_K"] = "1"
from prune_check import build_map, run

for n, algo, rate in ((50000, "nj", 0.1), (20000, "upgma", 0.1), (20000, "upgma", 0.7)):
    D, nn = build_map(n, 60, 100, "f32", rate, 0.2)
    res = {}
    for wide in ("1", "0", "1", "0"):
        os.environ["CAS_PRUNE_WIDE"] = wide
        j, l, st, ms = run(D, nn, algo, "fast", True)
        res.setdefault(wide, []).append((ms, j, l, st))
    same = all(np.array_equal(res["1"][0][1], r[1]) and np.array_equal(res["1"][0][2], r[2]) for w in res for r in res[w])
    print(f"n={nn} {algo} rate={rate}: wide units {min(r[0] for r in res['1']) / 1e3:.3f} s, short units "
          f"{min(r[0] for r in res['0']) / 1e3:.3f} s, identical joins {same}, rows listed / iteration "
          f"{res['1'][0][3]['rows_listed'] / nn:.0f}", flush=True)
    del D
