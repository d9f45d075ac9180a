"""UPGMA: pruned scan (CAS_PRUNE=1) against the dense scan -- identical joins, loop time, elements read.
Usage: upgma_prune_time.py [cells_with_dense ...] [--pruned-only cells ...]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
os.environ["PRUNE_CHECK_DEFAULT_MIN_K"] = "1"
from prune_check import build_map, run

both = [int(x) for x in sys.argv[1:] if not x.startswith("-") and sys.argv.index(x) < (sys.argv.index("--pruned-only") if "--pruned-only" in sys.argv else 10**9)]
only = [int(x) for x in sys.argv[sys.argv.index("--pruned-only") + 1:]] if "--pruned-only" in sys.argv else []
for n in both + only:
    for rate in (0.1, 0.7):
        D, nn = build_map(n, 60, 100, "f32", rate, 0.2)
        dense = sum(k * (k - 1) // 2 for k in range(3, nn + 1))
        jp, lp, sp, tp = run(D, nn, "upgma", "fast", True)
        line = f"n={nn} rate={rate} upgma: pruned {tp / 1e3:.2f} s (read {sp['scanned_elements'] / dense:.5f} of dense, rows listed/iteration {sp['rows_listed'] / max(nn - 2, 1):.0f})"
        if n in both:
            jd, ld, sd, td = run(D, nn, "upgma", "fast", False)
            same = np.array_equal(jd, jp) and np.array_equal(ld, lp) and sd["near_ties"] == sp["near_ties"] and sd["exact_ties"] == sp["exact_ties"]
            line += f", dense {td / 1e3:.2f} s, identical {same}, near ties {sd['near_ties']} exact ties {sd['exact_ties']}"
        print(line, flush=True)
        del D
