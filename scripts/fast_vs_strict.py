"""Fast (float32 map, maintained row sums) vs strict (float64, reference arithmetic) NJ / UPGMA on simulated
data: unrooted Robinson-Foulds distance between the two trees next to the near-tie counters of the strict
run.  The north star promises RF = 0 for the fast mode only when no near-ties occur."""
import sys
import numpy as np
sys.path.insert(0, ".")
from cassiopeia_b200 import engine
from cassiopeia_b200.synthetic import simulate_character_matrix
from oracle import treecmp

print("| algo | cells | mutation rate | priors | joins | near-tie iterations (strict, 1e-6) | exact-tie iterations | joins that differ | RF(fast, strict) | RF / splits |")
print("|---|---:|---:|:---:|---:|---:|---:|---:|---:|---:|")
for algo in ("nj", "upgma"):
    for n in (1024, 4096, 8192):
        for rate in (0.1, 0.7):
            for weighted in (True, False):
                cm, priors, table = simulate_character_matrix(n, 40, 50, seed=n, mutation_rate=rate, missing_fraction=0.1)
                if algo == "nj":
                    cm = np.vstack([cm, np.zeros((1, 40), dtype=cm.dtype)])
                with np.errstate(invalid="ignore"):
                    w = np.nan_to_num(-np.log(table)) if weighted else None
                N = cm.shape[0]
                names = [f"c{i}" for i in range(N)]
                res = {}
                for mode, dt in (("strict", "f64"), ("fast", "f32")):
                    D = engine.whd_map(cm, -1, w, 50, dtype=dt, method="exact")
                    joins, last2 = engine.agglomerate(D, N, algo, mode)
                    res[mode] = (joins, last2, dict(engine.last_stats))
                trees = {}
                for mode in res:
                    j, l, _ = res[mode]
                    trees[mode] = treecmp.tree_from_joins(j, N, names, l, root_name=(names[-1] if algo == "nj" else "root"),
                                                          mode=algo)
                # NJ trees are rooted at the implicit root sample (last row), which is not a leaf
                li = {nm: i for i, nm in enumerate(names if algo == "upgma" else names[:-1])}
                rf = treecmp.rf_distance(trees["strict"], trees["fast"], li)
                nsplit = 2 * (len(li) - 3)
                diff = int((res["strict"][0] != res["fast"][0]).any(axis=1).sum())
                st = res["strict"][2]
                print(f"| {algo} | {n} | {rate} | {'yes' if weighted else 'no'} | {N - 2} | {st['near_ties']} | {st['exact_ties']} | "
                      f"{diff} | {rf} | {rf / nsplit:.4f} |", flush=True)
