"""UPGMA loop on one GPU: dense scan vs value-only pruned scan vs tie-aware ("keyed") pruned scan (agglom_keyed.cuh).
Identical join lists, loop seconds, elements read.  Usage: upgma_keyed_time.py [--cells N ...] [--no-dense-above N]"""
import argparse, os, sys
import numpy as np, torch
sys.path.insert(0, ".")
sys.path.insert(0, "scripts")
os.environ["PRUNE_CHECK_DEFAULT_MIN_K"] = "1"  # keep the default switch to the dense scan below 2048 live nodes
from prune_check import build_map, run

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, nargs="+", default=[10000, 20000, 50000])
ap.add_argument("--no-dense-above", type=int, default=50000)
ap.add_argument("--no-value-only-above", type=int, default=20000)
a = ap.parse_args()
# (characters, states, dtype, weighted, mutation rate, missing): BASELINE config 2 is the plain / float64 / 40 x 50 row
shapes = [(40, 50, "f64", False, 0.1, 0.1, "plain Hamming f64 (C2 shape)"),
          (60, 100, "f32", True, 0.1, 0.2, "priors f32, rate 0.1"),
          (60, 100, "f64", True, 0.1, 0.2, "priors f64, rate 0.1"),
          (60, 100, "f32", True, 0.7, 0.2, "priors f32, rate 0.7")]
print("| cells | input | dense s | value-only pruned s | keyed pruned s | us / iteration (keyed) | read / dense (value-only) | read / dense (keyed) | rows listed / iteration (keyed) | exact-tie iterations (dense) | identical joins |")
print("|---:|---|---:|---:|---:|---:|---:|---:|---:|---:|---|")
for n in a.cells:
    for m, S, dtype, weighted, rate, miss, label in shapes:
        if n * n * (8 if dtype == "f64" else 4) > 60e9:
            continue
        D, nn = build_map(n, m, S, dtype, rate, miss, weighted)
        mode = "strict" if dtype == "f64" else "fast"
        dense_el = sum(k * (k - 1) // 2 for k in range(3, nn + 1))
        os.environ["CAS_UPGMA_KEYED"] = "1"
        jk, lk, sk, tk = run(D, nn, "upgma", mode, True)
        same, td, tv, ties, readv = "", float("nan"), float("nan"), "", float("nan")
        if n <= a.no_dense_above:
            jd, ld, sd, td = run(D, nn, "upgma", mode, False)
            same = str(bool(np.array_equal(jd, jk) and np.array_equal(ld, lk)))
            ties = str(sd["exact_ties"])
        if n <= a.no_value_only_above:
            os.environ["CAS_UPGMA_KEYED"] = "0"
            jv, lv, sv, tv = run(D, nn, "upgma", mode, True)
            os.environ["CAS_UPGMA_KEYED"] = "1"
            readv = sv["scanned_elements"] / dense_el
            same += "/" + str(bool(np.array_equal(jv, jk)))
        print(f"| {nn} | {label} | {td / 1e3:.3f} | {tv / 1e3:.3f} | {tk / 1e3:.3f} | {1e3 * tk / max(nn - 2, 1):.1f} | "
              f"{readv:.5f} | {sk['scanned_elements'] / dense_el:.5f} | {sk['rows_listed'] / max(nn - 2, 1):.0f} | {ties} | {same} |",
              flush=True)
        del D
