#!/usr/bin/env python
"""Times the REFERENCE's own numba / pandas CPU path in this container:

  1. `cassiopeia.data.utilities.compute_dissimilarity_map(cm, n, weighted_hamming_distance, weights, -1, threads=T)`
     (cassiopeia/data/utilities.py:189-295) at n in --map-cells, T = 1 and T = all cores -> Mpairs/s;
  2. full `NeighborJoiningSolver(add_root=True).solve(tree)` / `UPGMASolver().solve(tree)`
     (cassiopeia/solver/DistanceSolver.py:127-231) at n in --solve-cells.

The reference is imported from /root/reference through oracle/ref_loader.py (stubs for the absent third-party
packages); inputs come from the product's vectorised generator (same process as the reference's
simulators).  Writes a markdown table (stdout)."""
import argparse
import os
import sys
import time

import numpy as np
import pandas as pd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--map-cells", type=int, nargs="*", default=[2000, 6000, 10000])
    ap.add_argument("--solve-cells", type=int, nargs="*", default=[512, 1024, 2048, 4096])
    ap.add_argument("--chars", type=int, default=40)
    ap.add_argument("--states", type=int, default=50)
    a = ap.parse_args()
    from cassiopeia_b200.synthetic import simulate_character_matrix
    from oracle import ref_loader

    ref = ref_loader.load_reference()
    cores = os.cpu_count()
    print(f"# Reference CPU path (numba {__import__('numba').__version__}, numpy {np.__version__}, pandas "
          f"{pd.__version__}) on this container's {cores} host cores\n")
    print(f"Inputs: simulated {a.chars} characters x {a.states} states, mutation rate 0.1, 10 % missing, priors "
          "(negative_log).\n")
    whd = ref.dissimilarity_functions.weighted_hamming_distance
    print("## 1. `data.utilities.compute_dissimilarity_map` (weighted_hamming_distance with priors)\n")
    print("| cells | pairs | threads | seconds (first call, incl. numba compile) | seconds (second call) | Mpairs/s (second) |")
    print("|---:|---:|---:|---:|---:|---:|")
    for n in a.map_cells:
        cm, priors, _ = simulate_character_matrix(n, a.chars, a.states, seed=0, mutation_rate=0.1, missing_fraction=0.1)
        weights = ref.solver_utilities.transform_priors(priors, "negative_log")
        arr = cm.astype(np.int64)
        for T in (1, cores):
            t0 = time.perf_counter()
            ref.data_utilities.compute_dissimilarity_map(arr, n, whd, weights, -1, threads=T)
            t1 = time.perf_counter()
            ref.data_utilities.compute_dissimilarity_map(arr, n, whd, weights, -1, threads=T)
            t2 = time.perf_counter()
            pairs = n * (n - 1) // 2
            print(f"| {n} | {pairs} | {T} | {t1 - t0:.2f} | {t2 - t1:.2f} | {pairs / (t2 - t1) / 1e6:.2f} |", flush=True)
    print("\n## 2. full `solve(tree)` (map + single-threaded numba / pandas loop + tree tail)\n")
    print("| solver | cells | seconds |")
    print("|---|---:|---:|")
    for n in a.solve_cells:
        cm, priors, _ = simulate_character_matrix(n, a.chars, a.states, seed=0, mutation_rate=0.1, missing_fraction=0.1)
        frame = pd.DataFrame(cm, index=[f"c{i}" for i in range(n)])
        for name, make in (("NeighborJoiningSolver(add_root=True)", lambda: ref.NeighborJoiningSolver(add_root=True)),
                           ("UPGMASolver()", lambda: ref.UPGMASolver())):
            tree = ref.CassiopeiaTree(character_matrix=frame.copy(), priors=priors)
            t0 = time.perf_counter()
            make().solve(tree)
            t1 = time.perf_counter()
            print(f"| {name} | {n} | {t1 - t0:.1f} |", flush=True)
    print("\nThe loop is O(n^3) and single-threaded (pandas frame rebuilt every iteration, DistanceSolver.py:185-205): "
          "50 000 cells extrapolate to (50000 / 4096)^3 x the 4096-cell time, i.e. days.")


if __name__ == "__main__":
    main()
