#!/usr/bin/env python
"""Times the NJ loop modes on one GPU: fast (float32), exact (adaptive strict, float64) and, up to --strict-max
cells, strict (float64, every row sum re-accumulated each iteration); checks exact == strict and reports
RF(fast, exact), near-tie / ambiguity statistics.  Output: markdown table on stdout."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rf_from_joins(ja, la, jb, lb, n):
    """Unrooted Robinson-Foulds distance of two join lists over the same n leaves (bipartition hash sets)."""
    rng = np.random.default_rng(12345)
    key = rng.integers(1, 2 ** 62, size=n, dtype=np.int64)
    total = int(np.bitwise_xor.reduce(key))

    def splits(j):
        h = np.zeros(2 * n, dtype=np.int64)
        h[:n] = key
        out = set()
        for t, (a, b) in enumerate(j.tolist()):
            h[n + t] = h[a] ^ h[b]
            v = int(h[n + t])
            out.add(min(v, v ^ total))
        return out

    sa, sb = splits(ja), splits(jb)
    return len(sa ^ sb), len(sa)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, nargs="+", default=[20000, 50000])
    ap.add_argument("--chars", type=int, default=60)
    ap.add_argument("--states", type=int, default=100)
    ap.add_argument("--missing", type=float, default=0.2)
    ap.add_argument("--rate", type=float, default=0.1)
    ap.add_argument("--strict-max", type=int, default=20000)
    ap.add_argument("--no-priors", action="store_true")
    args = ap.parse_args()
    import torch
    from cassiopeia_b200 import engine
    from cassiopeia_b200.synthetic import simulate_character_matrix

    print("| cells | mode | map s | loop s | us/iter | near ties | exact ties | ambiguous | rows resolved | max rows | "
          "pairs changed | read frac | == strict | RF(fast, exact) |")
    print("|---:|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|---|")
    for n in args.cells:
        cm, priors, table = simulate_character_matrix(n, args.chars, args.states, seed=0, mutation_rate=args.rate,
                                                      missing_fraction=args.missing)
        cm = np.vstack([cm, np.zeros((1, args.chars), dtype=cm.dtype)])
        with np.errstate(invalid="ignore", divide="ignore"):
            w = None if args.no_priors else np.nan_to_num(-np.log(table), nan=0.0, posinf=0.0)
        nn = n + 1
        res = {}
        modes = ["fast", "exact"] + (["strict"] if n <= args.strict_max else [])
        for mode in modes:
            dt = "f32" if mode == "fast" else "f64"
            for rep in range(2):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                D = engine.whd_map(cm, -1, w, args.states, dtype=dt, method="exact")
                torch.cuda.synchronize()
                t1 = time.perf_counter()
                j, l = engine.agglomerate(D, nn, "nj", mode)
                torch.cuda.synchronize()
                t2 = time.perf_counter()
                del D
                if mode == "strict":
                    break
            res[mode] = (j, l, t1 - t0, t2 - t1, dict(engine.last_stats))
        dense = sum(k * (k - 1) // 2 for k in range(3, nn + 1))
        for mode in modes:
            j, l, tm, tl, st = res[mode]
            same = "" if "strict" not in res else str(bool(np.array_equal(j, res["strict"][0])))
            rf = ""
            if mode == "fast":
                d, tot = rf_from_joins(j, l, res["exact"][0], res["exact"][1], nn)
                rf = f"{d} / {tot}"
            print(f"| {n} | {mode} | {tm:.3f} | {tl:.3f} | {1e6 * tl / (nn - 2):.1f} | {st['near_ties']} | "
                  f"{st['exact_ties']} | {st['ambiguous_iterations']} | {st['rows_resolved']} | "
                  f"{st['max_rows_resolved']} | {st['pairs_changed']} | {st['scanned_elements'] / dense:.2e} | "
                  f"{same} | {rf} |", flush=True)


if __name__ == "__main__":
    main()
