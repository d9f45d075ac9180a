"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the REFERENCE.

Run in the build container (needs /root/reference):

    python -m oracle.make_golden            # all cases
    python -m oracle.make_golden sim64_s0   # one case

Every fixture holds the inputs (character matrix, prior table, missing
indicator) and the reference's own outputs for them:

* ``order``      -- row order of the reference map (CassiopeiaTree.py:1920-1951),
                    as indices into the character matrix rows (index n = the
                    implicit all-zero "root" row NJ appends, NeighborJoiningSolver.py:277-280)
* ``D``          -- the reference float64 map in that order (small cases), else
                    ``D_rows`` (a few full rows) + ``D_sha256``
* ``joins``      -- the reference join sequence in id space (oracle/ref_loader.py)
* ``clades_*``   -- hashed clade / size / depth tables of the reference's final
                    tree without and with collapse_mutationless_edges
"""
from __future__ import annotations

import hashlib
import os
import sys
import time
import warnings

import numpy as np
import pandas as pd

from oracle import ref_loader, treecmp

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def simulate(ref, n_cells, m, S, seed, mutation_rate):
    """benchmark.ipynb cell 8 parameters."""
    np.random.seed(seed)
    tree = ref.CompleteBinarySimulator(num_cells=n_cells).simulate_tree()
    sim = ref.Cas9LineageTracingDataSimulator(
        number_of_cassettes=m, size_of_cassette=1, mutation_rate=mutation_rate,
        number_of_states=S, state_priors=None, heritable_silencing_rate=9e-4,
        stochastic_silencing_rate=0.1, random_seed=seed,
    )
    sim.overlay_data(tree)
    cm = tree.character_matrix.copy()
    cm.index = [str(i) for i in cm.index]
    priors = {c: dict(sim.mutation_priors_per_character[c]) for c in range(m)}
    return cm, priors


def random_tie_heavy(n, m, S, seed, missing_frac, dup_frac):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, S + 1, size=(n, m))
    base[rng.random((n, m)) < missing_frac] = -1
    # force duplicates
    for r in range(n):
        if rng.random() < dup_frac and r > 0:
            base[r] = base[rng.integers(0, r)]
    cm = pd.DataFrame(base, index=[f"c{i}" for i in range(n)])
    pr = {}
    for c in range(m):
        p = rng.exponential(1.0, size=S)
        p = p / p.sum()
        pr[c] = {s + 1: float(p[s]) for s in range(S)}
    return cm, pr


def prior_table(priors, m):
    if not priors:
        return np.zeros((0, 0), dtype=np.float64)
    S = max(max(v.keys()) for v in priors.values())
    tab = np.full((m, S + 1), np.nan, dtype=np.float64)
    for c, sv in priors.items():
        for s, p in sv.items():
            tab[c, s] = p
    return tab


def run_case(name, cm, priors, algo, prior_transformation="negative_log", store_D=True, missing=-1):
    ref = ref_loader.load_reference()
    RecNJ, RecUPGMA = ref_loader.recording_solvers()
    n, m = cm.shape
    names = list(cm.index)
    leaf_index = {nm: i for i, nm in enumerate(names)}
    out = {
        "cm": cm.to_numpy().astype(np.int32),
        "priors": prior_table(priors, m),
        "missing": np.int64(missing),
        "algo": np.array(algo),
        "prior_transformation": np.array(prior_transformation),
    }
    t0 = time.time()
    for collapse in (False, True):
        tree = ref.CassiopeiaTree(character_matrix=cm, priors=priors if priors else None,
                                  missing_state_indicator=missing)
        if algo == "nj":
            solver = RecNJ(add_root=True, prior_transformation=prior_transformation)
        else:
            solver = RecUPGMA(prior_transformation=prior_transformation)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            solver.solve(tree, collapse_mutationless_edges=collapse)
        if not collapse:
            dmap = tree.get_dissimilarity_map()
            pos = {nm: i for i, nm in enumerate(names)}
            pos["root"] = n
            out["order"] = np.array([pos[x] for x in dmap.index], dtype=np.int64)
            D = dmap.to_numpy().astype(np.float64)
            assert (D == D.T).all()
            out["D_sha256"] = np.array(hashlib.sha256(np.ascontiguousarray(D).tobytes()).hexdigest())
            if store_D:
                out["D"] = D
            else:
                pick = sorted(set([0, 1, D.shape[0] // 3, D.shape[0] // 2, D.shape[0] - 2, D.shape[0] - 1]))
                out["D_row_ids"] = np.array(pick, dtype=np.int64)
                out["D_rows"] = D[pick]
            out["joins"] = np.array(solver.recorded_joins, dtype=np.int32).reshape(-1, 2)
        topo = tree.get_tree_topology()
        hs, ss, ds = treecmp.clade_table(topo, leaf_index)
        tag = "collapsed" if collapse else "plain"
        out[f"clades_{tag}_hash"] = hs
        out[f"clades_{tag}_size"] = ss
        out[f"clades_{tag}_depth"] = ds
        out[f"n_nodes_{tag}"] = np.int64(topo.number_of_nodes())
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: n={n} m={m} algo={algo} joins={len(out['joins'])} "
          f"{os.path.getsize(path)/1024:.0f} KiB in {time.time()-t0:.1f}s", flush=True)


def cases():
    ref = ref_loader.load_reference()
    cs = {}

    def sim_case(tag, n_cells, m, S, seed, rate):
        cache = {}

        def get():
            if "v" not in cache:
                cache["v"] = simulate(ref, n_cells, m, S, seed, rate)
            return cache["v"]
        return get

    s64 = sim_case("sim64", 64, 40, 50, 0, 0.1)
    s256 = sim_case("sim256", 256, 40, 50, 1, 0.7)
    s1024 = sim_case("sim1024", 1024, 40, 50, 0, 0.1)
    for algo in ("nj", "upgma"):
        cs[f"sim64_s0_{algo}_priors"] = lambda a=algo: run_case(f"sim64_s0_{a}_priors", *s64(), a)
        cs[f"sim64_s0_{algo}_plain"] = lambda a=algo: run_case(f"sim64_s0_{a}_plain", s64()[0], None, a)
        cs[f"sim256_s1_{algo}_priors"] = lambda a=algo: run_case(f"sim256_s1_{a}_priors", *s256(), a)
        cs[f"sim256_s1_{algo}_plain"] = lambda a=algo: run_case(f"sim256_s1_{a}_plain", s256()[0], None, a, store_D=False)
        cs[f"ties120_{algo}_plain"] = lambda a=algo: run_case(
            f"ties120_{a}_plain", random_tie_heavy(120, 8, 3, 7, 0.2, 0.15)[0], None, a)
        cs[f"ties120_{algo}_priors"] = lambda a=algo: run_case(
            f"ties120_{a}_priors", *random_tie_heavy(120, 8, 3, 7, 0.2, 0.15), a)
    cs["sim64_s0_nj_inverse"] = lambda: run_case("sim64_s0_nj_inverse", *s64(), "nj", "inverse")
    cs["sim64_s0_upgma_sqrtinv"] = lambda: run_case("sim64_s0_upgma_sqrtinv", *s64(), "upgma", "square_root_inverse")
    # config C1 of BASELINE.json: 2^10 leaves x 40 characters, priors
    cs["sim1024_s0_nj_priors"] = lambda: run_case("sim1024_s0_nj_priors", *s1024(), "nj", store_D=False)
    cs["sim1024_s0_upgma_plain"] = lambda: run_case("sim1024_s0_upgma_plain", s1024()[0], None, "upgma", store_D=False)
    return cs


def main(argv):
    if not ref_loader.reference_available():
        print("reference not available; nothing generated")
        return 1
    cs = cases()
    want = argv[1:] or list(cs)
    for name in want:
        cs[name]()
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
