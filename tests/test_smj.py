"""SharedMutationJoiningSolver.

CPU: the C oracle's loop against join sequences recorded from the reference itself
(oracle/make_golden_smj.py), and the host side of the B200 class (tree building, populate, collapse) against
the reference's final clade tables, with the device call replaced by the oracle.  GPU: ``cas_smj_solve``
join lists == the reference's, whole ``solve()`` trees == the reference's."""
import glob
import os

import numpy as np
import pandas as pd
import pytest

from conftest import GOLDEN_DIR
from oracle import pyoracle as po
from oracle import treecmp


def smj_goldens():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "smj", "*.npz")))


class SmjGolden:
    def __init__(self, path):
        z = np.load(path)
        self.z = z
        self.cm = z["cm"].astype(np.int64)
        self.missing = int(z["missing"])
        self.function = str(z["function"])
        self.priors = z["priors"] if z["priors"].size else None
        self.joins = z["joins"]
        self.names = [f"s{i}" for i in range(self.cm.shape[0])]

    @property
    def table(self):
        if self.priors is None:
            return None
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.nan_to_num(-np.log(self.priors), nan=0.0, posinf=0.0)

    def prior_dict(self):
        if self.priors is None:
            return None
        p = self.priors
        return {c: {s: float(p[c, s]) for s in range(p.shape[1]) if not np.isnan(p[c, s])} for c in range(p.shape[0])}

    def tree(self):
        import cassiopeia_b200 as cb

        return cb.CassiopeiaTree(character_matrix=pd.DataFrame(self.cm, index=self.names), priors=self.prior_dict(),
                                 missing_state_indicator=self.missing)

    def check_tree(self, tree, collapsed):
        tag = "collapsed" if collapsed else "plain"
        leaf_index = {nm: i for i, nm in enumerate(self.names)}
        hs, ss, ds = treecmp.clade_table(tree.get_tree_topology(), leaf_index)
        assert np.array_equal(hs, self.z[f"clades_{tag}_hash"])
        assert np.array_equal(ss, self.z[f"clades_{tag}_size"])
        assert np.array_equal(ds, self.z[f"clades_{tag}_depth"])
        assert tree.get_tree_topology().number_of_nodes() == int(self.z[f"n_nodes_{tag}"])


def solver_for(g):
    import cassiopeia_b200 as cb
    from cassiopeia_b200.solver import dissimilarity_functions as df

    return cb.SharedMutationJoiningSolver(similarity_function=getattr(df, g.function))


@pytest.mark.parametrize("path", smj_goldens(), ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_loop_matches_reference_joins(path):
    g = SmjGolden(path)
    assert g.joins.shape == (g.cm.shape[0] - 1, 2)
    assert np.array_equal(po.smj(g.function, g.cm, g.missing, g.table), g.joins)


@pytest.mark.parametrize("path", smj_goldens(), ids=lambda p: os.path.basename(p)[:-4])
@pytest.mark.parametrize("collapse", [False, True])
def test_host_side_builds_the_reference_tree(path, collapse, monkeypatch):
    """solve() with the device call replaced by the oracle: same clades, depths and node count as the
    reference's tree (tests the host logic without a GPU)."""
    from cassiopeia_b200 import engine

    g = SmjGolden(path)

    def fake(codes, missing_code, table, n_states, function, max_iters=-1):
        joins = po.smj(function, np.asarray(codes, dtype=np.int64), missing_code, table)
        return joins[:-1], joins[-1]

    monkeypatch.setattr(engine, "smj_solve", fake)
    tree = g.tree()
    solver_for(g).solve(tree, collapse_mutationless_edges=collapse)
    g.check_tree(tree, collapse)


def test_constructor_rejects_functions_without_a_kernel():
    import cassiopeia_b200 as cb
    from cassiopeia_b200.mixins import SharedMutationJoiningSolverError

    with pytest.raises(SharedMutationJoiningSolverError):
        cb.SharedMutationJoiningSolver(similarity_function=lambda a, b, m, w: 0.0)
    cb.SharedMutationJoiningSolver()  # default: hamming_similarity_without_missing


# ------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("path", smj_goldens(), ids=lambda p: os.path.basename(p)[:-4])
def test_gpu_loop_matches_reference_joins(path):
    from cassiopeia_b200 import engine

    g = SmjGolden(path)
    n_states = g.table.shape[1] - 1 if g.table is not None else max(int(g.cm.max()), 1)
    joins, last2 = engine.smj_solve(g.cm.astype(np.int32), g.missing, g.table, n_states, g.function)
    assert np.array_equal(joins, g.joins[:-1])
    assert np.array_equal(last2, g.joins[-1])


@pytest.mark.gpu
@pytest.mark.parametrize("path", smj_goldens(), ids=lambda p: os.path.basename(p)[:-4])
@pytest.mark.parametrize("collapse", [False, True])
def test_gpu_solve_builds_the_reference_tree(path, collapse):
    g = SmjGolden(path)
    tree = g.tree()
    solver = solver_for(g)
    solver.solve(tree, collapse_mutationless_edges=collapse)
    assert np.array_equal(solver.last_join_list[0], g.joins)
    g.check_tree(tree, collapse)


@pytest.mark.gpu
def test_gpu_loop_matches_oracle_at_size():
    from cassiopeia_b200 import engine
    from cassiopeia_b200.synthetic import simulate_character_matrix

    n, m, S = 900, 40, 20
    cm, priors, table = simulate_character_matrix(n, m, S, seed=9, mutation_rate=0.7, missing_fraction=0.15)
    with np.errstate(invalid="ignore", divide="ignore"):
        w = np.nan_to_num(-np.log(table), nan=0.0, posinf=0.0)
    for fn, ww in (("hamming_similarity_without_missing", None), ("hamming_similarity_without_missing", w),
                   ("weighted_hamming_similarity", w)):
        joins, last2 = engine.smj_solve(cm, -1, ww, S, fn)
        want = po.smj(fn, cm.astype(np.int64), -1, ww)
        assert np.array_equal(joins, want[:-1]) and np.array_equal(last2, want[-1]), fn
