"""Drop-in proof against the REFERENCE's own objects and tests (CPU box only: needs /root/reference).

* the reference's unmodified test files ``test/solver_tests/upgma_test.py`` and ``neighborjoining_solver_test.py`` run
  with ``cas.solver.UPGMASolver`` / ``cas.solver.NeighborJoiningSolver`` re-pointed at the B200 classes and the
  reference's real ``cas.data.CassiopeiaTree`` (cassiopeia/data/CassiopeiaTree.py:39);
* B200 solvers on reference trees give the reference solvers' trees (clade sets), with and without
  ``collapse_mutationless_edges``, explicit and implicit roots, stored maps, layers;
* a B200 solver as ``joining_solver`` of the reference's ``PercolationSolver`` (PercolationSolver.py:299).

The CUDA library cannot run here, so the engine's compute entry points are replaced, in these tests only, by the
oracle (tests/fake_engine.py): what is under test is the product's HOST side against the reference's objects."""
import importlib
import importlib.util
import os
import unittest

import re

import numpy as np
import pandas as pd
import pytest

from oracle import ref_loader, treecmp

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="needs the reference checkout")

REF_TESTS = os.path.join(ref_loader.REFERENCE_ROOT, "test", "solver_tests")

#: reference tests that use a user-supplied Python dissimilarity function without a precomputed map: the product
#: raises for those by design (SURVEY 8(b) "unsupported inputs must raise": there is no CPU fallback)
NEEDS_PYTHON_CALLABLE = {"test_duplicate_sample_neighbor_joining", "test_pp_solver",
                         "test_setup_root_finder_missing_dissimilarity_map"}


@pytest.fixture()
def ref(monkeypatch):
    import fake_engine

    r = ref_loader.load_reference()
    fake_engine.install(monkeypatch)
    return r


def _run_reference_test_file(ref, monkeypatch, fname, **repoint):
    solver_pkg = ref.cassiopeia.solver
    for name, cls in repoint.items():
        monkeypatch.setattr(solver_pkg, name, cls)
    # in the installed package `cas.solver.DistanceSolver` is the sub-module (the tests reach its DistanceSolverError)
    monkeypatch.setattr(solver_pkg, "DistanceSolver", importlib.import_module("cassiopeia.solver.DistanceSolver"))
    spec = importlib.util.spec_from_file_location("reftest_" + fname[:-3], os.path.join(REF_TESTS, fname))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    suite = unittest.defaultTestLoader.loadTestsFromModule(mod)
    result = unittest.TestResult()
    suite.run(result)
    bad = {t.id().split(".")[-1]: tb for t, tb in result.failures + result.errors}
    return result.testsRun, bad


def test_reference_upgma_test_file_passes_against_the_b200_class(ref, monkeypatch):
    import cassiopeia_b200 as cb

    ran, bad = _run_reference_test_file(ref, monkeypatch, "upgma_test.py", UPGMASolver=cb.UPGMASolver)
    assert ran == 7 and not bad, bad


def test_reference_nj_test_file_passes_against_the_b200_class(ref, monkeypatch):
    import cassiopeia_b200 as cb

    ran, bad = _run_reference_test_file(ref, monkeypatch, "neighborjoining_solver_test.py",
                                        NeighborJoiningSolver=cb.NeighborJoiningSolver)
    assert ran == 10
    assert set(bad) <= NEEDS_PYTHON_CALLABLE, {k: v[-400:] for k, v in bad.items() if k not in NEEDS_PYTHON_CALLABLE}
    for name in NEEDS_PYTHON_CALLABLE & set(bad):
        assert "no CPU fallback" in bad[name], bad[name][-400:]


def _inputs(n=60, m=12, seed=3, missing=0.15, dup=True):
    rng = np.random.default_rng(seed)
    cm = rng.integers(0, 5, size=(n, m))
    cm[rng.random((n, m)) < 0.55] = 0
    cm[rng.random((n, m)) < missing] = -1
    if dup:
        cm[7] = cm[3]
        cm[21] = cm[3]
    frame = pd.DataFrame(cm, index=[f"s{i}" for i in range(n)])
    priors = {c: {s: float(p) for s, p in zip(range(1, 5), rng.dirichlet(np.ones(4)))} for c in range(m)}
    return frame, priors


def _clades(tree):
    """(clade table: hashes, sizes, depths of every internal node; sorted leaf names)."""
    leaves = sorted(tree.leaves)
    index = {name: i for i, name in enumerate(leaves)}
    return treecmp.clade_table(tree.get_tree_topology(), index), leaves


def _same(a, b):
    return all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("collapse", [False, True])
@pytest.mark.parametrize("use_priors", [False, True])
@pytest.mark.parametrize("algo", ["nj", "upgma"])
def test_b200_solver_on_reference_tree_equals_reference_solver(ref, algo, use_priors, collapse):
    """`solver.solve(cas_tree)` with the reference's CassiopeiaTree: same clades as the reference's own solver."""
    import cassiopeia_b200 as cb

    frame, priors = _inputs()
    trees = []
    for impl in ("ref", "b200"):
        tree = ref.CassiopeiaTree(character_matrix=frame.copy(), priors=priors if use_priors else None)
        if algo == "nj":
            solver = (ref.NeighborJoiningSolver if impl == "ref" else cb.NeighborJoiningSolver)(add_root=True)
        else:
            solver = (ref.UPGMASolver if impl == "ref" else cb.UPGMASolver)()
        solver.solve(tree, collapse_mutationless_edges=collapse)
        assert isinstance(tree, ref.CassiopeiaTree)
        trees.append(tree)
    (a, la), (b, lb) = _clades(trees[0]), _clades(trees[1])
    assert la == lb
    assert _same(a, b)
    # side effects the reference's callers rely on: root sample recorded, root row gone from the character matrix,
    # the stored map readable as a frame in the reference's row order
    assert trees[1].root_sample_name == trees[0].root_sample_name
    assert list(trees[1].character_matrix.index) == list(trees[0].character_matrix.index)
    pd.testing.assert_frame_equal(trees[1].get_dissimilarity_map(), trees[0].get_dissimilarity_map())


def test_b200_nj_with_explicit_root_and_stored_map_on_reference_tree(ref):
    import cassiopeia_b200 as cb

    frame, priors = _inputs(n=40, seed=9)
    out = []
    for cls in (ref.NeighborJoiningSolver, cb.NeighborJoiningSolver):
        tree = ref.CassiopeiaTree(character_matrix=frame.copy(), priors=priors, root_sample_name="s5")
        tree.compute_dissimilarity_map(ref.dissimilarity_functions.weighted_hamming_distance, "negative_log")
        cls(add_root=False).solve(tree)
        out.append(_clades(tree))
    assert out[0][1] == out[1][1] and _same(out[0][0], out[1][0])


def test_b200_solver_as_joining_solver_of_the_reference_percolation_solver(ref):
    """PercolationSolver.py:299 calls `self.joining_solver.solve(lca_tree, collapse_mutationless_edges=False)`."""
    import cassiopeia_b200 as cb

    perc = importlib.import_module("cassiopeia.solver.PercolationSolver").PercolationSolver
    frame, priors = _inputs(n=48, m=10, seed=11, missing=0.05, dup=False)
    out = []
    for cls in (ref.NeighborJoiningSolver, cb.NeighborJoiningSolver):
        tree = ref.CassiopeiaTree(character_matrix=frame.copy(), priors=priors)
        joining = cls(dissimilarity_function=ref.dissimilarity_functions.weighted_hamming_distance, add_root=True)
        solver = perc(joining_solver=joining)
        solver.solve(tree)
        out.append(_clades(tree))
    assert out[0][1] == out[1][1] and _same(out[0][0], out[1][0])


@pytest.mark.parametrize("fn_name", ["weighted_hamming_distance", "hamming_similarity_without_missing",
                                     "hamming_similarity_normalized_over_missing", "weighted_hamming_similarity",
                                     "exponential_negative_hamming_distance"])
@pytest.mark.parametrize("use_priors", [False, True])
def test_tree_level_map_with_duplicate_cells_equals_reference_for_every_pair_function(ref, fn_name, use_priors):
    """The reference evaluates the function on the distinct rows and copies rows into the duplicates
    (CassiopeiaTree.py:1920-1951): cells with identical rows are at 0 from each other for EVERY function, also the
    similarities (f(x, x) != 0) and exp(-d) (f(x, x) = 1); a state carried only by duplicated cells is never looked
    up in the priors."""
    import cassiopeia_b200 as cb
    from cassiopeia_b200.solver import dissimilarity_functions as b200_fns

    frame, priors = _inputs(n=40, m=8, seed=21, missing=0.1, dup=True)
    frame.iloc[30] = frame.iloc[12]
    frame.iloc[31] = frame.iloc[12]
    # a state (9) carried only by two identical cells, without a prior: the reference never looks it up
    frame.iloc[33, 2] = 9
    frame.iloc[34] = frame.iloc[33]
    pri = priors if use_priors else None
    want_tree = ref.CassiopeiaTree(character_matrix=frame.copy(), priors=pri)
    got_tree = cb.CassiopeiaTree(character_matrix=frame.copy(), priors=pri)
    try:
        want_tree.compute_dissimilarity_map(getattr(ref.dissimilarity_functions, fn_name), "negative_log")
    except KeyError as e:
        # the distance functions look a state up when it takes part in a mismatch: state 9 does (against the other
        # distinct rows), so the reference raises for them -- and so must the product, with the same key
        with pytest.raises(KeyError) as got_err:
            got_tree.compute_dissimilarity_map(getattr(b200_fns, fn_name), "negative_log")
        assert got_err.value.args == e.args
        assert use_priors and fn_name in ("weighted_hamming_distance", "exponential_negative_hamming_distance")
        return
    got_tree.compute_dissimilarity_map(getattr(b200_fns, fn_name), "negative_log")
    want, got = want_tree.get_dissimilarity_map(), got_tree.get_dissimilarity_map()
    assert list(want.index) == list(got.index)
    np.testing.assert_array_equal(want.to_numpy(), got.to_numpy())


@pytest.mark.parametrize("algo", ["nj", "upgma"])
@pytest.mark.parametrize("collapse", [False, True])
def test_tail_worker_thread_does_not_change_the_tree(algo, collapse, monkeypatch):
    """solve() computes the join-independent inputs of the host tail in a worker thread while the loop runs
    (DistanceSolver._tail_precompute); with the thread switched off (CAS_TAIL_THREAD=0) the tree must be the same:
    nodes in the same order, the same edges with the same lengths, times and character states."""
    import fake_engine
    import cassiopeia_b200 as cb
    from cassiopeia_b200.synthetic import simulate_character_matrix

    fake_engine.install(monkeypatch)
    cm, priors, _ = simulate_character_matrix(120, 12, 5, seed=4, mutation_rate=0.4, missing_fraction=0.15)
    cm = np.vstack([cm, cm[:15]])  # duplicated cells
    names = [f"c{i}" for i in range(cm.shape[0])]
    trees = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("CAS_TAIL_THREAD", flag)
        tree = cb.CassiopeiaTree(character_matrix=pd.DataFrame(cm, index=names), priors=priors)
        solver = cb.NeighborJoiningSolver(add_root=True) if algo == "nj" else cb.UPGMASolver()
        solver.solve(tree, collapse_mutationless_edges=collapse)
        g = tree.get_tree_topology()
        # internal node names carry a per-solve random prefix (unique, not reproducible): keep the counter only
        norm = lambda name: re.sub(r"^cassiopeia_internal_node[0-9a-f]{12}", "internal_", name)
        trees[flag] = ([norm(x) for x in g.nodes], [(norm(u), norm(v), d.get("length")) for u, v, d in g.edges(data=True)],
                       {norm(x): tree.get_character_states(x) for x in g.nodes},
                       {norm(x): tree.get_time(x) for x in g.nodes}, tree.character_matrix.index.tolist())
    assert trees["1"] == trees["0"]
