"""Array-based solver tail (cassiopeia_b200/tree_tail.py) == the generic networkx chain that mirrors reference
DistanceSolver.py:181-231 (root_tree, populate_tree, collapse_unifurcations, collapse_mutationless_edges):
same node order, successor order, branch lengths, times and character states, on random join lists."""
import numpy as np
import pandas as pd
import pytest

import cassiopeia_b200 as cb


def random_joins(n, rng, caterpillar=False):
    """A valid id-space join list over n map rows: n-2 joins, two ids remain."""
    live = list(range(n))
    joins = []
    nxt = n
    while len(live) > 2:
        if caterpillar:
            i, j = 0, 1
        else:
            i, j = sorted(rng.choice(len(live), size=2, replace=False).tolist())
        a, b = live[i], live[j]
        joins.append((min(a, b), max(a, b)))
        live = [x for k, x in enumerate(live) if k != i and k != j] + [nxt]
        nxt += 1
    return np.asarray(joins, dtype=np.int32).reshape(-1, 2), np.asarray(sorted(live), dtype=np.int32)


def character_matrix(n, m, rng, low_information):
    if low_information:
        cm = rng.choice([0, 0, 0, 1, 2, -1], size=(n, m))
        cm[rng.random(n) < 0.4] = cm[0]            # many duplicates -> many mutationless edges
    else:
        cm = rng.integers(-1, 6, size=(n, m))
    return cm.astype(np.int64)


def snapshot(tree):
    g = tree._network
    return {
        "nodes": list(g.nodes),
        "succ": {v: list(g.successors(v)) for v in g.nodes},
        "length": {(u, v): d["length"] for u, v, d in g.edges(data=True)},
        "time": {v: g.nodes[v]["time"] for v in g.nodes},
        "states": {v: list(g.nodes[v]["character_states"]) for v in g.nodes},
    }


def deterministic_names(monkeypatch):
    """Identical internal-node names on both routes."""
    from cassiopeia_b200.solver import solver_utilities

    def gen():
        i = 0
        while True:
            yield f"internal{i}"
            i += 1

    monkeypatch.setattr(solver_utilities, "node_name_generator", gen)


@pytest.mark.parametrize("algo", ["nj_implicit", "nj_explicit", "upgma"])
@pytest.mark.parametrize("collapse", [False, True])
@pytest.mark.parametrize("seed", range(6))
def test_array_tail_equals_generic_chain(algo, collapse, seed, monkeypatch):
    rng = np.random.default_rng(1000 * seed + 17)
    n = int(rng.integers(3, 60))
    m = int(rng.integers(1, 9))
    cm = character_matrix(n, m, rng, low_information=seed % 2 == 0)
    leaves = [f"c{i}" for i in range(n)]
    deterministic_names(monkeypatch)
    snaps = []
    for fast in (False, True):
        tree = cb.CassiopeiaTree(character_matrix=pd.DataFrame(cm, index=leaves))
        if algo == "upgma":
            solver = cb.UPGMASolver()
            names = list(leaves)
            tree.root_sample_name = "root"
        elif algo == "nj_implicit":
            solver = cb.NeighborJoiningSolver(add_root=True)
            names = leaves + ["root"]
            tree.root_sample_name = "root"
            rooted = tree.character_matrix.copy()
            rooted.loc["root"] = 0
            tree.character_matrix = rooted
        else:
            solver = cb.NeighborJoiningSolver(add_root=False)
            names = list(leaves)
            tree.root_sample_name = leaves[int(rng.integers(0, n))] if not snaps else snaps_root
        snaps_root = tree.root_sample_name
        jrng = np.random.default_rng(seed)
        joins, last2 = random_joins(len(names), jrng, caterpillar=seed == 5)
        solver.fast_tail = fast
        solver.finish_tree(tree, names, joins, last2, None, collapse)
        snaps.append(snapshot(tree))
    generic, fast = snaps
    assert fast["nodes"] == generic["nodes"]
    assert fast["succ"] == generic["succ"]
    assert fast["length"] == generic["length"]
    assert fast["time"] == generic["time"]
    assert fast["states"] == generic["states"]


def test_root_in_last_edge_and_overridden_hook(monkeypatch):
    """The root sample may be one of the two remaining nodes; a subclass that overrides root_tree keeps the
    generic route."""
    cm = np.array([[1, 0], [1, 2], [0, 2], [0, 0]], dtype=np.int64)
    names = ["a", "b", "c", "r"]
    joins = np.array([[0, 1], [2, 4]], dtype=np.int32)   # (a,b)->4, (c,4)->5 ; remaining: r(3), 5
    last2 = np.array([3, 5], dtype=np.int32)
    deterministic_names(monkeypatch)
    out = []
    for fast in (False, True):
        tree = cb.CassiopeiaTree(character_matrix=pd.DataFrame(cm, index=names), root_sample_name="r")
        s = cb.NeighborJoiningSolver()
        s.fast_tail = fast
        s.finish_tree(tree, names, joins, last2, None, True)
        out.append((sorted(tree.leaves), sorted(map(sorted, (tree.children(v) for v in tree.internal_nodes))),
                    tree.get_times()))
        assert tree.root == "r"
    assert out[0][0] == out[1][0] and out[0][1] == out[1][1]
    assert sorted(out[0][2].values()) == sorted(out[1][2].values())

    calls = []

    class Custom(cb.NeighborJoiningSolver):
        def root_tree(self, tree, root_sample, remaining_samples):
            calls.append(root_sample)
            return super().root_tree(tree, root_sample, remaining_samples)

    tree = cb.CassiopeiaTree(character_matrix=pd.DataFrame(cm, index=names), root_sample_name="r")
    Custom().finish_tree(tree, names, joins, last2, None, False)
    assert calls == ["r"]


@pytest.mark.parametrize("n,algo,seed", [(3, "nj", 0), (3, "upgma", 1), (40, "nj", 2), (257, "upgma", 3),
                                         (2000, "nj", 4)])
def test_native_passes_equal_interpreter_passes(n, algo, seed):
    """cas_tail_preorder / cas_tail_heights / cas_tail_successor_lists against their Python twins, on random
    join lists (including a caterpillar, whose depth is n) and random removal flags."""
    from cassiopeia_b200 import tree_tail

    rng = np.random.default_rng(seed)
    for shape in ("random", "caterpillar"):
        live = list(range(n))
        joins = []
        for t in range(n - 2):
            if shape == "random":
                i, j = sorted(rng.choice(len(live), size=2, replace=False).tolist())
            else:
                i, j = len(live) - 2, len(live) - 1
            joins.append((live[i], live[j]))
            del live[j]
            del live[i]
            live.append(n + t)
        rt = tree_tail.RootedJoinTree(n, np.asarray(joins).reshape(-1, 2), live, algo,
                                      None if algo == "upgma" else int(rng.integers(0, n)))
        order, depth = rt.preorder()
        order_py, depth_py = rt.preorder_py()
        assert np.array_equal(order, order_py) and np.array_equal(depth, depth_py)
        assert np.array_equal(rt.heights(), rt.heights_py())
        removed = rng.random(rt.M) < 0.4
        removed[rt.root] = False
        removed[rt.child0 < 0] = False  # leaves are never spliced
        got = tree_tail.successor_lists(order, rt.child0, rt.child1, removed)
        want = tree_tail.successor_lists_py(order, rt.child0, rt.child1, removed)
        kept = np.flatnonzero(~removed)
        for a, b in zip(got, want):
            # head / tail of removed leaves-of-lists and next of list tails are -1 in both
            assert np.array_equal(a, b)
        assert (got[0][kept[rt.child0[kept] >= 0]] >= -1).all()


@pytest.mark.parametrize("algo", ["nj", "upgma"])
@pytest.mark.parametrize("collapse", [False, True])
def test_c_dictionary_filler_equals_interpreter_loop(algo, collapse, monkeypatch):
    """`_graphfill.fill` (csrc/graphfill.c) and the interpreter loop of build_final_tree leave identical
    networkx dictionaries: key order of all three, attribute values, and one shared attribute dict per edge."""
    from cassiopeia_b200 import tree_tail

    if tree_tail._graphfill is None:
        pytest.skip("cassiopeia_b200/_graphfill.so was not built (make -C cassiopeia_b200/csrc needs Python.h)")
    filler = tree_tail._graphfill
    rng = np.random.default_rng(3)
    n, m = 300, 6
    cm = character_matrix(n, m, rng, low_information=True)
    names = [f"c{i}" for i in range(n)]
    joins, last2 = random_joins(n, rng)
    internal = [f"i{t}" for t in range(n - 2)]
    graphs = []
    for use_c in (True, False):
        if not use_c:
            monkeypatch.setattr(tree_tail, "_graphfill", None)
        g, states = tree_tail.build_final_tree(names, internal, joins, last2, algo, "c7" if algo == "nj" else None,
                                               cm, -1, collapse)
        graphs.append((g, states))
    (gc_, sc), (gp, sp) = graphs
    assert sc == sp and gc_.graph == gp.graph
    assert list(gc_._node) == list(gp._node) and gc_._node == gp._node
    assert list(gc_._succ) == list(gp._succ) and list(gc_._pred) == list(gp._pred)
    for u in gp._node:
        assert list(gc_._succ[u].items()) == list(gp._succ[u].items())
        assert list(gc_._pred[u].items()) == list(gp._pred[u].items())
        for v, attrs in gc_._succ[u].items():
            assert gc_._pred[v][u] is attrs
    with pytest.raises(ValueError):
        filler.fill({}, {}, {}, ["a", "a"], np.array([0, 1]), np.array([-1, -1]), np.array([-1, -1]),
                                  np.array([0, 1]))
    with pytest.raises(TypeError):
        filler.fill({}, {}, {}, ["a"], np.array([0], dtype=np.int32), np.array([-1]), np.array([-1]),
                                  np.array([0]))
