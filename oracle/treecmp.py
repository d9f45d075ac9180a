"""TEST INFRASTRUCTURE ONLY -- topology comparison helpers (no ete3 here).

A rooted tree is summarised by the multiset of its clades; each clade (set of
leaves under an internal node) is hashed additively over 64-bit mixed leaf
indices, so the summary is O(#nodes) numbers and can be stored as a fixture.
Two rooted trees on the same leaf set are topologically identical iff their
clade sets agree; the unrooted Robinson-Foulds distance is the symmetric
difference of their bipartition sets.
"""
from __future__ import annotations

from typing import Dict, Iterable, Sequence, Tuple

import networkx as nx
import numpy as np

_MASK = (1 << 64) - 1


def _mix(x: int) -> int:
    """splitmix64 finaliser."""
    x = (x + 0x9E3779B97F4A7C15) & _MASK
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _MASK
    return x ^ (x >> 31)


def clade_table(tree: nx.DiGraph, leaf_index: Dict[str, int]) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(hashes, sizes, depths) of every internal node, sorted by (hash, size).

    depth = number of edges from the root (the reference sets every edge length
    to 1 in populate_tree, CassiopeiaTree.py:193-196).
    """
    roots = [v for v in tree.nodes if tree.in_degree(v) == 0]
    assert len(roots) == 1, f"expected one root, found {len(roots)}"
    root = roots[0]
    h: Dict[str, int] = {}
    sz: Dict[str, int] = {}
    depth = {root: 0}
    order = list(nx.dfs_preorder_nodes(tree, root))
    for v in order:
        for c in tree.successors(v):
            depth[c] = depth[v] + 1
    for v in reversed(order):
        kids = list(tree.successors(v))
        if not kids:
            h[v] = _mix(leaf_index[str(v)] + 1)
            sz[v] = 1
        else:
            h[v] = sum(h[c] for c in kids) & _MASK
            sz[v] = sum(sz[c] for c in kids)
    internal = [v for v in order if tree.out_degree(v) > 0]
    rows = sorted((h[v], sz[v], depth[v]) for v in internal)
    if not rows:
        z = np.zeros(0, dtype=np.uint64)
        return z, z.astype(np.int64), z.astype(np.int64)
    hs = np.array([r[0] for r in rows], dtype=np.uint64)
    ss = np.array([r[1] for r in rows], dtype=np.int64)
    ds = np.array([r[2] for r in rows], dtype=np.int64)
    return hs, ss, ds


def rf_distance(tree_a: nx.DiGraph, tree_b: nx.DiGraph, leaf_index: Dict[str, int]) -> int:
    """Unrooted Robinson-Foulds distance via hashed bipartitions."""
    total = sum(_mix(i + 1) for i in leaf_index.values()) & _MASK
    n = len(leaf_index)

    def splits(tree):
        hs, ss, _ = clade_table(tree, leaf_index)
        out = set()
        for hv, s in zip(hs.tolist(), ss.tolist()):
            if s <= 1 or s >= n - 1:
                continue  # trivial split
            comp = (total - hv) & _MASK
            out.add(min((hv, s), (comp, n - s)))
        return out

    return len(splits(tree_a) ^ splits(tree_b))


def tree_from_joins(joins: Iterable[Sequence[int]], n: int, names: Sequence[str], last2=None,
                    root_name: str = "root", mode: str = "upgma") -> nx.DiGraph:
    """Rooted tree from an id-space join list (leaf p -> names[p], t-th join -> id n+t).

    mode "upgma": a new root adopts the last two clusters (UPGMASolver.py:107-116).
    mode "nj": edge between the last two, oriented away from `root_name`, which
    must be one of `names` (NeighborJoiningSolver.py:115-121).
    """
    g = nx.Graph()
    label = {p: names[p] for p in range(n)}
    g.add_nodes_from(names)
    nxt = n
    for a, b in joins:
        label[nxt] = f"internal{nxt}"
        g.add_edges_from([(label[nxt], label[int(a)]), (label[nxt], label[int(b)])])
        nxt += 1
    if last2 is None:
        joined = set(int(x) for ab in joins for x in ab)
        last2 = [i for i in range(nxt) if i not in joined]
    assert len(last2) == 2
    if mode == "upgma":
        g.add_edges_from([(root_name, label[int(last2[0])]), (root_name, label[int(last2[1])])])
    else:
        g.add_edge(label[int(last2[0])], label[int(last2[1])])
    d = nx.DiGraph()
    d.add_node(root_name)
    for e in nx.dfs_edges(g, source=root_name):
        d.add_edge(*e)
    return d
