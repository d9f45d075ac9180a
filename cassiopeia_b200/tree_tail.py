"""Array-based tail of ``DistanceSolver.solve``.

The reference finishes a solve with a chain of per-node Python/networkx passes
(cassiopeia/solver/DistanceSolver.py:207-231): ``root_tree`` (NeighborJoiningSolver.py:100-121,
UPGMASolver.py:89-116), ``populate_tree`` (CassiopeiaTree.py:141-203), ``collapse_unifurcations``
(:1669-1715) and, optionally, ``reconstruct_ancestral_characters`` (:648-675 with
data/utilities.py:29-97) + ``collapse_mutationless_edges`` (:1717-1775).  On 200 000 cells those passes take
longer than the sharded GPU loop's exchange overhead.  This module computes the *result* of that chain from
the id-space join list with integer arrays:

* the unrooted join graph is never materialised: node ``N+t`` has adjacency ``[a_t, b_t, up]`` in networkx
  insertion order, so the depth-first orientation from the root, the child order and the depth ("time") of
  every node follow from the ``up`` pointers and the root-to-last-edge path, whose edges are the only ones
  that get reversed;
* Camin-Sokal LCA states are a semilattice (missing = identity, 0 = absorbing, unequal cut states -> 0), so
  they are evaluated level by level (node height) with vectorised numpy over all nodes of a level;
* an edge is mutationless iff the child is internal and its state row equals its parent's, equality is
  transitive along chains, so the contracted tree is "every kept node hangs from its nearest kept ancestor";
  the reference's successor order (un-spliced children first, then the spliced children's lists, recursively)
  is reproduced with O(1) linked-list concatenation;
* one final ``nx.DiGraph`` is built in the reference's node order (depth-first preorder of the rooted binary
  tree, minus removed nodes) with ``length`` attributes preset, so ``populate_tree`` (the tree object's own,
  reference or ours) yields the same lengths and times as the reference's collapse chain.

``tests/test_tree_tail.py`` checks node order, successor order, lengths, times and character states against
the generic networkx chain on random join lists.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import networkx as nx
import numpy as np

from . import _lib

try:  # CPython extension built next to libcas_b200.so (csrc/graphfill.c); the interpreter loop is the fallback
    from . import _graphfill
except ImportError:  # pragma: no cover
    _graphfill = None


#: graph attribute set on trees whose node attributes (`character_states`, `time`) and edge `length`s are
#: already what `populate_tree` would compute
PREPARED = "cas_b200_prepared"


def _lca2(a: np.ndarray, b: np.ndarray, missing: int) -> np.ndarray:
    """Camin-Sokal LCA of two state rows (reference data/utilities.py:29-97, unambiguous states)."""
    return np.where(a == missing, b, np.where(b == missing, a, np.where(a == b, a, 0)))


class RootedJoinTree:
    """Rooted binary tree of a join list in id space.

    ids: leaves 0..N-1 (map rows), join t -> N+t, UPGMA's extra root -> N+T.
    ``child0/child1`` follow networkx adjacency order; ``parent[root] = -1``.
    """

    def __init__(self, n_leaves: int, joins: np.ndarray, last2: Sequence[int], algo: str,
                 root_leaf: Optional[int]):
        N = int(n_leaves)
        joins = np.asarray(joins, dtype=np.int64).reshape(-1, 2)
        T = joins.shape[0]
        if T != N - 2 or N < 3:
            raise ValueError("RootedJoinTree needs a complete join list of at least 3 samples")
        M = N + T + (1 if algo == "upgma" else 0)
        a, b = joins[:, 0], joins[:, 1]
        new = N + np.arange(T, dtype=np.int64)
        up = np.full(M, -1, dtype=np.int64)   # the join that consumes the node
        up[a] = new
        up[b] = new
        r0, r1 = int(last2[0]), int(last2[1])
        if up[r0] != -1 or up[r1] != -1 or np.count_nonzero(up[: N + T] == -1) != 2:
            raise ValueError("join list and remaining samples do not describe one tree")
        parent = up.copy()
        c0 = np.full(M, -1, dtype=np.int64)
        c1 = np.full(M, -1, dtype=np.int64)
        c0[new] = a
        c1[new] = b
        if algo == "upgma":
            root = N + T
            parent[r0] = root
            parent[r1] = root
            c0[root], c1[root] = r0, r1
        else:
            if root_leaf is None or not (0 <= root_leaf < N):
                raise ValueError("the root sample must be one of the map's rows")
            root = int(root_leaf)
            # third neighbour in adjacency order: the consuming join, or the partner of the final edge
            third = up.copy()
            third[r0], third[r1] = r1, r0
            # path root = v0, v1 = up(v0), ... , vk in {r0, r1}: these edges are walked downwards
            path = [root]
            while up[path[-1]] != -1:
                path.append(int(up[path[-1]]))
            vk = path[-1]
            other = r1 if vk == r0 else r0
            parent[other] = vk
            parent[root] = -1
            c0[root] = third[root]
            for prev, v in zip(path[:-1], path[1:]):
                parent[v] = prev
                # adjacency [a, b, third] minus the node we came from
                keep = c1[v] if c0[v] == prev else c0[v]
                c0[v], c1[v] = keep, third[v]
        self.N, self.T, self.M, self.root, self.algo = N, T, M, int(root), algo
        self.parent, self.child0, self.child1 = parent, c0, c1
        self._preorder: Optional[np.ndarray] = None
        self._depth: Optional[np.ndarray] = None

    # ------------------------------------------------------------------ orders
    def preorder(self) -> Tuple[np.ndarray, np.ndarray]:
        """Depth-first preorder (first child first) = node insertion order of the reference's rooted
        DiGraph, and the depth of every node (``cas_tail_preorder``)."""
        if self._preorder is None:
            order = np.empty(self.M, dtype=np.int64)
            depth = np.zeros(self.M, dtype=np.int64)
            c0, c1 = np.ascontiguousarray(self.child0), np.ascontiguousarray(self.child1)
            L = _lib.load()
            count = int(L.cas_tail_preorder(c0.ctypes.data, c1.ctypes.data, self.M, self.root,
                                            order.ctypes.data, depth.ctypes.data))
            if count < 0:
                _lib.check(count, "cas_tail_preorder")
            if count != self.M:
                raise ValueError("join list does not describe one tree")
            self._preorder, self._depth = order, depth
        return self._preorder, self._depth

    def preorder_py(self) -> Tuple[np.ndarray, np.ndarray]:
        """Interpreter version of `preorder` (tests cross-check the two)."""
        c0, c1 = self.child0.tolist(), self.child1.tolist()
        depth = [0] * self.M
        order: List[int] = []
        stack = [self.root]
        push, pop, out = stack.append, stack.pop, order.append
        while stack:
            v = pop()
            out(v)
            x, y = c0[v], c1[v]
            if x >= 0:
                d = depth[v] + 1
                if y >= 0:
                    depth[y] = d
                    push(y)
                depth[x] = d
                push(x)
        return np.asarray(order, dtype=np.int64), np.asarray(depth, dtype=np.int64)

    def heights(self) -> np.ndarray:
        """Height above the deepest leaf (leaves 0); children precede parents in reverse preorder
        (``cas_tail_heights``)."""
        order, _ = self.preorder()
        h = np.zeros(self.M, dtype=np.int64)
        c0, c1 = np.ascontiguousarray(self.child0), np.ascontiguousarray(self.child1)
        L = _lib.load()
        _lib.check(L.cas_tail_heights(order.ctypes.data, order.shape[0], c0.ctypes.data, c1.ctypes.data,
                                      h.ctypes.data), "cas_tail_heights")
        return h

    def heights_py(self) -> np.ndarray:
        order, _ = self.preorder()
        c0, c1 = self.child0.tolist(), self.child1.tolist()
        h = [0] * self.M
        for v in order[::-1].tolist():
            x = c0[v]
            if x >= 0:
                y = c1[v]
                hx = h[x]
                if y >= 0 and h[y] > hx:
                    hx = h[y]
                h[v] = hx + 1
        return np.asarray(h, dtype=np.int64)

    def lca_states(self, leaf_states: np.ndarray, missing: int) -> np.ndarray:
        """[M, m] state rows: leaves as given ([N, m]; the row of an NJ root sample is ignored), internal nodes
        by the Camin-Sokal rule, one vectorised step per height level."""
        S = np.empty((self.M, leaf_states.shape[1]), dtype=leaf_states.dtype)
        S[: self.N] = leaf_states
        h = self.heights()
        internal = np.flatnonzero(self.child0 >= 0)
        internal = internal[np.argsort(h[internal], kind="stable")]
        hs = h[internal]
        bounds = np.flatnonzero(np.diff(hs)) + 1
        for idx in np.split(internal, bounds):
            x, y = self.child0[idx], self.child1[idx]
            two = y >= 0
            S[idx] = S[x]
            if two.all():
                S[idx] = _lca2(S[x], S[y], missing)
            else:
                S[idx[two]] = _lca2(S[x[two]], S[y[two]], missing)
        return S


def successor_lists(order: np.ndarray, child0: np.ndarray, child1: np.ndarray,
                    removed: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """(head, tail, next) linked lists of every node's successors once the `removed` nodes are spliced out
    (``cas_tail_successor_lists``)."""
    M = child0.shape[0]
    head = np.full(M, -1, dtype=np.int64)
    tail = np.full(M, -1, dtype=np.int64)
    nxt = np.full(M, -1, dtype=np.int64)
    order = np.ascontiguousarray(order, dtype=np.int64)
    c0, c1 = np.ascontiguousarray(child0, dtype=np.int64), np.ascontiguousarray(child1, dtype=np.int64)
    rem = np.ascontiguousarray(removed, dtype=np.uint8)
    L = _lib.load()
    _lib.check(L.cas_tail_successor_lists(order.ctypes.data, order.shape[0], c0.ctypes.data, c1.ctypes.data,
                                          rem.ctypes.data, head.ctypes.data, tail.ctypes.data, nxt.ctypes.data),
               "cas_tail_successor_lists")
    return head, tail, nxt


def successor_lists_py(order: np.ndarray, child0: np.ndarray, child1: np.ndarray,
                       removed: np.ndarray) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Interpreter version of `successor_lists` (tests cross-check the two)."""
    M = child0.shape[0]
    head = [-1] * M
    tail = [-1] * M
    nxt = [-1] * M
    rem = removed.tolist()
    c0l, c1l = child0.tolist(), child1.tolist()
    for v in order[::-1].tolist():       # children before parents
        x = c0l[v]
        if x < 0:
            continue
        kids = (x, c1l[v]) if c1l[v] >= 0 else (x,)
        h = t = -1
        for c in kids:                   # kept children, in order
            if not rem[c]:
                if h < 0:
                    h = t = c
                else:
                    nxt[t] = c
                    t = c
        for c in kids:                   # then the spliced children's lists
            if rem[c] and head[c] >= 0:
                if h < 0:
                    h, t = head[c], tail[c]
                else:
                    nxt[t] = head[c]
                    t = tail[c]
        head[v], tail[v] = h, t
    return (np.asarray(head, dtype=np.int64), np.asarray(tail, dtype=np.int64), np.asarray(nxt, dtype=np.int64))


def build_final_tree(names: Sequence[str], internal_names: Sequence[str], joins: np.ndarray, last2: Sequence[int],
                     algo: str, root_sample_name: Optional[str], leaf_states: Optional[np.ndarray], missing: int,
                     collapse_mutationless_edges: bool) -> Tuple[nx.DiGraph, Dict[str, List[int]]]:
    """The rooted tree the reference's tail would leave behind, as one DiGraph with ``length`` attributes, and
    the reconstructed character states of its internal nodes (empty unless collapsing was requested).

    names: the map's samples in row order (including an NJ root sample); internal_names: one name per join;
    leaf_states: [N, m] integer states in the same row order (needed only for collapsing).
    """
    N = len(names)
    root_leaf = None
    if algo != "upgma":
        root_leaf = list(names).index(root_sample_name)
    rt = RootedJoinTree(N, joins, last2, algo, root_leaf)
    order, depth = rt.preorder()
    M, root = rt.M, rt.root
    label = list(names) + list(internal_names)
    if algo == "upgma":
        label.append("root")
    c0, c1 = rt.child0, rt.child1
    internal = c0 >= 0

    removed = np.zeros(M, dtype=bool)
    unifurcation_child = -1
    if algo != "upgma":
        # collapse_unifurcations: only the root sample has a single child; it is spliced out
        unifurcation_child = int(c0[root])
        if c0[unifurcation_child] < 0:
            raise ValueError("degenerate tree (root's neighbour is a leaf)")
        removed[unifurcation_child] = True

    states = None
    if collapse_mutationless_edges:
        if leaf_states is None:
            raise ValueError("collapsing mutationless edges needs the leaves' character states")
        states = rt.lca_states(np.asarray(leaf_states), missing)
        if unifurcation_child >= 0:
            states[root] = states[unifurcation_child]  # LCA of the spliced child's children
        # parent in the tree after collapse_unifurcations
        par = rt.parent.copy()
        if unifurcation_child >= 0:
            par[par == unifurcation_child] = root
        cand = internal & ~removed
        cand[root] = False
        idx = np.flatnonzero(cand)
        same = (states[idx] == states[par[idx]]).all(axis=1)
        removed[idx[same]] = True

    # successor lists of kept nodes as linked lists: own kept children first, then the lists of removed
    # children, in child order (reference: add_edge appends, post-order splicing)
    head_a, tail, nxt_a = successor_lists(order, c0, c1, removed)
    kept_a = np.ascontiguousarray(order[~removed[order]])
    g = nx.DiGraph()
    node, succ, pred = getattr(g, "_node", None), getattr(g, "_succ", None), getattr(g, "_pred", None)
    direct = node is not None and succ is not None and pred is not None and not node
    if direct and _graphfill is not None:
        # the same dictionaries as the interpreter loop below, filled by the C extension (csrc/graphfill.c)
        _graphfill.fill(node, succ, pred, label, kept_a, head_a, nxt_a, np.ascontiguousarray(depth))
        if all(isinstance(x, str) for x in names):
            g.graph[PREPARED] = True
        return g, _internal_states(states, internal, removed, label)
    head, nxt = head_a.tolist(), nxt_a.tolist()
    kept_order = kept_a.tolist()
    dl = depth.tolist()
    if direct:
        # networkx keeps {node: attrs}, {u: {v: attrs}}, {v: {u: attrs}} with the edge attribute dict shared
        # between the successor and the predecessor map; filling them directly skips the per-call checks of
        # add_nodes_from / add_edges_from (half of this function's time at 10^5 nodes)
        # node attributes as populate_tree leaves them (CassiopeiaTree.py:174-203): an empty state list and
        # the time = sum of the branch lengths from the root = depth in the uncollapsed tree.  The tree
        # object's populate_tree recomputes the same values; ours skips its per-node passes when it finds
        # the marker below.
        for v in kept_order:
            lv = label[v]
            node[lv] = {"character_states": [], "time": dl[v]}
            succ[lv] = {}
            pred[lv] = {}
        if all(isinstance(x, str) for x in names):
            g.graph[PREPARED] = True
        for v in kept_order:
            c = head[v]
            if c < 0:
                continue
            lv, dv = label[v], dl[v]
            sv = succ[lv]
            while c >= 0:
                lc = label[c]
                attrs = {"length": dl[c] - dv}
                sv[lc] = attrs
                pred[lc][lv] = attrs
                c = nxt[c]
    else:  # pragma: no cover - a networkx without these internals
        g.add_nodes_from(label[v] for v in kept_order)
        edges = []
        for v in kept_order:
            c = head[v]
            lv, dv = label[v], dl[v]
            while c >= 0:
                edges.append((lv, label[c], {"length": dl[c] - dv}))
                c = nxt[c]
        g.add_edges_from(edges)

    return g, _internal_states(states, internal, removed, label)


def _internal_states(states, internal, removed, label) -> Dict[str, List[int]]:
    if states is None:
        return {}
    keep_int = np.flatnonzero(internal & ~removed)
    rows = states[keep_int].tolist()
    return {label[v]: row for v, row in zip(keep_int.tolist(), rows)}
