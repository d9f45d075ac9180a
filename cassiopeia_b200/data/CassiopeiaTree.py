"""Minimal tree container for the distance-solver path.

Mirrors the subset of the reference's ``CassiopeiaTree``
(cassiopeia/data/CassiopeiaTree.py:39) that ``DistanceSolver.solve`` touches: the
character matrix / layers / priors / dissimilarity-map accessors (:229-260, :1777-1958),
``populate_tree`` (:141-203), ``collapse_unifurcations`` (:1669-1715),
``reconstruct_ancestral_characters`` (:648-675) and ``collapse_mutationless_edges``
(:1717-1775).  The B200 solvers duck-type on this interface, so they accept either
this class or the reference's own object.  ``compute_dissimilarity_map`` runs on the
GPU; nothing here falls back to a CPU map.
"""
from __future__ import annotations

import collections
import warnings
from typing import Callable, Dict, Iterator, List, Optional, Tuple

import networkx as nx
import numpy as np
import pandas as pd

from ..mixins import CassiopeiaTreeError, CassiopeiaTreeWarning


class CassiopeiaTree:
    def __init__(self, character_matrix: Optional[pd.DataFrame] = None, missing_state_indicator: int = -1,
                 cell_meta: Optional[pd.DataFrame] = None, character_meta: Optional[pd.DataFrame] = None,
                 priors: Optional[Dict[int, Dict[int, float]]] = None, tree: Optional[nx.DiGraph] = None,
                 dissimilarity_map: Optional[pd.DataFrame] = None, parameters: Optional[dict] = None,
                 root_sample_name: Optional[str] = None) -> None:
        self.missing_state_indicator = missing_state_indicator
        self.cell_meta = cell_meta
        self.character_meta = character_meta
        self.priors = priors
        self.parameters = dict(parameters) if parameters else {}
        self._network: Optional[nx.DiGraph] = None
        self._layers: Dict[str, pd.DataFrame] = {}
        self._character_matrix: Optional[pd.DataFrame] = None
        if character_matrix is not None:
            self.character_matrix = character_matrix
        if tree is not None:
            self.populate_tree(tree.copy())
        self._dissimilarity_map: Optional[pd.DataFrame] = None
        if dissimilarity_map is not None:
            self.set_dissimilarity_map(dissimilarity_map)
        self.root_sample_name = root_sample_name

    # ------------------------------------------------------------------ data
    @property
    def layers(self) -> Dict[str, pd.DataFrame]:
        return self._layers

    @property
    def character_matrix(self) -> Optional[pd.DataFrame]:
        return self._character_matrix

    @character_matrix.setter
    def character_matrix(self, character_matrix: pd.DataFrame) -> None:
        index = character_matrix.index
        if index.inferred_type != "string" and not all(type(i) == str for i in index):
            raise CassiopeiaTreeError("Index of character matrix must consist of strings.")
        self._character_matrix = character_matrix.copy()

    @property
    def n_cell(self) -> int:
        if self._character_matrix is None:
            if self._network is None:
                raise CassiopeiaTreeError("This is an empty object with no tree or character matrix.")
            return len(self.leaves)
        return self._character_matrix.shape[0]

    @property
    def n_character(self) -> int:
        if self._character_matrix is None:
            raise CassiopeiaTreeError("No character matrix is detected in this tree.")
        return self._character_matrix.shape[1]

    # ------------------------------------------------------- dissimilarity map
    def get_dissimilarity_map(self) -> Optional[pd.DataFrame]:
        return None if self._dissimilarity_map is None else self._dissimilarity_map.copy()

    def set_dissimilarity_map(self, dissimilarity_map: pd.DataFrame) -> None:
        cm = self._character_matrix
        if cm is not None:
            names = collections.Counter(cm.index)
            if (cm.shape[0] != dissimilarity_map.shape[0]
                    or names != collections.Counter(dissimilarity_map.index)
                    or names != collections.Counter(dissimilarity_map.columns)):
                warnings.warn("The samples in the existing character matrix and specified "
                              "dissimilarity map do not agree.", CassiopeiaTreeWarning)
        self._dissimilarity_map = dissimilarity_map.copy()

    def get_dissimilarity(self, node: str) -> Dict[str, float]:
        if self._dissimilarity_map is None:
            raise CassiopeiaTreeError("No dissimilarity map is defined.")
        if node not in self._dissimilarity_map.index:
            raise CassiopeiaTreeError(f"`{node}` is not in the dissimilarity map.")
        return self._dissimilarity_map.loc[node].to_dict()

    def set_dissimilarity(self, node: str, dissimilarity: Dict[str, float]) -> None:
        """Adds / overwrites one sample's row and column (reference :1829-1863)."""
        if self._dissimilarity_map is None:
            raise CassiopeiaTreeError("This function may only be used with an existing dissimilarity map.")
        dm = self._dissimilarity_map
        if set(dissimilarity.keys()) - {node} != set(dm.index) - {node}:
            raise CassiopeiaTreeError("Provided dissimilarity dictionary does not provide "
                                      "dissimilarities to all other leaves.")
        names = list(dm.index)
        arr = dm.to_numpy(dtype=np.float64, copy=True) if len(names) else np.zeros((0, 0))
        if node not in dm.index:
            arr = np.pad(arr, ((0, 1), (0, 1)))
            names.append(node)
        pos = {nm: i for i, nm in enumerate(names)}
        p = pos[node]
        arr[p, :] = 0
        arr[:, p] = 0
        for other, d in dissimilarity.items():
            arr[p, pos[other]] = d
            arr[pos[other], p] = d
        self._dissimilarity_map = pd.DataFrame(arr, index=names, columns=names)

    def compute_dissimilarity_map(self, dissimilarity_function: Optional[Callable] = None,
                                  prior_transformation: str = "negative_log", layer: Optional[str] = None,
                                  threads: int = 1) -> None:
        """Same contract as reference :1865-1958, computed by the CUDA map kernel.

        `threads` is accepted for signature parity and ignored (the reference uses it to
        fan the map out over processes)."""
        from . import utilities

        character_matrix = self.layers[layer] if layer is not None else self.character_matrix
        if character_matrix is None:
            raise CassiopeiaTreeError("No character matrix is detected in this tree.")
        names, D = utilities.dissimilarity_map_frame_inputs(
            character_matrix, self.priors, self.missing_state_indicator, dissimilarity_function,
            prior_transformation)
        self.set_dissimilarity_map(pd.DataFrame(D, index=names, columns=names))

    # ------------------------------------------------------------- topology
    def _check_tree(self) -> None:
        if self._network is None:
            raise CassiopeiaTreeError("Tree has not been initialized.")

    def populate_tree(self, tree: nx.DiGraph, layer: Optional[str] = None) -> None:
        if not isinstance(tree, nx.DiGraph):
            raise CassiopeiaTreeError("Please pass a Networkx DiGraph object.")
        if tree.graph.pop("cas_b200_prepared", False):
            # built by tree_tail.build_final_tree: string names, `character_states` = [] and `time` on every
            # node, `length` on every edge -- exactly what the passes below would leave behind
            self._network = tree
            character_matrix = self.layers[layer] if layer else self.character_matrix
            prepared = self.__dict__.pop("_prepared_leaf_rows", None)
            if character_matrix is not None:
                if prepared is not None and len(prepared[0]) == character_matrix.shape[0]:
                    self._set_leaf_rows(*prepared)
                else:
                    self.set_character_states_at_leaves(layer=layer)
            return
        if any(not isinstance(n, str) for n in tree.nodes):
            tree = nx.relabel_nodes(tree, {n: str(n) for n in tree.nodes})
        self._network = tree
        # networkx's node / adjacency dictionaries are used directly: the view objects cost a Python
        # call per access, which dominates on trees with 10^5 nodes
        node_attrs = tree._node
        for attrs in node_attrs.values():
            attrs["character_states"] = []
        character_matrix = self.layers[layer] if layer else self.character_matrix
        if character_matrix is not None:
            self.set_character_states_at_leaves(layer=layer)
        succ = tree._succ
        for nbrs in succ.values():
            for attrs in nbrs.values():
                if "length" not in attrs:
                    attrs["length"] = 1
        root = self.root
        node_attrs[root]["time"] = 0
        stack = [root]
        while stack:
            u = stack.pop()
            t = node_attrs[u]["time"]
            for v, attrs in succ[u].items():
                node_attrs[v]["time"] = t + attrs["length"]
                stack.append(v)

    def set_character_states_at_leaves(self, character_matrix=None, layer: Optional[str] = None) -> None:
        self._check_tree()
        if character_matrix is not None and layer is not None:
            raise CassiopeiaTreeError("You may only specify one of the following: character_matrix or layer.")
        if character_matrix is None:
            character_matrix = self.layers[layer] if layer else self.character_matrix
        elif isinstance(character_matrix, dict):
            character_matrix = pd.DataFrame.from_dict(character_matrix, orient="index")
        leaves = self.leaves
        index = character_matrix.index.tolist()  # (iterating a pandas string index directly is 30x slower)
        if set(leaves) != set(index):
            raise CassiopeiaTreeError("Character matrix index does not match set of leaves.")
        rows = character_matrix.to_numpy().tolist()
        self._set_leaf_rows(index, rows, leaves)

    def accept_prepared_leaf_rows(self, index: List[str], rows: List[List[int]]) -> None:
        """`index` / `rows` = ``character_matrix.index.tolist()`` / ``character_matrix.to_numpy().tolist()`` of the
        matrix (or layer) the NEXT ``populate_tree`` call of a prepared graph will read: the B200 solvers compute
        them while the GPU loop runs."""
        self._prepared_leaf_rows = (index, rows)

    def _set_leaf_rows(self, index, rows, leaves=None) -> None:
        leaves = self.leaves if leaves is None else leaves
        if len(leaves) != len(index) or set(leaves) != set(index):
            raise CassiopeiaTreeError("Character matrix index does not match set of leaves.")
        by_name = dict(zip(index, rows))
        node_attrs = self._network._node
        for n in leaves:
            node_attrs[n]["character_states"] = by_name[n]

    @property
    def root(self) -> str:
        self._check_tree()
        pred = self._network._pred
        return next(n for n in self._network._node if not pred[n])

    @property
    def leaves(self) -> List[str]:
        self._check_tree()
        succ = self._network._succ
        return [n for n in self._network._node if not succ[n]]

    @property
    def internal_nodes(self) -> List[str]:
        self._check_tree()
        succ = self._network._succ
        return [n for n in self._network._node if succ[n]]

    @property
    def nodes(self) -> List[str]:
        self._check_tree()
        return list(self._network.nodes)

    @property
    def edges(self) -> List[Tuple[str, str]]:
        self._check_tree()
        return list(self._network.edges)

    def is_leaf(self, node: str) -> bool:
        self._check_tree()
        return self._network.out_degree(node) == 0

    def is_root(self, node: str) -> bool:
        self._check_tree()
        return node == self.root

    def is_internal_node(self, node: str) -> bool:
        self._check_tree()
        return self._network.out_degree(node) > 0

    def parent(self, node: str) -> str:
        self._check_tree()
        return next(iter(self._network.predecessors(node)))

    def children(self, node: str) -> List[str]:
        self._check_tree()
        return list(self._network.successors(node))

    def get_character_states(self, node: str) -> List[int]:
        self._check_tree()
        return self._network.nodes[node]["character_states"][:]

    def set_character_states(self, node: str, states: List[int]) -> None:
        self._check_tree()
        self._network.nodes[node]["character_states"] = list(states)

    def set_internal_character_states(self, states_by_node: Dict[str, List[int]]) -> None:
        """Bulk form of ``set_character_states`` for internal nodes (used by the array-based solver tail)."""
        self._check_tree()
        nodes = self._network.nodes
        for node, states in states_by_node.items():
            nodes[node]["character_states"] = states

    def get_branch_length(self, parent: str, child: str) -> float:
        self._check_tree()
        return self._network[parent][child]["length"]

    def get_time(self, node: str) -> float:
        self._check_tree()
        return self._network.nodes[node]["time"]

    def get_times(self) -> Dict[str, float]:
        self._check_tree()
        return {n: self._network.nodes[n]["time"] for n in self._network.nodes}

    def get_tree_topology(self) -> nx.DiGraph:
        self._check_tree()
        g = nx.DiGraph()
        g.add_nodes_from(self._network.nodes)
        g.add_edges_from(self._network.edges)
        return g

    def depth_first_traverse_nodes(self, source: Optional[str] = None, postorder: bool = True) -> Iterator[str]:
        self._check_tree()
        if source is None:
            source = self.root
        if postorder:
            return nx.dfs_postorder_nodes(self._network, source=source)
        return nx.dfs_preorder_nodes(self._network, source=source)

    def depth_first_traverse_edges(self, source: Optional[str] = None) -> Iterator[Tuple[str, str]]:
        self._check_tree()
        if source is None:
            source = self.root
        return nx.dfs_edges(self._network, source=source)

    def get_newick(self, record_branch_lengths: bool = False, record_node_names: bool = False) -> str:
        """Newick string in the reference's format (cassiopeia/data/utilities.py:139-186),
        built iteratively so caterpillar trees of any depth work."""
        self._check_tree()
        g = self._network
        root = self.root
        text: Dict[str, str] = {}
        for node in nx.dfs_postorder_nodes(g, source=root):
            weight = ""
            if record_branch_lengths and g.in_degree(node) > 0:
                weight = ":" + str(g[next(iter(g.predecessors(node)))][node]["length"])
            kids = list(g.successors(node))
            if not kids:
                text[node] = str(node) + weight
            else:
                inner = ",".join(text.pop(c) for c in kids)
                text[node] = "(" + inner + ")" + (str(node) if record_node_names else "") + weight
        return text[root] + ";"

    # ------------------------------------------------------------ collapsing
    def _splice(self, parent: str, node: str) -> None:
        """Removes `node`, attaching its children to `parent` with summed branch lengths."""
        g = self._network
        t = g[parent][node]["length"]
        for child in list(g.successors(node)):
            g.add_edge(parent, child, length=t + g[node][child]["length"])
        g.remove_node(node)

    def collapse_unifurcations(self, source: Optional[str] = None) -> None:
        self._check_tree()
        g = self._network
        if source is None:
            source = self.root
        for node in list(nx.dfs_postorder_nodes(g, source=source)):
            if g.out_degree(node) != 1:
                continue
            if node == source:
                self._splice(node, next(iter(g.successors(node))))
            else:
                parent = next(iter(g.predecessors(node)))
                child = next(iter(g.successors(node)))
                g.add_edge(parent, child, length=g[parent][node]["length"] + g[node][child]["length"])
                g.remove_node(node)

    def reconstruct_ancestral_characters(self) -> None:
        """Camin-Sokal LCA states bottom-up (reference :648-675 with
        cassiopeia/data/utilities.py:29-97 restricted to unambiguous integer states):
        all children missing -> missing; the present children agree -> that state; else 0."""
        self._check_tree()
        g = self._network
        miss = self.missing_state_indicator
        for n in list(nx.dfs_postorder_nodes(g, source=self.root)):
            kids = list(g.successors(n))
            if not kids:
                if g.nodes[n]["character_states"] == []:
                    raise CassiopeiaTreeError("Character states have not been initialized at leaves.")
                continue
            vecs = np.asarray([g.nodes[c]["character_states"] for c in kids])
            if vecs.dtype == object:
                raise CassiopeiaTreeError("ambiguous character states are not supported on this path")
            pres = vecs != miss
            hi = np.where(pres, vecs, np.iinfo(vecs.dtype).min).max(axis=0)
            lo = np.where(pres, vecs, np.iinfo(vecs.dtype).max).min(axis=0)
            out = np.where(pres.any(axis=0), np.where(hi == lo, hi, 0), miss)
            g.nodes[n]["character_states"] = out.tolist()

    def collapse_mutationless_edges(self, infer_ancestral_characters: bool) -> None:
        self._check_tree()
        g = self._network
        if infer_ancestral_characters:
            self.reconstruct_ancestral_characters()
        for n in list(nx.dfs_postorder_nodes(g, source=self.root)):
            if g.out_degree(n) == 0:
                if g.nodes[n]["character_states"] == []:
                    raise CassiopeiaTreeError("Character states have not been initialized at leaves.")
                continue
            if g.nodes[n]["character_states"] == []:
                raise CassiopeiaTreeError("Character states empty at internal node.")
            for child in list(g.successors(n)):
                if g.out_degree(child) > 0 and g.nodes[n]["character_states"] == g.nodes[child]["character_states"]:
                    self._splice(n, child)
