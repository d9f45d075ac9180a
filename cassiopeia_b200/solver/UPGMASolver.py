"""UPGMA on the B200 behind the reference's class API (cassiopeia/solver/UPGMASolver.py:20-250)."""
from __future__ import annotations

from collections import defaultdict
from typing import Callable, List, Optional, Tuple

import networkx as nx

from ..mixins import DistanceSolverError
from . import dissimilarity_functions
from .DistanceSolver import B200_IMPLEMENTATIONS, DistanceSolver


class UPGMASolver(DistanceSolver):
    """UPGMA: join the closest pair, size-weighted average linkage, root above the last two.

    Constructor keywords are the reference's (:56-87); see ``NeighborJoiningSolver`` for the
    meaning of ``fast`` / ``implementation``.  UPGMA values involve no reductions, so the
    float64 paths ("b200_exact" = "b200_strict" here) reproduce the reference bit for bit at any size.
    """

    _algo = "upgma"

    def __init__(self, dissimilarity_function: Optional[Callable] = dissimilarity_functions.weighted_hamming_distance,
                 prior_transformation: str = "negative_log", fast: bool = False,
                 implementation: str = "b200_fast", threads: int = 1):
        if fast:
            if implementation in B200_IMPLEMENTATIONS:
                self._implementation = implementation
            elif implementation == "ccphylo_upgma":
                raise DistanceSolverError(
                    "'ccphylo_upgma' needs the external CCPhylo binary, which this build does not wrap. "
                    "Options are: 'b200_fast', 'b200_exact', 'b200_strict'")
            else:
                raise DistanceSolverError(
                    "Invalid fast implementation of UPGMA. Options are: 'b200_fast', 'b200_exact', 'b200_strict'")
        else:
            self._implementation = "b200_exact"
        super().__init__(dissimilarity_function=dissimilarity_function, add_root=True,
                         prior_transformation=prior_transformation, threads=threads)
        self._cluster_to_cluster_size = defaultdict(int)

    def root_tree(self, tree: nx.Graph, root_sample: str, remaining_samples: List[str]) -> nx.DiGraph:
        """New "root" node above the last two clusters (:89-116)."""
        tree.add_node("root")
        tree.add_edges_from([("root", remaining_samples[0]), ("root", remaining_samples[1])])
        rooted_tree = nx.DiGraph()
        for u, v in nx.dfs_edges(tree, source="root"):
            rooted_tree.add_edge(u, v)
        return rooted_tree

    def _cluster_sizes(self, cherry, new_node) -> Tuple[int, int]:
        """Sizes for the single-step hook; leaves count 1 (:165-170)."""
        i_size = max(1, self._cluster_to_cluster_size[cherry[0]])
        j_size = max(1, self._cluster_to_cluster_size[cherry[1]])
        self._cluster_to_cluster_size[new_node] = i_size + j_size
        return i_size, j_size

    def setup_root_finder(self, cassiopeia_tree) -> None:
        """The root is created by ``root_tree``; only its name is fixed here (:240-250)."""
        cassiopeia_tree.root_sample_name = "root"


UPGMASolver._stock_root_tree = UPGMASolver.root_tree
