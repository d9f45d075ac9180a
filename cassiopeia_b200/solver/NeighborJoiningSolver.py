"""Neighbor-Joining on the B200 behind the reference's class API
(cassiopeia/solver/NeighborJoiningSolver.py:23-309)."""
from __future__ import annotations

from typing import Callable, List, Optional

import networkx as nx
import numpy as np
import pandas as pd

from .. import engine
from ..data import utilities as data_utilities
from ..mixins import DistanceSolverError
from . import dissimilarity_functions, solver_utilities
from .DistanceSolver import B200_IMPLEMENTATIONS, DistanceSolver


class NeighborJoiningSolver(DistanceSolver):
    """Saitou-Nei Neighbor-Joining.

    Constructor keywords are the reference's (:68-98).  ``fast=False`` (default) produces the reference's
    join sequence (``"b200_exact"``: float64 map, no FMA, row sums maintained incrementally under rigorous
    error bounds, the reference's sequential ``d.sum(axis=1)`` evaluated only for the rows of pairs the
    maintained sums cannot separate -- DESIGN 4.3).  ``fast=True`` selects ``implementation``:
    ``"b200_fast"`` (float32 map, maintained sums: the reference's tree whenever no near-tie occurs),
    ``"b200_exact"`` or ``"b200_strict"`` (every row sum re-accumulated sequentially every iteration, the
    reference's arithmetic literally; same joins as ``"b200_exact"``, kept for verification).  The reference's
    CCPhylo implementations are an external binary and are not available here.
    """

    _algo = "nj"

    def __init__(self, dissimilarity_function: Optional[Callable] = dissimilarity_functions.weighted_hamming_distance,
                 add_root: bool = False, prior_transformation: str = "negative_log", fast: bool = False,
                 implementation: str = "b200_fast", threads: int = 1):
        if fast:
            if implementation in B200_IMPLEMENTATIONS:
                self._implementation = implementation
            elif implementation in ("ccphylo_dnj", "ccphylo_hnj", "ccphylo_nj"):
                raise DistanceSolverError(
                    f"'{implementation}' needs the external CCPhylo binary, which this build does not "
                    "wrap. Options are: 'b200_fast', 'b200_exact', 'b200_strict'")
            else:
                raise DistanceSolverError(
                    "Invalid fast implementation of Neighbor-Joining. Options are: 'b200_fast', 'b200_exact', 'b200_strict'")
        else:
            self._implementation = "b200_exact"
        super().__init__(dissimilarity_function=dissimilarity_function, add_root=add_root,
                         prior_transformation=prior_transformation, threads=threads)

    def root_tree(self, tree: nx.Graph, root_sample: str, remaining_samples: List[str]) -> nx.DiGraph:
        """Joins the last two nodes and orients every edge away from `root_sample` (:100-121)."""
        tree.add_edge(remaining_samples[0], remaining_samples[1])
        rooted_tree = nx.DiGraph()
        for u, v in nx.dfs_edges(tree, source=root_sample):
            rooted_tree.add_edge(u, v)
        return rooted_tree

    @staticmethod
    def compute_q(dissimilarity_map: np.ndarray) -> np.ndarray:
        """Q-criterion matrix, Q(i,j) = d(i,j) - 1/(n-2) (sum d(i,:) + sum d(j,:)), zero diagonal
        (:142-171), evaluated by the CUDA kernels in the reference's operation order."""
        return engine.nj_q(np.asarray(dissimilarity_map, dtype=np.float64))

    def setup_root_finder(self, cassiopeia_tree) -> None:
        """Implicit all-zero "root" sample (:262-309).  Without a stored map the map of the
        rooted character matrix is computed on the GPU (the layer is ignored here, as in
        the reference); with one, only the root row is added."""
        character_matrix = cassiopeia_tree.character_matrix.copy()
        rooted_character_matrix = character_matrix.copy()
        rooted_character_matrix.loc["root"] = [0] * rooted_character_matrix.shape[1]
        cassiopeia_tree.root_sample_name = "root"
        cassiopeia_tree.character_matrix = rooted_character_matrix

        if self.dissimilarity_function is None:
            raise DistanceSolverError(
                "Please specify a dissimilarity function to add an implicit root, or specify an explicit root")

        try:
            if cassiopeia_tree.get_dissimilarity_map() is None:
                self._pending_device_map = self._compute_device_map(cassiopeia_tree, rooted_character_matrix)
            elif dissimilarity_functions.is_gpu_kernel(self.dissimilarity_function):
                # one row block of the CUDA map kernel instead of n Python calls
                dissimilarity = data_utilities.row_to_all_distances(
                    rooted_character_matrix, "root", cassiopeia_tree.priors,
                    cassiopeia_tree.missing_state_indicator, self.prior_transformation,
                    self.dissimilarity_function)
                dissimilarity["root"] = 0
                cassiopeia_tree.set_dissimilarity("root", dissimilarity)
            else:
                # a user-supplied Python function can only be evaluated on the host, per leaf, exactly as
                # the reference does; the map itself was supplied by the user
                weights = None
                if cassiopeia_tree.priors:
                    weights = solver_utilities.transform_priors(cassiopeia_tree.priors, self.prior_transformation)
                root_states = rooted_character_matrix.loc["root"].values
                dissimilarity = {"root": 0}
                for leaf in character_matrix.index:
                    dissimilarity[leaf] = self.dissimilarity_function(
                        root_states, rooted_character_matrix.loc[leaf].values,
                        cassiopeia_tree.missing_state_indicator, weights)
                cassiopeia_tree.set_dissimilarity("root", dissimilarity)
        finally:
            cassiopeia_tree.character_matrix = character_matrix


NeighborJoiningSolver._stock_root_tree = NeighborJoiningSolver.root_tree
