"""Shared-Mutation-Joining on the B200 behind the reference's class API
(cassiopeia/solver/SharedMutationJoiningSolver.py:29-357)."""
from __future__ import annotations

from typing import Callable, Optional

import networkx as nx
import numpy as np

from .. import engine, host_prep
from ..data import utilities as data_utilities
from ..mixins import SharedMutationJoiningSolverError
from . import dissimilarity_functions, solver_utilities


class SharedMutationJoiningSolver:
    """Bottom-up clustering that joins, at every step, the two samples sharing the most mutations.

    Constructor keywords are the reference's (:60-74).  ``solve`` follows :99-209: similarity map of the
    character matrix rows *as given* (no de-duplication), `while N > 1` join the first row-major arg-max, give
    the new node the Camin-Sokal LCA row and recompute its similarities, ``populate_tree``, optional
    ``collapse_mutationless_edges``.  The map and the whole loop run on the GPU (``cas_pairwise_map`` +
    ``cas_smj_solve``, float64, bit-identical to the reference's arithmetic); only similarity functions the
    CUDA kernels implement are accepted (no CPU fallback): ``hamming_similarity_without_missing`` (default),
    ``hamming_similarity_normalized_over_missing``, ``weighted_hamming_similarity`` and, formally,
    ``weighted_hamming_distance`` / ``exponential_negative_hamming_distance``.
    """

    def __init__(self, similarity_function: Optional[Callable] = dissimilarity_functions.hamming_similarity_without_missing,
                 prior_transformation: str = "negative_log"):
        self.prior_transformation = prior_transformation
        self.similarity_function = similarity_function
        name = None if similarity_function is None else dissimilarity_functions.gpu_kernel_name(similarity_function)
        if name is None:
            raise SharedMutationJoiningSolverError(
                "the B200 SharedMutationJoiningSolver needs one of the similarity functions of "
                "`dissimilarity_functions` that are built as CUDA kernels (there is no CPU fallback)")
        self._function_name = name
        self.last_join_list = None  # (joins incl. the root join, names) of the most recent solve

    def solve(self, cassiopeia_tree, layer: Optional[str] = None, collapse_mutationless_edges: bool = False,
              logfile: str = "stdout.log") -> None:
        character_matrix = (cassiopeia_tree.layers[layer] if layer else cassiopeia_tree.character_matrix).copy()
        weights = None
        if cassiopeia_tree.priors:
            weights = solver_utilities.transform_priors(cassiopeia_tree.priors, self.prior_transformation)
        names = list(character_matrix.index)
        n = len(names)
        cm = host_prep.as_int_matrix(character_matrix)
        lookup = "match" if self._function_name in data_utilities._MATCH_LOOKUP else "mismatch"
        codes, missing_code, table, n_states = host_prep.compact_states(
            cm, cassiopeia_tree.missing_state_indicator, weights, lookup=lookup)
        if n >= 2:
            joins, last2 = engine.smj_solve(codes, missing_code, table, n_states, self._function_name)
            joins = np.vstack([joins.reshape(-1, 2), np.asarray(last2, dtype=np.int32).reshape(1, 2)])
        else:
            joins = np.zeros((0, 2), dtype=np.int32)
        self.last_join_list = (joins, names)

        # the reference's tree building (:166-185): a DiGraph with an edge from every new node to its two children
        node_name_generator = solver_utilities.node_name_generator()
        tree = nx.DiGraph()
        tree.add_nodes_from(names)
        label = list(names)
        for a, b in joins.tolist():
            new_node_name = next(node_name_generator)
            label.append(new_node_name)
            tree.add_node(new_node_name)
            tree.add_edges_from([(new_node_name, label[a]), (new_node_name, label[b])])

        cassiopeia_tree.populate_tree(tree, layer=layer)
        if collapse_mutationless_edges:
            cassiopeia_tree.collapse_mutationless_edges(infer_ancestral_characters=True)
