"""B200-native drop-in for Cassiopeia's distance-solver hot path.

``solver.NeighborJoiningSolver`` / ``solver.UPGMASolver`` keep the reference's
constructor keywords and ``solve(CassiopeiaTree, layer, collapse_mutationless_edges)``
contract; the dissimilarity map and the agglomeration loop run in hand-written sm_100a
CUDA kernels behind the C ABI of ``include/cas_b200.h``.
"""
from . import data, mixins, solver
from .data import CassiopeiaTree
from .solver import NeighborJoiningSolver, SharedMutationJoiningSolver, UPGMASolver

__version__ = "0.1.0"
__all__ = ["data", "mixins", "solver", "CassiopeiaTree", "NeighborJoiningSolver", "UPGMASolver",
           "SharedMutationJoiningSolver"]
