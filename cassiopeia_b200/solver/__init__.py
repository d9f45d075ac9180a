"""Distance solvers of the path (reference: cassiopeia/solver)."""
from . import dissimilarity_functions as dissimilarity
from . import dissimilarity_functions, solver_utilities
from .DistanceSolver import DistanceSolver
from .NeighborJoiningSolver import NeighborJoiningSolver
from .SharedMutationJoiningSolver import SharedMutationJoiningSolver
from .UPGMASolver import UPGMASolver

__all__ = ["DistanceSolver", "NeighborJoiningSolver", "UPGMASolver", "SharedMutationJoiningSolver", "dissimilarity",
           "dissimilarity_functions", "solver_utilities"]
