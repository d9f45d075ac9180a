"""Data containers of the path (reference: cassiopeia/data)."""
from .CassiopeiaTree import CassiopeiaTree
from .utilities import compute_dissimilarity_map

__all__ = ["CassiopeiaTree", "compute_dissimilarity_map"]
