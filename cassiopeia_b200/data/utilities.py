"""GPU counterparts of cassiopeia/data/utilities.py on the hot path."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import numpy as np

from .. import engine, host_prep
from ..mixins import DistanceSolverError
from ..solver import dissimilarity_functions, solver_utilities


_MATCH_LOOKUP = ("hamming_similarity_without_missing", "hamming_similarity_normalized_over_missing",
                 "weighted_hamming_similarity")


def _require_gpu_function(dissimilarity_function: Optional[Callable]) -> str:
    """Name of the CUDA pair function implementing `dissimilarity_function`; raises otherwise."""
    name = None if dissimilarity_function is None else dissimilarity_functions.gpu_kernel_name(dissimilarity_function)
    if name is None:
        raise DistanceSolverError(
            "only the pair functions of `dissimilarity_functions` (weighted_hamming_distance, the hamming "
            "similarities, exponential_negative_hamming_distance) are built as CUDA kernels; for any other "
            "function populate the tree with a precomputed dissimilarity map (there is no CPU fallback)"
        )
    return name


def device_map(cm: np.ndarray, missing: int, weights: Optional[Dict[int, Dict[int, float]]],
               dtype: str = "f64", method: str = "exact", function: str = "weighted_hamming_distance",
               count_rows: Optional[int] = None):
    """[n, ld] CUDA tensor holding the full symmetric map of the int matrix `cm` (rows as given)."""
    lookup = "match" if function in _MATCH_LOOKUP else "mismatch"
    codes, missing_code, table, n_states = host_prep.compact_states(cm, missing, weights, lookup=lookup,
                                                                    count_rows=count_rows)
    if function != "weighted_hamming_distance":
        method = "exact"
    return engine.whd_map(codes, missing_code, table, n_states, dtype=dtype, method=method, function=function)


def compute_dissimilarity_map(cm: np.ndarray, C: int, dissimilarity_function: Callable,
                              weights: Optional[Dict[int, Dict[int, float]]] = None,
                              missing_state_indicator: int = -1, threads: int = 1) -> np.ndarray:
    """Condensed float64 vector of the C*(C-1)/2 pairwise dissimilarities, row-major pair
    order -- the contract of cassiopeia/data/utilities.py:189-295 (`threads` is ignored)."""
    function = _require_gpu_function(dissimilarity_function)
    cm = host_prep.as_int_matrix(cm)[:C]
    if C < 2:
        return np.zeros(0, dtype=np.float64)
    D = device_map(cm, missing_state_indicator, weights, dtype="f64", method="exact", function=function)
    full = D[:, :C].cpu().numpy()
    iu = np.triu_indices(C, k=1)
    return np.ascontiguousarray(full[iu])


def dissimilarity_map_frame_inputs(character_matrix, priors, missing_state_indicator: int,
                                   dissimilarity_function: Optional[Callable],
                                   prior_transformation: str = "negative_log",
                                   dtype: str = "f64", method: str = "exact",
                                   to_host: bool = True) -> Tuple[List[str], object]:
    """(row names in the reference's map order, full n x n map).

    Restates cassiopeia/data/CassiopeiaTree.py:1914-1958: priors -> weights, the
    de-duplication row order, the square map.  `to_host=False` keeps the map on the
    device as an [n, ld] CUDA tensor (used by the solvers for large n)."""
    function = _require_gpu_function(dissimilarity_function)
    weights = None
    if priors:
        weights = solver_utilities.transform_priors(priors, prior_transformation)
    cm = host_prep.as_int_matrix(character_matrix)
    order, first = host_prep.reference_row_order(cm, return_first=True)
    names = np.asarray(character_matrix.index, dtype=object)[order].tolist()  # one gather, not n __getitem__ calls
    n_unique, rep = host_prep.duplicate_groups(cm, order, first=first if cm.shape[1] else None)
    D = device_map(cm[order], missing_state_indicator, weights, dtype=dtype, method=method, function=function,
                   count_rows=n_unique)
    if n_unique < cm.shape[0] and function not in ("weighted_hamming_distance", "hamming_distance"):
        # The reference evaluates the function on the distinct rows only and copies a distinct row -- zero
        # diagonal included -- into each of its duplicates (CassiopeiaTree.py:1937-1951): cells with identical
        # rows are at 0 from each other whatever the function (a similarity of a row with itself is not 0).
        _zero_duplicate_groups(D, rep, n_unique)
    if to_host:
        n = cm.shape[0]
        return names, D[:, :n].cpu().numpy()
    return names, D


def _zero_duplicate_groups(D, rep: np.ndarray, n_unique: int) -> None:
    """D[p, q] = 0 for every pair of rows of the same duplicate group (device tensor, in place)."""
    import torch

    n = rep.shape[0]
    dup_pos = np.arange(n_unique, n, dtype=np.int64)
    groups = {}
    for p, r in zip(dup_pos.tolist(), rep[n_unique:].tolist()):
        groups.setdefault(r, [r]).append(p)
    for members in groups.values():
        idx = torch.as_tensor(members, dtype=torch.long, device=D.device)
        D[idx[:, None], idx[None, :]] = 0


def row_to_all_distances(character_matrix, row_name, priors, missing_state_indicator: int,
                         prior_transformation: str = "negative_log",
                         dissimilarity_function: Optional[Callable] = None):
    """Distances from the sample `row_name` to every sample of `character_matrix` (float64, exact kernel),
    as {name: distance}.  One map row block on the GPU -- used for the implicit-root row when a map already
    exists (reference NeighborJoiningSolver.py:293-307 evaluates the function per leaf in Python)."""
    weights = None
    if priors:
        weights = solver_utilities.transform_priors(priors, prior_transformation)
    cm = host_prep.as_int_matrix(character_matrix)
    names = list(character_matrix.index)
    r = names.index(row_name)
    function = ("weighted_hamming_distance" if dissimilarity_function is None
                else _require_gpu_function(dissimilarity_function))
    lookup = "match" if function in _MATCH_LOOKUP else "mismatch"
    codes, missing_code, table, n_states = host_prep.compact_states(cm, missing_state_indicator, weights,
                                                                    lookup=lookup)
    D = engine.whd_map(codes, missing_code, table, n_states, dtype="f64", method="exact", rows=(r, r + 1),
                       function=function)
    row = D[0, : len(names)].cpu().numpy()
    return {nm: float(row[i]) for i, nm in enumerate(names)}
