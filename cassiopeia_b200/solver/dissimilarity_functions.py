"""Pair functions understood by the B200 map kernels.

Same names, signatures and values as cassiopeia/solver/dissimilarity_functions.py:12-300.  The CUDA
kernels (``cas_whd_map`` / ``cas_pairwise_map``) implement every one of them for whole maps; the Python
bodies below are the single-pair definitions, used only for O(n) host work (e.g. the implicit-root row when
a map already exists, NeighborJoiningSolver.py:293-307) and as documentation of what the kernels compute.
Whole maps never go through them.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np


def weighted_hamming_distance(s1: List[int], s2: List[int], missing_state_indicator=-1,
                              weights: Optional[Dict[int, Dict[int, float]]] = None) -> float:
    """Reference :12-72."""
    d = 0
    num_present = 0
    for i in range(len(s1)):
        a, b = s1[i], s2[i]
        if a == missing_state_indicator or b == missing_state_indicator:
            continue
        num_present += 1
        if a != b:
            if a == 0 or b == 0:
                if weights:
                    d += weights[i][a] if a != 0 else weights[i][b]
                else:
                    d += 1
            elif weights:
                d += weights[i][a] + weights[i][b]
            else:
                d += 2
    if num_present == 0:
        return 0
    return d / num_present


def hamming_similarity_without_missing(s1: List[int], s2: List[int], missing_state_indicator: int,
                                       weights: Optional[Dict[int, Dict[int, float]]] = None) -> float:
    """Number (or weight) of shared mutations, missing and uncut positions skipped (reference :75-113)."""
    similarity = 0
    for i in range(len(s1)):
        a, b = s1[i], s2[i]
        if a == missing_state_indicator or b == missing_state_indicator or a == 0 or b == 0:
            continue
        if a == b:
            similarity += weights[i][a] if weights else 1
    return similarity


def hamming_similarity_normalized_over_missing(s1: List[int], s2: List[int], missing_state_indicator: int,
                                               weights: Optional[Dict[int, Dict[int, float]]] = None) -> float:
    """Shared mutations over the number of positions present in both samples (reference :116-160)."""
    similarity = 0
    num_present = 0
    for i in range(len(s1)):
        a, b = s1[i], s2[i]
        if a == missing_state_indicator or b == missing_state_indicator:
            continue
        num_present += 1
        if a == 0 or b == 0:
            continue
        if a == b:
            similarity += weights[i][a] if weights else 1
    if num_present == 0:
        return 0
    return similarity / num_present


def hamming_distance(s1, s2, ignore_missing_state: bool = False, missing_state_indicator: int = -1) -> int:
    """Positions at which the samples disagree; missing counts as a state unless ignored (reference :163-194)."""
    dist = 0
    for i in range(len(s1)):
        if s1[i] != s2[i]:
            if (s1[i] == missing_state_indicator or s2[i] == missing_state_indicator) and ignore_missing_state:
                dist += 0
            else:
                dist += 1
    return dist


def weighted_hamming_similarity(s1: List[int], s2: List[int], missing_state_indicator: int,
                                weights: Optional[Dict[int, Dict[int, float]]] = None) -> float:
    """Reference :197-240: +2 w (or 2) per shared mutation, +1 per shared uncut site when unweighted."""
    d = 0
    num_present = 0
    for i in range(len(s1)):
        a, b = s1[i], s2[i]
        if a == missing_state_indicator or b == missing_state_indicator:
            continue
        num_present += 1
        if a == b:
            if a != 0:
                d += 2 * weights[i][a] if weights else 2
            elif not weights:
                d += 1
    if num_present == 0:
        return 0
    return d / num_present


def exponential_negative_hamming_distance(s1: List[int], s2: List[int], missing_state_indicator=-1,
                                          weights: Optional[Dict[int, Dict[int, float]]] = None) -> float:
    """exp(-weighted_hamming_distance); 0 when no position is present in both (reference :243-300)."""
    if not any(a != missing_state_indicator and b != missing_state_indicator for a, b in zip(s1, s2)):
        return 0
    return np.exp(-weighted_hamming_distance(s1, s2, missing_state_indicator, weights))


_KERNEL_FUNCTIONS = ("weighted_hamming_distance", "hamming_similarity_without_missing",
                     "hamming_similarity_normalized_over_missing", "weighted_hamming_similarity",
                     "exponential_negative_hamming_distance")


def gpu_kernel_name(fn) -> Optional[str]:
    """Name of the CUDA pair function that implements `fn` (ours or the reference's function of the same
    name), or None.  ``hamming_distance`` has a different signature (its third positional argument is
    ``ignore_missing_state``) and cannot be passed as a map function by the reference either; it is reachable
    through ``engine.whd_map(function="hamming_distance")``."""
    name = getattr(fn, "__name__", "")
    if name not in _KERNEL_FUNCTIONS:
        py_func = getattr(fn, "py_func", None)  # a numba dispatcher of the reference
        name = getattr(py_func, "__name__", "")
        if name not in _KERNEL_FUNCTIONS:
            return None
    if fn is globals().get(name):
        return name
    mod = getattr(fn, "__module__", "") or getattr(getattr(fn, "py_func", None), "__module__", "") or ""
    return name if mod.endswith("dissimilarity_functions") else None


def is_gpu_kernel(fn) -> bool:
    """True when `fn` is (an alias of) a function the CUDA kernels implement."""
    return gpu_kernel_name(fn) is not None
