"""Dissimilarity functions understood by the B200 solvers.

``weighted_hamming_distance`` (reference: cassiopeia/solver/dissimilarity_functions.py:12-72)
is the function the CUDA map kernels implement.  The Python body below is the
single-pair definition used for O(n) host work (e.g. the implicit-root row when a map
already exists, NeighborJoiningSolver.py:293-307); whole maps never go through it.
"""
from __future__ import annotations

from typing import Dict, List, Optional


def weighted_hamming_distance(s1: List[int], s2: List[int], missing_state_indicator=-1,
                              weights: Optional[Dict[int, Dict[int, float]]] = None) -> float:
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


def is_gpu_kernel(fn) -> bool:
    """True when `fn` is (an alias of) the function the CUDA kernels implement."""
    if fn is weighted_hamming_distance:
        return True
    mod = getattr(fn, "__module__", "") or ""
    return getattr(fn, "__name__", "") == "weighted_hamming_distance" and mod.endswith("dissimilarity_functions")
