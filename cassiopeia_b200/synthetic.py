"""Vectorised synthetic lineage-tracing inputs ("large inputs").

The reference's simulators (``CompleteBinarySimulator`` + ``Cas9LineageTracingDataSimulator``)
are per-node Python; this generator follows the same process with numpy so that 10^4-10^5
cells take seconds: a complete binary tree of depth ceil(log2 n) truncated to n leaves; on
every edge each still-uncut site mutates with probability 1 - exp(-rate * t), t = 1/(depth+1),
to a state drawn from that character's prior; states are inherited; finally each leaf entry
drops out to the missing indicator independently.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np


def simulate_character_matrix(n_cells: int, n_characters: int, n_states: int, seed: int = 0,
                              mutation_rate: float = 0.1, missing_fraction: float = 0.1,
                              missing_state_indicator: int = -1
                              ) -> Tuple[np.ndarray, Dict[int, Dict[int, float]], np.ndarray]:
    """Returns (int32 [n_cells, n_characters] matrix, priors dict, priors table [m, S+1] with NaN at 0)."""
    rng = np.random.default_rng(seed)
    m, S = int(n_characters), int(n_states)
    p = rng.exponential(1.0, size=(m, S))
    p /= p.sum(axis=1, keepdims=True)
    cdf = np.cumsum(p, axis=1)
    cdf[:, -1] = 1.0
    depth = max(1, math.ceil(math.log2(max(n_cells, 2))))
    p_edge = 1.0 - math.exp(-mutation_rate / (depth + 1))
    states = np.zeros((1, m), dtype=np.int32)
    for level in range(1, depth + 1):
        states = np.repeat(states, 2, axis=0)
        # only ancestors of the first n_cells leaves are simulated
        keep = math.ceil(n_cells / (2 ** (depth - level)))
        states = states[:keep]
        mutate = (states == 0) & (rng.random(states.shape) < p_edge)
        u = rng.random(states.shape)
        for c in range(m):
            rows = np.nonzero(mutate[:, c])[0]
            if rows.size:
                states[rows, c] = np.searchsorted(cdf[c], u[rows, c], side="right").astype(np.int32) + 1
    cm = np.ascontiguousarray(states[:n_cells])
    np.minimum(cm, S, out=cm)
    drop = rng.random(cm.shape) < missing_fraction
    cm[drop] = missing_state_indicator
    priors = {c: {s + 1: float(p[c, s]) for s in range(S)} for c in range(m)}
    table = np.concatenate([np.full((m, 1), np.nan), p], axis=1)
    return cm, priors, table
