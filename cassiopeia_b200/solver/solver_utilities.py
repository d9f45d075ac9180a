"""Host-side helpers of the solvers (reference: cassiopeia/solver/solver_utilities.py)."""
from __future__ import annotations

import itertools
import os
from typing import Dict, Generator, Optional

import numpy as np

from ..mixins import PriorTransformationError

_TRANSFORMS = ("negative_log", "inverse", "square_root_inverse")


def node_name_generator() -> Generator[str, None, None]:
    """Unique internal-node names (reference solver_utilities.py:16-29 hashes timestamps;
    the names are arbitrary, only their uniqueness matters)."""
    salt = os.urandom(6).hex()
    for i in itertools.count():
        yield f"cassiopeia_internal_node{salt}{i:012x}"


def transform_priors(priors: Optional[Dict[int, Dict[int, float]]],
                     prior_transformation: str = "negative_log") -> Dict[int, Dict[int, float]]:
    """Prior probabilities -> weights, numpy float64 like the reference
    (solver_utilities.py:53-103): -log p | 1/p | sqrt(1/p); p must lie in (0, 1]."""
    if prior_transformation not in _TRANSFORMS:
        raise PriorTransformationError("Please select one of the supported prior transformations.")
    weights: Dict[int, Dict[int, float]] = {}
    for character, states in priors.items():
        out = {}
        for state, p in states.items():
            if p <= 0.0 or p > 1.0:
                raise PriorTransformationError("Please make sure all priors have a value between 0 and 1")
            if prior_transformation == "negative_log":
                out[state] = -np.log(p)
            elif prior_transformation == "inverse":
                out[state] = 1 / p
            else:
                out[state] = np.sqrt(1 / p)
        weights[character] = out
    return weights
