"""Exceptions and warnings of the path (reference: cassiopeia/mixins/errors.py, warnings.py).

When the real ``cassiopeia`` package is importable its classes are re-used so that
``except cassiopeia.mixins.DistanceSolverError`` keeps working for callers that swap
in the B200 solvers; otherwise equivalent classes are defined here.
"""
from __future__ import annotations

try:  # pragma: no cover - depends on the environment
    from cassiopeia.mixins.errors import (CassiopeiaError, CassiopeiaTreeError, DistanceSolverError,
                                          PriorTransformationError, SharedMutationJoiningSolverError)
    from cassiopeia.mixins.warnings import CassiopeiaTreeWarning, SharedMutationJoiningSolverWarning
except Exception:  # reference package absent (e.g. on the GPU box)

    class CassiopeiaError(Exception):
        """Base class of the errors raised on this path."""

    class CassiopeiaTreeError(CassiopeiaError):
        """Errors of the tree container (cassiopeia/mixins/errors.py)."""

    class DistanceSolverError(CassiopeiaError):
        """Errors of the distance solvers (cassiopeia/mixins/errors.py:25)."""

    class PriorTransformationError(CassiopeiaError):
        """Invalid prior transformation or prior value."""

    class SharedMutationJoiningSolverError(CassiopeiaError):
        """Errors of the SharedMutationJoiningSolver (cassiopeia/mixins/errors.py:77)."""

    class CassiopeiaTreeWarning(UserWarning):
        """Warnings of the tree container."""

    class SharedMutationJoiningSolverWarning(UserWarning):
        """Warnings of the SharedMutationJoiningSolver (cassiopeia/mixins/warnings.py:14)."""


def is_ambiguous_state(state) -> bool:
    """A tuple/list state is 'ambiguous' (cassiopeia/mixins/utilities.py)."""
    return isinstance(state, (tuple, list))
