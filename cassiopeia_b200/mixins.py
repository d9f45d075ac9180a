"""Exceptions and warnings of the path (reference: cassiopeia/mixins/errors.py, warnings.py).

When the real ``cassiopeia`` package is importable its classes are re-used so that
``except cassiopeia.mixins.DistanceSolverError`` keeps working for callers that swap
in the B200 solvers; otherwise equivalent classes are defined here.  If the reference
package is imported only AFTER this module (any import order must work for a drop-in),
an error raised here is an instance of a class derived from both this module's class and
the reference's class of the same name, so either ``except`` clause catches it.
"""
from __future__ import annotations

import sys

_combined = {}


def _with_reference(cls):
    """`cls`, or a subclass of it that also derives from the reference's class of the same name."""
    ref_mod = sys.modules.get("cassiopeia.mixins.errors")
    ref_cls = getattr(ref_mod, cls.__name__, None) if ref_mod is not None else None
    if ref_cls is None or not isinstance(ref_cls, type) or issubclass(cls, ref_cls):
        return cls
    key = (cls, ref_cls)
    if key not in _combined:
        _combined[key] = type(cls.__name__, (cls, ref_cls), {"__module__": cls.__module__})
    return _combined[key]

try:  # pragma: no cover - depends on the environment
    from cassiopeia.mixins.errors import (CassiopeiaError, CassiopeiaTreeError, DistanceSolverError,
                                          PriorTransformationError, SharedMutationJoiningSolverError)
    from cassiopeia.mixins.warnings import CassiopeiaTreeWarning, SharedMutationJoiningSolverWarning
except Exception:  # reference package absent (e.g. on the GPU box)

    class CassiopeiaError(Exception):
        """Base class of the errors raised on this path."""

        def __new__(cls, *args, **kwargs):
            return super().__new__(_with_reference(cls), *args, **kwargs)

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
