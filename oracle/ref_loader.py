"""TEST INFRASTRUCTURE ONLY -- loader for the *reference* hot path.

Imports the reference's own Python/numba implementation of the distance-solver
path from ``/root/reference`` (read-only, present only in the build container,
never on the GPU box).  It is used by ``oracle/make_golden.py`` to produce the
golden fixtures under ``tests/golden/`` and by ``tests/test_oracle_vs_reference.py``
(skipped when ``/root/reference`` is absent) to pin the C restatement in
``oracle/oracle.c``.  Nothing in ``cassiopeia_b200/`` may import this module.

``import cassiopeia`` fails in this image (``ete3``, ``ngs_tools``, ``treedata``,
``pysam`` ... are not installed and ``cassiopeia/__init__.py:5-12`` imports every
sub-package), so we register stub modules for the missing third-party packages
and bare package shells for the heavy sub-packages, then import only the
modules on the path.
"""
from __future__ import annotations

import importlib
import logging
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("CASSIOPEIA_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "cassiopeia", "solver"))


class _Permissive:
    """Stand-in class for third-party types that are only named, never used."""

    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):  # pragma: no cover - never reached on the path
        raise AttributeError(name)


def _stub(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def _shell(name: str, subdir: str) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__path__ = [os.path.join(REFERENCE_ROOT, "cassiopeia", subdir)] if subdir else [
        os.path.join(REFERENCE_ROOT, "cassiopeia")
    ]
    mod.__package__ = name
    sys.modules[name] = mod
    return mod


_loaded = None


def load_reference():
    """Returns a namespace with the reference classes/functions on the hot path."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    if "cassiopeia" in sys.modules and not getattr(sys.modules["cassiopeia"], "_b200_shell", False):
        raise RuntimeError("a different `cassiopeia` package is already imported")

    # third-party packages that are named at import time but unused on the path
    if "ete3" not in sys.modules:
        _stub("ete3", Tree=_Permissive, TreeNode=_Permissive)
    if "ngs_tools" not in sys.modules:
        ngs = _stub("ngs_tools")
        ngs.logging = types.SimpleNamespace(Logger=lambda name: logging.getLogger(name))
        _stub("ngs_tools.logging", Logger=lambda name: logging.getLogger(name))
        ngs.utils = types.SimpleNamespace()
    for name in ("treedata", "pysam", "pylab"):
        if name not in sys.modules:
            _stub(name, TreeData=_Permissive)

    top = _shell("cassiopeia", "")
    top._b200_shell = True
    for sub in ("solver", "preprocess", "simulator", "critique"):
        setattr(top, sub, _shell(f"cassiopeia.{sub}", sub))

    mixins = importlib.import_module("cassiopeia.mixins")
    data = importlib.import_module("cassiopeia.data")
    top.mixins, top.data = mixins, data

    solver = sys.modules["cassiopeia.solver"]
    ds = importlib.import_module("cassiopeia.solver.DistanceSolver")
    nj = importlib.import_module("cassiopeia.solver.NeighborJoiningSolver")
    up = importlib.import_module("cassiopeia.solver.UPGMASolver")
    cs = importlib.import_module("cassiopeia.solver.CassiopeiaSolver")
    solver.CassiopeiaSolver = cs
    dfn = importlib.import_module("cassiopeia.solver.dissimilarity_functions")
    sut = importlib.import_module("cassiopeia.solver.solver_utilities")
    solver.dissimilarity_functions = dfn
    solver.dissimilarity = dfn
    solver.solver_utilities = sut
    solver.NeighborJoiningSolver = nj.NeighborJoiningSolver
    solver.UPGMASolver = up.UPGMASolver
    solver.DistanceSolver = ds.DistanceSolver
    smj = importlib.import_module("cassiopeia.solver.SharedMutationJoiningSolver")
    solver.SharedMutationJoiningSolver = smj.SharedMutationJoiningSolver

    sim = sys.modules["cassiopeia.simulator"]
    cbs = importlib.import_module("cassiopeia.simulator.CompleteBinarySimulator")
    cas9 = importlib.import_module("cassiopeia.simulator.Cas9LineageTracingDataSimulator")
    sim.CompleteBinarySimulator = cbs.CompleteBinarySimulator
    sim.Cas9LineageTracingDataSimulator = cas9.Cas9LineageTracingDataSimulator

    _loaded = types.SimpleNamespace(
        cassiopeia=top,
        CassiopeiaTree=data.CassiopeiaTree,
        data_utilities=importlib.import_module("cassiopeia.data.utilities"),
        DistanceSolver=ds.DistanceSolver,
        NeighborJoiningSolver=nj.NeighborJoiningSolver,
        UPGMASolver=up.UPGMASolver,
        SharedMutationJoiningSolver=smj.SharedMutationJoiningSolver,
        dissimilarity_functions=dfn,
        solver_utilities=sut,
        CompleteBinarySimulator=cbs.CompleteBinarySimulator,
        Cas9LineageTracingDataSimulator=cas9.Cas9LineageTracingDataSimulator,
        DistanceSolverError=mixins.DistanceSolverError,
        CassiopeiaTreeError=mixins.CassiopeiaTreeError,
        PriorTransformationError=mixins.PriorTransformationError,
    )
    return _loaded


class _JoinRecorder:
    """Mixin that records the reference's join sequence in *id space*.

    Leaf p (position in the initial dissimilarity map) has id p; the t-th new
    node has id n+t.  ``find_cherry`` (reference ``DistanceSolver.py:187``)
    returns positions in the current map whose row order is "delete i and j
    preserving order, append the new node last" (``NeighborJoiningSolver.py:203-214``)
    so a parallel python list of ids reproduces the bookkeeping.
    """

    def find_cherry(self, dissimilarity_matrix):
        n = dissimilarity_matrix.shape[0]
        if getattr(self, "_rec_ids", None) is None or len(self._rec_ids) != n:
            # first call of a solve(): ids are 0..n-1
            self._rec_ids = list(range(n))
            self._rec_next = n
            self.recorded_joins = []
        i, j = super().find_cherry(dissimilarity_matrix)
        i, j = int(i), int(j)
        ids = self._rec_ids
        self.recorded_joins.append((ids[i], ids[j]))
        a, b = ids[i], ids[j]
        ids.remove(a)
        ids.remove(b)
        ids.append(self._rec_next)
        self._rec_next += 1
        return i, j


def recording_smj_solver():
    """Reference SharedMutationJoiningSolver subclass that exposes ``recorded_joins`` (n-1 joins)."""
    ref = load_reference()

    class RecSMJ(_JoinRecorder, ref.SharedMutationJoiningSolver):
        pass

    return RecSMJ


def recording_solvers():
    """Reference NJ / UPGMA solver subclasses that expose ``recorded_joins``."""
    ref = load_reference()

    class RecNJ(_JoinRecorder, ref.NeighborJoiningSolver):
        pass

    class RecUPGMA(_JoinRecorder, ref.UPGMASolver):
        pass

    return RecNJ, RecUPGMA
