"""TEST INFRASTRUCTURE ONLY -- a CPU stand-in for ``cassiopeia_b200.engine`` backed by the oracle.

The drop-in tests (tests/test_reference_dropin.py) exercise the HOST side of the product -- the solver classes, their
handling of the reference's real ``CassiopeiaTree`` objects, rooting, collapsing -- on the CPU box, where the CUDA
library cannot run.  They monkeypatch the engine's entry points with these functions for the duration of a test; the
product itself never imports this module and has no CPU path."""
import numpy as np
import torch

from cassiopeia_b200 import engine
from oracle import pyoracle as po


def _ld(n):
    return engine.padded_ld(n)


def whd_map(cm, missing, weights, n_states, dtype="f64", method="exact", rows=None, out=None,
            function="weighted_hamming_distance", ignore_missing=False):
    cm = np.asarray(cm.cpu().numpy() if hasattr(cm, "cpu") else cm, dtype=np.int64)
    n = cm.shape[0]
    table = None if weights is None else np.asarray(weights, dtype=np.float64)
    if function == "weighted_hamming_distance":
        full = po.whd_map(cm, int(missing), table)
    else:
        full = po.pairwise_map(function, cm, int(missing), table, ignore_missing=ignore_missing)
    r0, r1 = (0, n) if rows is None else rows
    host = np.zeros((r1 - r0, _ld(n)), dtype=np.float64 if dtype == "f64" else np.float32)
    host[:, :n] = full[r0:r1]
    t = torch.from_numpy(host)
    if out is not None:
        out[: r1 - r0].copy_(t)
        return out
    return t


def agglomerate(D, n, algo, mode, max_iters=-1):
    M = np.ascontiguousarray(D[:, :n].double().numpy())
    fn = po.nj if algo == "nj" else po.upgma
    return fn(M, max_iters=max_iters)


def nj_q(D):
    D = np.asarray(D, dtype=np.float64)
    n = D.shape[0]
    rs = np.zeros(n)
    for i in range(n):
        s = 0.0
        for j in range(n):
            s += D[i, j]
        rs[i] = s
    q = np.zeros((n, n))
    t = 1 / (n - 2)
    for i in range(n):
        for j in range(i):
            q[i, j] = q[j, i] = D[i, j] - (t * (rs[i] + rs[j]))
    return q


def update_map(D, i, j, algo, size_i=1, size_j=1):
    D = np.asarray(D, dtype=np.float64)
    n = D.shape[0]
    keep = [v for v in range(n) if v not in (i, j)]
    if algo == "nj":
        new = np.array([0.5 * ((D[v, i] + D[v, j]) - D[i, j]) for v in keep])
    else:
        new = np.array([((size_i * D[v, i]) + (size_j * D[v, j])) / (size_i + size_j) for v in keep])
    out = np.zeros((n - 1, n - 1))
    out[: n - 2, : n - 2] = D[np.ix_(keep, keep)]
    out[n - 2, : n - 2] = new
    out[: n - 2, n - 2] = new
    return out


def install(monkeypatch):
    """Route the engine's compute entry points to the oracle for one test."""
    for name, fn in (("whd_map", whd_map), ("agglomerate", agglomerate), ("nj_q", nj_q), ("update_map", update_map)):
        monkeypatch.setattr(engine, name, fn)
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self, raising=False)
