"""TEST INFRASTRUCTURE ONLY -- ctypes/numpy front-end of ``oracle/oracle.c``.

Also holds small numpy restatements of the host-side steps of the reference
path (prior transformation, de-duplication row order) so tests can check the
product's host logic against something independent of it.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Dict, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compiles oracle.c -> liboracle.so (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "liboracle.so"])
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        i64, i32, dbl, vp = ctypes.c_int64, ctypes.c_int32, ctypes.c_double, ctypes.c_void_p
        L.oracle_whd_pair.restype = dbl
        L.oracle_whd_pair.argtypes = [vp, vp, ctypes.c_int, i64, vp, ctypes.c_int]
        L.oracle_whd_map_rows.restype = None
        L.oracle_whd_map_rows.argtypes = [vp, i64, ctypes.c_int, i64, vp, ctypes.c_int, i64, i64, vp, ctypes.c_int]
        L.oracle_pair_fn.restype = dbl
        L.oracle_pair_fn.argtypes = [ctypes.c_int, ctypes.c_int, vp, vp, ctypes.c_int, i64, vp, ctypes.c_int]
        L.oracle_pairwise_map_rows.restype = None
        L.oracle_pairwise_map_rows.argtypes = [ctypes.c_int, ctypes.c_int, vp, i64, ctypes.c_int, i64, vp,
                                               ctypes.c_int, i64, i64, vp, ctypes.c_int]
        L.oracle_smj.restype = ctypes.c_int
        L.oracle_smj.argtypes = [ctypes.c_int, vp, i64, ctypes.c_int, i64, vp, ctypes.c_int, vp, vp]
        L.oracle_nj.restype = ctypes.c_int
        L.oracle_nj.argtypes = [vp, i64, vp, vp, i64, ctypes.c_int]
        L.oracle_upgma.restype = ctypes.c_int
        L.oracle_upgma.argtypes = [vp, i64, vp, vp, i64, ctypes.c_int]
        L.oracle_nj_adaptive.restype = ctypes.c_int
        L.oracle_nj_adaptive.argtypes = [vp, i64, vp, vp, vp, ctypes.c_int]
        L.oracle_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def max_threads() -> int:
    return int(lib().oracle_max_threads())


# --------------------------------------------------------------------------
# host-side restatements


def transform_priors(priors: Dict[int, Dict[int, float]], how: str = "negative_log"):
    """cassiopeia/solver/solver_utilities.py:53-103, numpy float64."""
    if how not in ("negative_log", "inverse", "square_root_inverse"):
        raise ValueError("unsupported prior transformation")
    out = {}
    for c, states in priors.items():
        o = {}
        for s, p in states.items():
            if p <= 0.0 or p > 1.0:
                raise ValueError("prior outside (0, 1]")
            if how == "negative_log":
                o[s] = -np.log(p)
            elif how == "inverse":
                o[s] = 1 / p
            else:
                o[s] = np.sqrt(1 / p)
        out[c] = o
    return out


def dense_weight_table(weights: Optional[Dict[int, Dict[int, float]]], m: int, max_state: int):
    """[m, max_state+1] float64 table indexed by raw state value; NaN = no prior."""
    if not weights:
        return None
    W = max_state + 1
    tab = np.full((m, W), np.nan, dtype=np.float64)
    for c, sw in weights.items():
        for s, w in sw.items():
            if 0 <= s < W and 0 <= c < m:
                tab[c, s] = w
    return tab


def reference_row_order(cm: np.ndarray) -> np.ndarray:
    """Row order of the reference's map (cassiopeia/data/CassiopeiaTree.py:1920-1951):
    unique rows in first-appearance order, then, for each unique row in that
    order, its duplicates in original order."""
    seen = {}
    uniq = []
    dups = {}
    for r in range(cm.shape[0]):
        key = cm[r].tobytes()
        if key in seen:
            dups[seen[key]].append(r)
        else:
            seen[key] = r
            uniq.append(r)
            dups[r] = []
    order = list(uniq)
    for u in uniq:
        order.extend(dups[u])
    return np.asarray(order, dtype=np.int64)


# --------------------------------------------------------------------------
# C entry points


def whd_pair(s1, s2, missing=-1, table: Optional[np.ndarray] = None) -> float:
    a = np.ascontiguousarray(s1, dtype=np.int64)
    b = np.ascontiguousarray(s2, dtype=np.int64)
    if table is not None:
        table = np.ascontiguousarray(table, dtype=np.float64)
    return float(
        lib().oracle_whd_pair(
            a.ctypes.data, b.ctypes.data, int(a.shape[0]), int(missing),
            table.ctypes.data if table is not None else None,
            int(table.shape[1]) if table is not None else 0,
        )
    )


def whd_map(cm: np.ndarray, missing: int = -1, table: Optional[np.ndarray] = None,
            rows: Optional[Tuple[int, int]] = None, threads: int = 0) -> np.ndarray:
    """Full symmetric map (or the row block `rows`) in float64."""
    cm = np.ascontiguousarray(cm, dtype=np.int64)
    n, m = cm.shape
    if table is not None:
        table = np.ascontiguousarray(table, dtype=np.float64)
        assert table.shape[0] == m
    r0, r1 = (0, n) if rows is None else rows
    out = np.empty((r1 - r0, n), dtype=np.float64)
    lib().oracle_whd_map_rows(
        cm.ctypes.data, n, m, int(missing),
        table.ctypes.data if table is not None else None,
        int(table.shape[1]) if table is not None else 0,
        int(r0), int(r1), out.ctypes.data, int(threads),
    )
    return out


#: pair-function codes (the product's CAS_FN_* values; names = reference function names)
PAIR_FUNCTIONS = {
    "weighted_hamming_distance": 0,
    "hamming_similarity_without_missing": 1,
    "hamming_similarity_normalized_over_missing": 2,
    "hamming_distance": 3,
    "weighted_hamming_similarity": 4,
    "exponential_negative_hamming_distance": 5,
}


def pair_fn(fn: str, s1, s2, missing=-1, table: Optional[np.ndarray] = None, ignore_missing: bool = False) -> float:
    a = np.ascontiguousarray(s1, dtype=np.int64)
    b = np.ascontiguousarray(s2, dtype=np.int64)
    if table is not None:
        table = np.ascontiguousarray(table, dtype=np.float64)
    return float(lib().oracle_pair_fn(
        PAIR_FUNCTIONS[fn], 1 if ignore_missing else 0, a.ctypes.data, b.ctypes.data, int(a.shape[0]),
        int(missing), table.ctypes.data if table is not None else None,
        int(table.shape[1]) if table is not None else 0))


def pairwise_map(fn: str, cm: np.ndarray, missing: int = -1, table: Optional[np.ndarray] = None,
                 ignore_missing: bool = False, threads: int = 0) -> np.ndarray:
    """Full symmetric float64 map of any pair function of the reference's dissimilarity_functions.py."""
    cm = np.ascontiguousarray(cm, dtype=np.int64)
    n, m = cm.shape
    if table is not None:
        table = np.ascontiguousarray(table, dtype=np.float64)
    out = np.empty((n, n), dtype=np.float64)
    lib().oracle_pairwise_map_rows(
        PAIR_FUNCTIONS[fn], 1 if ignore_missing else 0, cm.ctypes.data, n, m, int(missing),
        table.ctypes.data if table is not None else None, int(table.shape[1]) if table is not None else 0,
        0, n, out.ctypes.data, int(threads))
    return out


def _loop(fn, D: np.ndarray, max_iters: int, threads: int):
    D = np.array(D, dtype=np.float64, order="C", copy=True)
    n = D.shape[0]
    assert D.shape == (n, n)
    total = max(n - 2, 0)
    joins = np.full((total, 2), -1, dtype=np.int32)
    last2 = np.full(2, -1, dtype=np.int32)
    rc = fn(D.ctypes.data, n, joins.ctypes.data, last2.ctypes.data, int(max_iters), int(threads))
    if rc != 0:
        raise RuntimeError(f"oracle loop failed rc={rc}")
    if max_iters >= 0:
        joins = joins[: min(total, max_iters)]
    return joins, last2


def nj(D: np.ndarray, max_iters: int = -1, threads: int = 0):
    """Join list (ids) of the reference NJ loop on the symmetric map D."""
    return _loop(lib().oracle_nj, D, max_iters, threads)


def upgma(D: np.ndarray, max_iters: int = -1, threads: int = 0):
    return _loop(lib().oracle_upgma, D, max_iters, threads)


def nj_adaptive(D: np.ndarray, threads: int = 0):
    """CPU model of the product's adaptive-strict NJ loop (maintained row sums + rigorous error window +
    sequential sums only for the rows of ambiguous pairs).  Returns (joins, last2, stats) with stats =
    {ambiguous_iterations, rows_resolved, max_rows, max_window, winner_changed}."""
    D = np.array(D, dtype=np.float64, order="C", copy=True)
    n = D.shape[0]
    joins = np.full((max(n - 2, 0), 2), -1, dtype=np.int32)
    last2 = np.full(2, -1, dtype=np.int32)
    st = np.zeros(5, dtype=np.int64)
    rc = lib().oracle_nj_adaptive(D.ctypes.data, n, joins.ctypes.data, last2.ctypes.data, st.ctypes.data, int(threads))
    if rc != 0:
        raise RuntimeError(f"oracle_nj_adaptive failed rc={rc}")
    stats = {"ambiguous_iterations": int(st[0]), "rows_resolved": int(st[1]), "max_rows": int(st[2]),
             "max_window": float(st[3:4].view(np.float64)[0]), "winner_changed": int(st[4])}
    return joins, last2, stats


def smj(fn: str, cm: np.ndarray, missing: int = -1, table: Optional[np.ndarray] = None):
    """Join list (ids, n-1 joins: the last one is the root) of the reference's SharedMutationJoiningSolver loop
    with the similarity function `fn` on the rows of `cm` as given (the solver does not de-duplicate)."""
    cm = np.ascontiguousarray(cm, dtype=np.int64)
    n, m = cm.shape
    if table is not None:
        table = np.ascontiguousarray(table, dtype=np.float64)
    S = pairwise_map(fn, cm, missing, table)
    joins = np.full((max(n - 1, 0), 2), -1, dtype=np.int32)
    rc = lib().oracle_smj(PAIR_FUNCTIONS[fn], cm.ctypes.data, n, m, int(missing),
                          table.ctypes.data if table is not None else None,
                          int(table.shape[1]) if table is not None else 0, S.ctypes.data, joins.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oracle smj failed rc={rc}")
    return joins
