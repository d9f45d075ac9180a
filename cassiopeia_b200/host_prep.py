"""Host-side preparation of the dissimilarity-map inputs (numpy, O(n*m)).

* ``reference_row_order``: the row order of the reference's map
  (cassiopeia/data/CassiopeiaTree.py:1920-1951): unique rows in first-appearance order,
  then, for every unique row in that order, its duplicates in original order.  Computing
  the full map on the permuted matrix equals the reference's dedup-then-expand result
  (duplicates are at distance 0 from each other and share all other distances).
* ``compact_states``: remaps arbitrary integer state labels to the compact codes the
  kernels take (0 = uncut, 1..n_states, missing indicator untouched) and builds the dense
  ``[m, n_states+1]`` float64 weight table.  Equality of states is preserved, so
  distances are unchanged.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

from .mixins import DistanceSolverError

DIRECT_STATE_LIMIT = 4094  # raw labels in [0, limit] are used as codes without remapping


def as_int_matrix(character_matrix) -> np.ndarray:
    """DataFrame / ndarray -> contiguous int64 ndarray; ambiguous (object) states are rejected."""
    arr = character_matrix.to_numpy() if hasattr(character_matrix, "to_numpy") else np.asarray(character_matrix)
    if arr.dtype == object:
        try:
            arr = arr.astype(np.int64)
        except (TypeError, ValueError) as e:
            raise DistanceSolverError(
                "ambiguous (tuple) character states are not supported by the B200 kernels"
            ) from e
    if not np.issubdtype(arr.dtype, np.integer):
        if np.issubdtype(arr.dtype, np.floating) and np.all(np.isfinite(arr)) and np.all(arr == np.round(arr)):
            arr = arr.astype(np.int64)
        else:
            raise DistanceSolverError("character matrix must hold integer states")
    return np.ascontiguousarray(arr, dtype=np.int64)


def _first_occurrence_exact(cm: np.ndarray) -> np.ndarray:
    """first[r] = index of the first row equal to row r (sort of the rows as opaque byte strings)."""
    rows = cm.view(np.dtype((np.void, cm.dtype.itemsize * cm.shape[1]))).ravel()
    _, first_idx, inverse = np.unique(rows, return_index=True, return_inverse=True)
    return first_idx[inverse.ravel()]


# fixed odd 64-bit multipliers of the row hash (any values do: the hash only proposes groups, equality decides)
_HASH_MULT = np.random.default_rng(0x5EED).integers(1, 2**63, size=4096, dtype=np.uint64) | np.uint64(1)


def _first_occurrence(cm: np.ndarray) -> np.ndarray:
    """Same as `_first_occurrence_exact`, 10x faster on wide matrices: group the rows by a 64-bit
    multiplicative hash (one integer sort instead of a sort of m-word keys), then verify every row against
    the first row of its group; a hash collision (never seen) falls back to the exact sort."""
    n, m = cm.shape
    if m > _HASH_MULT.size or n < 2:
        return _first_occurrence_exact(cm)
    h = (cm.astype(np.uint64) * _HASH_MULT[:m]).sum(axis=1, dtype=np.uint64)  # wraps mod 2^64
    _, first_idx, inverse = np.unique(h, return_index=True, return_inverse=True)
    first = first_idx[inverse.ravel()]
    dup = np.flatnonzero(first != np.arange(n))
    if dup.size and not np.array_equal(cm[dup], cm[first[dup]]):
        return _first_occurrence_exact(cm)
    return first


def duplicate_groups(cm: np.ndarray, order: np.ndarray, first: Optional[np.ndarray] = None):
    """(n_unique, rep) for the matrix permuted by `order` = reference_row_order(cm): the first n_unique rows are
    the distinct rows; rep[p] = position (in the permuted matrix) of the distinct row that row p duplicates
    (rep[p] == p for the distinct rows).  `first`: the first-occurrence vector `reference_row_order(cm,
    return_first=True)` already computed for the same matrix."""
    cm = np.ascontiguousarray(cm)
    n = cm.shape[0]
    if n == 0 or cm.shape[1] == 0:
        return (min(n, 1), np.zeros(n, dtype=np.int64))
    if first is None:
        first = _first_occurrence(cm)
    pos_of = np.empty(n, dtype=np.int64)
    pos_of[order] = np.arange(n, dtype=np.int64)
    n_unique = int((first == np.arange(n)).sum())
    return n_unique, pos_of[first[order]]


def reference_row_order(cm: np.ndarray, return_first: bool = False):
    cm = np.ascontiguousarray(cm)
    n = cm.shape[0]
    if n == 0:
        return (np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)) if return_first else np.zeros(0, dtype=np.int64)
    if cm.shape[1] == 0:
        first = np.zeros(n, dtype=np.int64)
    else:
        first = _first_occurrence(cm)  # first occurrence of each row's content
    idx = np.arange(n, dtype=np.int64)
    is_first = first == idx
    uniq = idx[is_first]
    dups = idx[~is_first]
    # duplicates grouped by their unique row's first index (== unique-row order), then by own index
    dups = dups[np.lexsort((dups, first[dups]))]
    order = np.concatenate([uniq, dups])
    return (order, first) if return_first else order


def compact_states(cm: np.ndarray, missing: int, weights: Optional[Dict[int, Dict[int, float]]],
                   lookup: str = "mismatch", count_rows: Optional[int] = None
                   ) -> Tuple[np.ndarray, int, Optional[np.ndarray], int]:
    """Returns (codes int32 [n, m], missing_code, weight_table or None, n_states).

    `count_rows`: only the first `count_rows` rows decide which weights are looked up (the tree-level map of the
    reference evaluates the function on the DISTINCT rows only, CassiopeiaTree.py:1920-1932, so a state shared
    only by duplicated cells is never looked up there); every state present anywhere still gets its weight if the
    priors hold one.

    Raises KeyError where the reference's numba dict lookup would.  lookup="mismatch"
    (weighted_hamming_distance, exponential_negative_hamming_distance): a weighted character
    that shows at least two distinct present states must have a weight for each of its
    non-zero states (a state is only looked up when it takes part in a mismatch,
    dissimilarity_functions.py:56-69).  lookup="match" (the similarity functions, :104-110,
    :150-156, :228-233): a non-zero state is looked up when two samples share it, so every
    non-zero state carried by at least two samples needs a weight.
    """
    n, m = cm.shape
    present = cm != missing
    weighted = bool(weights)
    # smallest / largest PRESENT state without gathering the present entries: the indicator only matters
    # when it is itself the extreme value of the matrix
    nonneg, smax = True, 0
    if cm.size and present.any():
        lo, smax = int(cm.min()), int(cm.max())
        if lo < 0:
            # negative present states force the remapping; the indicator alone (the usual -1) does not
            nonneg = lo == missing and not ((cm < 0) & present).any()
        if smax == missing:
            smax = int(np.where(present, cm, lo).max())
    direct = nonneg and smax <= DIRECT_STATE_LIMIT
    if direct:
        codes = cm.astype(np.int32)
        n_states = max(smax, 1)
        missing_code = int(missing)
        if 0 <= missing_code <= n_states:
            # the indicator collides with the code range: move it out of the way
            missing_code = -1
            codes = np.where(present, codes, -1).astype(np.int32)
        elif not (-(2**31) <= missing_code < 2**31):
            missing_code = -1
            codes = np.where(present, codes, -1).astype(np.int32)
        label_of = None
    else:
        codes = np.full((n, m), -1, dtype=np.int32)
        missing_code = -1
        label_of = []
        n_states = 1
        for c in range(m):
            col = cm[:, c]
            pres = present[:, c]
            labels = np.unique(col[pres])
            nz = labels[labels != 0]
            lut_keys = np.concatenate([[0], nz])
            code = np.searchsorted(nz, col[pres]) + 1
            code = np.where(col[pres] == 0, 0, code)
            codes[pres, c] = code
            label_of.append(lut_keys)
            n_states = max(n_states, len(nz))
    table = None
    if weighted:
        table = np.zeros((m, n_states + 1), dtype=np.float64)
        # distinct present codes per character and how many samples carry each: codes are small
        # non-negative integers, so a counting pass per column replaces a sort
        width = n_states + 2  # bin 0 collects the missing entries
        itype = np.int32 if m * width < 2**31 - 1 else np.int64
        offsets = (1 + np.arange(m, dtype=np.int64) * width).astype(itype)
        if missing_code == -1:
            shifted = codes.astype(itype, copy=False) + offsets  # missing (-1) lands in bin 0 of its character
        else:
            shifted = np.where(codes != missing_code, codes, -1).astype(itype) + offsets
        hist2d = np.bincount(shifted.ravel(), minlength=m * width).reshape(m, width)[:, 1:]
        if count_rows is not None and count_rows < n:
            hist_rule = np.bincount(shifted[:count_rows].ravel(), minlength=m * width).reshape(m, width)[:, 1:]
        else:
            hist_rule = hist2d
        for c in range(m):
            hist = hist_rule[c]
            seen = np.flatnonzero(hist)
            counts = hist[seen]
            if lookup == "mismatch":
                if seen.size < 2:
                    continue  # no mismatch possible at this character: never looked up
                needed = seen[seen != 0]
            else:
                needed = seen[(seen != 0) & (counts >= 2)]
                if needed.size == 0:
                    continue
            if c not in weights:
                raise KeyError(c)
            wc = weights[c]
            for code in needed.tolist():
                label = code if label_of is None else int(label_of[c][code])
                if label not in wc:
                    raise KeyError(label)
                table[c, code] = wc[label]
            if hist_rule is not hist2d:
                # states carried only by duplicated rows: not part of the rule above, weight filled in when known
                for code in np.flatnonzero(hist2d[c]).tolist():
                    if code != 0 and table[c, code] == 0.0:
                        label = code if label_of is None else int(label_of[c][code])
                        if label in wc:
                            table[c, code] = wc[label]
    return np.ascontiguousarray(codes), missing_code, table, int(n_states)
