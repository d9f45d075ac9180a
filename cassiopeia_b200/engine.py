"""Thin device-side driver over the C ABI.

PyTorch is used only for plumbing: it owns the device buffers (character matrix,
weights, the n x ld dissimilarity map, scratch) and supplies the CUDA stream.
All arithmetic happens in the hand-written kernels of ``libcas_b200.so``.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import numpy as np

from . import _lib
from ._lib import (CAS_ALGO_NJ, CAS_ALGO_UPGMA, CAS_F32, CAS_F64, CAS_MAP_EXACT, CAS_MAP_TENSOR,
                   CAS_MODE_EXACT, CAS_MODE_FAST, CAS_MODE_STRICT, B200LibraryError)

_ALGO = {"nj": CAS_ALGO_NJ, "upgma": CAS_ALGO_UPGMA}
#: loop arithmetic (include/cas_b200.h).  "exact" and "strict" both reproduce the reference's join list; "exact"
#: evaluates the reference's sequential row sums only where the maintained sums cannot decide.
_MODE = {"strict": CAS_MODE_STRICT, "fast": CAS_MODE_FAST, "exact": CAS_MODE_EXACT}
_METHOD = {"exact": CAS_MAP_EXACT, "tensor": CAS_MAP_TENSOR}


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise B200LibraryError("no CUDA device visible: the B200 kernels cannot run (no CPU fallback)")
    return torch


def padded_ld(n: int) -> int:
    """Row pitch (elements) of a map with n columns: multiple of 32 => 128-byte aligned rows."""
    return (int(n) + 31) // 32 * 32


def _stream_ptr(torch) -> int:
    return int(torch.cuda.current_stream().cuda_stream)


def device_info() -> dict:
    L = _lib.load()
    _torch()
    sms, maj, mnr = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    fr, tot = ctypes.c_size_t(), ctypes.c_size_t()
    _lib.check(L.cas_device_info(ctypes.byref(sms), ctypes.byref(maj), ctypes.byref(mnr),
                                 ctypes.byref(fr), ctypes.byref(tot)), "cas_device_info")
    return {"sm_count": sms.value, "cc": (maj.value, mnr.value), "free_bytes": fr.value,
            "total_bytes": tot.value}


#: diagnostics of the most recent `agglomerate` call: near / exact ties and the number of map elements the
#: pruned scan actually read (0 for the dense scan, which reads k(k-1)/2 per iteration)
last_stats = {"near_ties": 0, "exact_ties": 0, "scanned_elements": 0, "rows_listed": 0, "drift": 0.0,
              "ambiguous_iterations": 0, "rows_resolved": 0, "max_rows_resolved": 0, "pairs_changed": 0}


def last_launch_count() -> int:
    return int(_lib.load().cas_last_launch_count())


#: pair-function codes of the C ABI (CAS_FN_* in include/cas_b200.h), keyed by the reference's function names
PAIR_FUNCTIONS = {
    "weighted_hamming_distance": 0,
    "hamming_similarity_without_missing": 1,
    "hamming_similarity_normalized_over_missing": 2,
    "hamming_distance": 3,
    "weighted_hamming_similarity": 4,
    "exponential_negative_hamming_distance": 5,
}


def whd_map(cm, missing: int, weights: Optional[np.ndarray], n_states: int, dtype: str = "f64",
            method: str = "exact", rows: Optional[Tuple[int, int]] = None, out=None,
            function: str = "weighted_hamming_distance", ignore_missing: bool = False):
    """Pairwise weighted-Hamming map of an int32 [n, m] matrix of compact state codes.

    cm may be a numpy array (copied to the device) or a CUDA int32 tensor.  Returns a
    CUDA tensor of shape [rows, ld] (ld = padded_ld(n)); columns >= n are padding.

    `function` names any pair function of the reference's dissimilarity_functions.py (PAIR_FUNCTIONS);
    all but `weighted_hamming_distance` run on the exact CUDA-core kernel only.  `ignore_missing` is
    `hamming_distance`'s `ignore_missing_state`.
    """
    torch = _torch()
    L = _lib.load()
    if isinstance(cm, np.ndarray):
        cm_t = torch.from_numpy(np.ascontiguousarray(cm, dtype=np.int32)).cuda()
    else:
        cm_t = cm.to(device="cuda", dtype=torch.int32).contiguous()
    n, m = int(cm_t.shape[0]), int(cm_t.shape[1])
    w_t = None
    if weights is not None:
        w_np = np.ascontiguousarray(weights, dtype=np.float64)
        if w_np.shape != (m, n_states + 1):
            raise ValueError(f"weights table must be [{m}, {n_states + 1}], got {w_np.shape}")
        w_t = torch.from_numpy(w_np).cuda()
    r0, r1 = (0, n) if rows is None else (int(rows[0]), int(rows[1]))
    ld = padded_ld(n)
    tdtype = torch.float64 if dtype == "f64" else torch.float32
    if out is None:
        out = torch.empty((r1 - r0, ld), dtype=tdtype, device="cuda")
        if ld != n:
            out[:, n:].zero_()
    else:
        assert out.is_cuda and out.dtype == tdtype and out.shape[0] >= r1 - r0 and out.stride(0) == ld
    code = _METHOD[method]
    ws_bytes = int(L.cas_whd_workspace_bytes(n, m, int(n_states), 1 if w_t is not None else 0, code))
    ws = torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device="cuda")
    fn_code = PAIR_FUNCTIONS[function]
    if fn_code != 0:
        if method != "exact":
            raise ValueError(f"`{function}` is built on the exact kernel only (method='exact')")
        rc = L.cas_pairwise_map(fn_code, 1 if ignore_missing else 0, cm_t.data_ptr(), n, m, int(missing),
                                w_t.data_ptr() if w_t is not None else None, int(n_states), out.data_ptr(), ld,
                                CAS_F64 if dtype == "f64" else CAS_F32, r0, r1, ws.data_ptr(), ws_bytes,
                                _stream_ptr(torch))
        _lib.check(rc, "cas_pairwise_map")
    else:
        rc = L.cas_whd_map(cm_t.data_ptr(), n, m, int(missing), w_t.data_ptr() if w_t is not None else None,
                           int(n_states), out.data_ptr(), ld, CAS_F64 if dtype == "f64" else CAS_F32,
                           r0, r1, code, ws.data_ptr(), ws_bytes, _stream_ptr(torch))
        _lib.check(rc, "cas_whd_map")
    # first int of the workspace: "state outside [0, n_states]" flag set by the encode kernels
    flag = int(ws[:4].view(torch.int32).item())
    if flag:
        raise ValueError("character matrix holds a state outside [0, n_states] that is not the missing indicator")
    return out


def agglomerate(D, n: int, algo: str, mode: str, max_iters: int = -1):
    """Runs the NJ / UPGMA loop on the device map D ([n, ld] CUDA tensor, destroyed).

    Returns (joins [n-2, 2] int32 numpy in id space, last2 [2] int32 numpy).
    """
    torch = _torch()
    L = _lib.load()
    assert D.is_cuda and D.dim() == 2 and D.stride(1) == 1
    ld = int(D.stride(0))
    dtype = CAS_F64 if D.dtype == torch.float64 else CAS_F32
    if D.dtype not in (torch.float64, torch.float32):
        raise ValueError("map must be float32 or float64")
    njoin = max(int(n) - 2, 0)
    joins = torch.full((njoin + 1, 2), -1, dtype=torch.int32, device="cuda")
    last2 = torch.full((2,), -1, dtype=torch.int32, device="cuda")
    ws_bytes = int(L.cas_agglom_workspace_bytes(int(n), _ALGO[algo], _MODE[mode]))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    fn = L.cas_nj_solve if algo == "nj" else L.cas_upgma_solve
    rc = fn(D.data_ptr(), int(n), ld, dtype, _MODE[mode], int(max_iters), joins.data_ptr(),
            last2.data_ptr(), ws.data_ptr(), ws_bytes, _stream_ptr(torch))
    _lib.check(rc, f"cas_{algo}_solve")
    launches = int(L.cas_last_launch_count())
    near, exact = ctypes.c_int64(0), ctypes.c_int64(0)
    _lib.check(L.cas_agglom_stats(ws.data_ptr(), int(n), _stream_ptr(torch), ctypes.byref(near),
                                  ctypes.byref(exact)), "cas_agglom_stats")
    last_stats["near_ties"], last_stats["exact_ties"] = int(near.value), int(exact.value)
    scanned = ctypes.c_int64(0)
    _lib.check(L.cas_agglom_scanned_elements(ws.data_ptr(), int(n), _stream_ptr(torch), ctypes.byref(scanned)),
               "cas_agglom_scanned_elements")
    last_stats["scanned_elements"] = int(scanned.value)
    listed, drift = ctypes.c_int64(0), ctypes.c_double(0.0)
    _lib.check(L.cas_agglom_prune_diag(ws.data_ptr(), int(n), _stream_ptr(torch), ctypes.byref(listed),
                                       ctypes.byref(drift)), "cas_agglom_prune_diag")
    last_stats["rows_listed"], last_stats["drift"] = int(listed.value), float(drift.value)
    x4 = (ctypes.c_int64 * 4)()
    _lib.check(L.cas_agglom_exact_stats(ws.data_ptr(), int(n), _stream_ptr(torch), x4), "cas_agglom_exact_stats")
    (last_stats["ambiguous_iterations"], last_stats["rows_resolved"], last_stats["max_rows_resolved"],
     last_stats["pairs_changed"]) = (int(v) for v in x4)
    if last_stats["ambiguous_iterations"]:
        rows = min(last_stats["ambiguous_iterations"], 4096)
        log = np.zeros((rows, 4), dtype=np.int32)
        _lib.check(L.cas_agglom_exact_log(ws.data_ptr(), int(n), _stream_ptr(torch), log.ctypes.data, rows),
                   "cas_agglom_exact_log")
        last_stats["ambiguous_log"] = log[:, :3]  # (iteration, live size, rows resolved)
        last_stats["ambiguous_window"] = log[:, 3].copy().view(np.float32)  # the error window W of those iterations
    else:
        last_stats["ambiguous_log"] = np.zeros((0, 3), dtype=np.int32)
    done = njoin if max_iters < 0 else min(njoin, int(max_iters))
    joins_h = joins[:done].cpu().numpy()
    last2_h = last2.cpu().numpy()
    if last2_h[0] == -2:
        raise B200LibraryError("the loop found no finite entry to join (is the dissimilarity map all NaN / inf?)")
    return joins_h, last2_h


def int8_mma_peak_tops(iters: int = 20000) -> float:
    """Measured issue rate of the int8 tensor pipe, TOP/s over the whole GPU (``cas_int8_mma_peak``: tcgen05.mma
    kind::i8 128 x 256 x 32 back to back on shared-memory operands, one CTA per SM)."""
    torch = _torch()
    L = _lib.load()
    out = ctypes.c_double(0.0)
    _lib.check(L.cas_int8_mma_peak(int(iters), _stream_ptr(torch), ctypes.byref(out)), "cas_int8_mma_peak")
    return out.value / 1e12


def loop_launch_floor_us(k: int, iters: int = 2000, kernels_per_iter: int = 3) -> float:
    """Device microseconds per iteration of the pruned loop's bare launch pattern at live size k
    (``cas_loop_launch_floor``: same grids and programmatic dependent launches, kernels that do nothing)."""
    torch = _torch()
    L = _lib.load()
    ws_bytes = int(L.cas_agglom_workspace_bytes(int(k), CAS_ALGO_NJ, CAS_MODE_FAST))
    ws = torch.zeros(ws_bytes, dtype=torch.uint8, device="cuda")
    out = None
    for rep in range(2):  # first pass warms the launch path up
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        _lib.check(L.cas_loop_launch_floor(int(k), int(iters), int(kernels_per_iter), ws.data_ptr(), ws_bytes,
                                           _stream_ptr(torch)), "cas_loop_launch_floor")
        e1.record()
        torch.cuda.synchronize()
        out = 1e3 * e0.elapsed_time(e1) / max(int(iters), 1)
    return out


def smj_solve(cm, missing: int, weights: Optional[np.ndarray], n_states: int,
              function: str = "hamming_similarity_without_missing", max_iters: int = -1):
    """Shared-Mutation-Joining loop (``cas_smj_solve``) on the int32 [n, m] matrix of compact codes `cm`, rows
    as given: the float64 similarity map is built with ``cas_pairwise_map`` and consumed on the device.

    Returns (joins [n-2, 2] int32 numpy in id space, last2 [2]); the reference's last join (its root) is
    the pair `last2`."""
    torch = _torch()
    L = _lib.load()
    if function == "hamming_distance":
        raise ValueError("hamming_distance has no (s1, s2, missing, weights) signature and cannot drive the loop")
    cm_t = (torch.from_numpy(np.ascontiguousarray(cm, dtype=np.int32)).cuda() if isinstance(cm, np.ndarray)
            else cm.to(device="cuda", dtype=torch.int32).contiguous())
    n, m = int(cm_t.shape[0]), int(cm_t.shape[1])
    S_map = whd_map(cm_t, missing, weights, n_states, dtype="f64", method="exact", function=function)
    w_t = None
    if weights is not None:
        w_t = torch.from_numpy(np.ascontiguousarray(weights, dtype=np.float64)).cuda()
    njoin = max(n - 2, 0)
    joins = torch.full((njoin + 1, 2), -1, dtype=torch.int32, device="cuda")
    last2 = torch.full((2,), -1, dtype=torch.int32, device="cuda")
    ws_bytes = int(L.cas_smj_workspace_bytes(n, m))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    rc = L.cas_smj_solve(cm_t.data_ptr(), n, m, int(missing), w_t.data_ptr() if w_t is not None else None,
                         int(n_states), PAIR_FUNCTIONS[function], S_map.data_ptr(), int(S_map.stride(0)),
                         int(max_iters), joins.data_ptr(), last2.data_ptr(), ws.data_ptr(), ws_bytes,
                         _stream_ptr(torch))
    _lib.check(rc, "cas_smj_solve")
    done = njoin if max_iters < 0 else min(njoin, int(max_iters))
    return joins[:done].cpu().numpy(), last2.cpu().numpy()


def solve_host(cm: np.ndarray, missing: int, weights: Optional[np.ndarray], n_states: int,
               algo: str, mode: str, method: str = "exact"):
    """End to end from HOST buffers through ``cas_solve_host`` (H2D/D2H inside the call).

    Returns (joins, last2, timings_ms[h2d, map, loop, d2h]).
    """
    _torch()
    L = _lib.load()
    cm = np.ascontiguousarray(cm, dtype=np.int32)
    n, m = cm.shape
    w = None
    if weights is not None:
        w = np.ascontiguousarray(weights, dtype=np.float64)
        assert w.shape == (m, n_states + 1)
    joins = np.full((max(n - 2, 0) + 1, 2), -1, dtype=np.int32)
    last2 = np.full(2, -1, dtype=np.int32)
    tm = np.zeros(4, dtype=np.float64)
    rc = L.cas_solve_host(cm.ctypes.data, n, m, int(missing), w.ctypes.data if w is not None else None,
                          int(n_states), _ALGO[algo], _MODE[mode], _METHOD[method],
                          joins.ctypes.data, last2.ctypes.data, tm.ctypes.data)
    _lib.check(rc, "cas_solve_host")
    return joins[: max(n - 2, 0)], last2, tm


def solve_host_release() -> None:
    """Frees the device buffers ``cas_solve_host`` keeps between calls."""
    _lib.check(_lib.load().cas_solve_host_release(), "cas_solve_host_release")


def nj_q(D: np.ndarray) -> np.ndarray:
    """Q-criterion matrix of a float64 [n, n] map (single-step hook, ``cas_nj_q``)."""
    torch = _torch()
    L = _lib.load()
    D = np.ascontiguousarray(D, dtype=np.float64)
    n = D.shape[0]
    d = torch.from_numpy(D).cuda()
    q = torch.empty((n, n), dtype=torch.float64, device="cuda")
    rs = torch.empty(n, dtype=torch.float64, device="cuda")
    _lib.check(L.cas_nj_q(d.data_ptr(), n, n, q.data_ptr(), rs.data_ptr(), _stream_ptr(torch)), "cas_nj_q")
    return q.cpu().numpy()


def update_map(D: np.ndarray, i: int, j: int, algo: str, size_i: int = 1, size_j: int = 1) -> np.ndarray:
    """(n-1) x (n-1) map after joining rows i and j (single-step hook, ``cas_update_map``)."""
    torch = _torch()
    L = _lib.load()
    D = np.ascontiguousarray(D, dtype=np.float64)
    n = D.shape[0]
    d = torch.from_numpy(D).cuda()
    out = torch.empty((n - 1, n - 1), dtype=torch.float64, device="cuda")
    _lib.check(L.cas_update_map(d.data_ptr(), n, n, _ALGO[algo], int(i), int(j), int(size_i), int(size_j),
                                out.data_ptr(), n - 1, _stream_ptr(torch)), "cas_update_map")
    return out.cpu().numpy()
