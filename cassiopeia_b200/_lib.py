"""ctypes binding of ``libcas_b200.so`` (the C ABI declared in ``include/cas_b200.h``).

There is no CPU fallback: if the shared library is missing or cannot be loaded
every compute entry point raises :class:`B200LibraryError`.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CAS_B200_LIBRARY", os.path.join(_HERE, "libcas_b200.so"))

# constants of include/cas_b200.h
CAS_ABI_VERSION = 2
CAS_F32, CAS_F64 = 0, 1
CAS_ALGO_NJ, CAS_ALGO_UPGMA = 0, 1
CAS_MODE_STRICT, CAS_MODE_FAST, CAS_MODE_EXACT = 0, 1, 2
CAS_MAP_EXACT, CAS_MAP_TENSOR = 0, 1

EXPORTS = (
    "cas_abi_version",
    "cas_last_error",
    "cas_device_info",
    "cas_whd_workspace_bytes",
    "cas_whd_map",
    "cas_pairwise_map",
    "cas_smj_workspace_bytes",
    "cas_smj_solve",
    "cas_agglom_workspace_bytes",
    "cas_nj_solve",
    "cas_upgma_solve",
    "cas_last_launch_count",
    "cas_solve_host",
    "cas_solve_host_release",
    "cas_nj_q",
    "cas_update_map",
    "cas_agglom_stats",
    "cas_agglom_scanned_elements",
    "cas_agglom_prune_diag",
    "cas_agglom_exact_stats",
    "cas_agglom_exact_log",
    "cas_loop_launch_floor",
    "cas_int8_mma_peak",
    "cas_sharded_local_rows",
    "cas_sharded_local_capacity",
    "cas_sharded_workspace_bytes",
    "cas_sharded_scanned_elements",
    "cas_whd_map_cyclic",
    "cas_nccl_unique_id",
    "cas_nccl_comm_create",
    "cas_nccl_comm_destroy",
    "cas_sharded_solve",
    "cas_sharded_solve_emulated",
    "cas_exchange_bytes",
    "cas_exchange_alloc",
    "cas_exchange_open",
    "cas_exchange_close",
    "cas_exchange_free",
    "cas_exchange_error",
    "cas_sharded_solve_p2p",
    "cas_sharded_solve_emulated_p2p",
    "cas_sharded_exact_stats",
    "cas_tail_preorder",
    "cas_tail_heights",
    "cas_tail_successor_lists",
)


class B200LibraryError(RuntimeError):
    """libcas_b200.so is missing / failed to load / returned an error code."""


_lib = None
_lock = threading.Lock()


def _declare(L: ctypes.CDLL) -> None:
    i32, i64, sz, vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t, ctypes.c_void_p
    L.cas_abi_version.restype = ctypes.c_int
    L.cas_abi_version.argtypes = []
    L.cas_last_error.restype = ctypes.c_char_p
    L.cas_last_error.argtypes = []
    L.cas_last_launch_count.restype = i64
    L.cas_last_launch_count.argtypes = []
    L.cas_device_info.restype = ctypes.c_int
    L.cas_device_info.argtypes = [ctypes.POINTER(i32)] * 3 + [ctypes.POINTER(sz)] * 2
    L.cas_whd_workspace_bytes.restype = sz
    L.cas_whd_workspace_bytes.argtypes = [i64, i32, i32, i32, i32]
    L.cas_whd_map.restype = ctypes.c_int
    L.cas_whd_map.argtypes = [vp, i64, i32, i32, vp, i32, vp, i64, i32, i64, i64, i32, vp, sz, vp]
    L.cas_pairwise_map.restype = ctypes.c_int
    L.cas_pairwise_map.argtypes = [i32, i32, vp, i64, i32, i32, vp, i32, vp, i64, i32, i64, i64, vp, sz, vp]
    L.cas_smj_workspace_bytes.restype = sz
    L.cas_smj_workspace_bytes.argtypes = [i64, i32]
    L.cas_smj_solve.restype = ctypes.c_int
    L.cas_smj_solve.argtypes = [vp, i64, i32, i32, vp, i32, i32, vp, i64, i64, vp, vp, vp, sz, vp]
    L.cas_agglom_workspace_bytes.restype = sz
    L.cas_agglom_workspace_bytes.argtypes = [i64, i32, i32]
    for fn in (L.cas_nj_solve, L.cas_upgma_solve):
        fn.restype = ctypes.c_int
        fn.argtypes = [vp, i64, i64, i32, i32, i64, vp, vp, vp, sz, vp]
    L.cas_agglom_stats.restype = ctypes.c_int
    L.cas_agglom_stats.argtypes = [vp, i64, vp, ctypes.POINTER(i64), ctypes.POINTER(i64)]
    L.cas_agglom_scanned_elements.restype = ctypes.c_int
    L.cas_agglom_scanned_elements.argtypes = [vp, i64, vp, ctypes.POINTER(i64)]
    L.cas_agglom_prune_diag.restype = ctypes.c_int
    L.cas_agglom_prune_diag.argtypes = [vp, i64, vp, ctypes.POINTER(i64), ctypes.POINTER(ctypes.c_double)]
    L.cas_agglom_exact_stats.restype = ctypes.c_int
    L.cas_agglom_exact_stats.argtypes = [vp, i64, vp, vp]
    L.cas_agglom_exact_log.restype = ctypes.c_int
    L.cas_agglom_exact_log.argtypes = [vp, i64, vp, vp, i64]
    L.cas_loop_launch_floor.restype = ctypes.c_int
    L.cas_loop_launch_floor.argtypes = [i64, i64, i32, vp, sz, vp]
    L.cas_int8_mma_peak.restype = ctypes.c_int
    L.cas_int8_mma_peak.argtypes = [i32, vp, vp]
    L.cas_tail_preorder.restype = i64
    L.cas_tail_preorder.argtypes = [vp, vp, i64, i64, vp, vp]
    L.cas_tail_heights.restype = ctypes.c_int
    L.cas_tail_heights.argtypes = [vp, i64, vp, vp, vp]
    L.cas_tail_successor_lists.restype = ctypes.c_int
    L.cas_tail_successor_lists.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp]
    L.cas_nj_q.restype = ctypes.c_int
    L.cas_nj_q.argtypes = [vp, i64, i64, vp, vp, vp]
    L.cas_update_map.restype = ctypes.c_int
    L.cas_update_map.argtypes = [vp, i64, i64, i32, i64, i64, i64, i64, vp, i64, vp]
    L.cas_sharded_scanned_elements.restype = ctypes.c_int
    L.cas_sharded_scanned_elements.argtypes = [vp, i64, i32, i32, vp, ctypes.POINTER(i64)]
    L.cas_sharded_local_rows.restype = i64
    L.cas_sharded_local_rows.argtypes = [i64, i32, i32]
    L.cas_sharded_local_capacity.restype = i64
    L.cas_sharded_local_capacity.argtypes = [i64, i32]
    L.cas_sharded_workspace_bytes.restype = sz
    L.cas_sharded_workspace_bytes.argtypes = [i64, i32, i32]
    L.cas_whd_map_cyclic.restype = ctypes.c_int
    L.cas_whd_map_cyclic.argtypes = [vp, i64, i32, i32, vp, i32, vp, i64, i32, i32, i32, i32, vp, sz, vp]
    L.cas_nccl_unique_id.restype = ctypes.c_int
    L.cas_nccl_unique_id.argtypes = [vp]
    L.cas_nccl_comm_create.restype = ctypes.c_int
    L.cas_nccl_comm_create.argtypes = [vp, i32, i32, ctypes.POINTER(vp)]
    L.cas_nccl_comm_destroy.restype = ctypes.c_int
    L.cas_nccl_comm_destroy.argtypes = [vp]
    L.cas_sharded_solve.restype = ctypes.c_int
    L.cas_sharded_solve.argtypes = [vp, i64, i64, i32, i32, i64, i32, i32, vp, vp, vp, vp, sz, vp]
    L.cas_sharded_solve_emulated.restype = ctypes.c_int
    L.cas_sharded_solve_emulated.argtypes = [vp, i64, i64, i32, i32, i64, i32, vp, vp, vp, sz, vp]
    u64 = ctypes.c_uint64
    L.cas_exchange_bytes.restype = sz
    L.cas_exchange_bytes.argtypes = [i64, i32, i32]
    L.cas_exchange_alloc.restype = ctypes.c_int
    L.cas_exchange_alloc.argtypes = [sz, ctypes.POINTER(vp), vp]
    L.cas_exchange_open.restype = ctypes.c_int
    L.cas_exchange_open.argtypes = [vp, ctypes.POINTER(vp)]
    for fn in (L.cas_exchange_close, L.cas_exchange_free):
        fn.restype = ctypes.c_int
        fn.argtypes = [vp]
    L.cas_exchange_error.restype = ctypes.c_int
    L.cas_exchange_error.argtypes = [vp, vp, ctypes.POINTER(i32)]
    L.cas_sharded_solve_p2p.restype = ctypes.c_int
    L.cas_sharded_solve_p2p.argtypes = [vp, i64, i64, i32, i32, i32, i64, i32, i32, vp, u64, vp, vp, vp, sz, vp]
    L.cas_sharded_solve_emulated_p2p.restype = ctypes.c_int
    L.cas_sharded_solve_emulated_p2p.argtypes = [vp, i64, i64, i32, i32, i32, i64, i32, vp, u64, vp, vp, vp, sz, vp]
    L.cas_sharded_exact_stats.restype = ctypes.c_int
    L.cas_sharded_exact_stats.argtypes = [vp, i64, i32, i32, vp, vp]
    L.cas_solve_host_release.restype = ctypes.c_int
    L.cas_solve_host_release.argtypes = []
    L.cas_solve_host.restype = ctypes.c_int
    L.cas_solve_host.argtypes = [vp, i64, i32, i32, vp, i32, i32, i32, i32, vp, vp, vp]


def load() -> ctypes.CDLL:
    """Loads the library once; raises B200LibraryError when it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise B200LibraryError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C cassiopeia_b200/csrc` (there is no CPU fallback)"
            )
        try:
            L = ctypes.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise B200LibraryError(f"cannot load {LIB_PATH}: {e}") from e
        missing = [s for s in EXPORTS if not hasattr(L, s)]
        if missing:
            raise B200LibraryError(f"{LIB_PATH} lacks symbols {missing}")
        _declare(L)
        if L.cas_abi_version() != CAS_ABI_VERSION:
            raise B200LibraryError(
                f"ABI mismatch: library {L.cas_abi_version()} vs binding {CAS_ABI_VERSION}"
            )
        _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().cas_last_error().decode("utf-8", "replace")
        raise B200LibraryError(f"{what} failed (code {rc}): {msg}")
