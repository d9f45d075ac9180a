"""The C-ABI library loads on a CPU-only box and exports every symbol include/cas_b200.h declares."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "cas_b200.h")).read()
    return sorted(set(re.findall(r"CAS_API[^;]*?\b(cas_[a-z0-9_]+)\s*\(", text, flags=re.S)))


def test_header_declares_the_expected_entry_points():
    names = _declared()
    for must in ("cas_whd_map", "cas_nj_solve", "cas_upgma_solve", "cas_solve_host", "cas_last_error"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from cassiopeia_b200 import _lib

    assert os.path.exists(_lib.LIB_PATH), "build with __graft_entry__.build() first"
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(L, name), f"{name} declared in cas_b200.h but not exported"
    assert set(_lib.EXPORTS) <= set(_declared())
    L.cas_abi_version.restype = ctypes.c_int
    assert L.cas_abi_version() == _lib.CAS_ABI_VERSION


def test_binding_loads_without_a_gpu():
    from cassiopeia_b200 import _lib

    L = _lib.load()
    assert L.cas_last_error() is not None
