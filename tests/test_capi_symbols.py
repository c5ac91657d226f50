"""CPU: the C-ABI library loads and exports every symbol include/sci_b200.h declares,
and the ctypes binding table covers exactly that set (no compute calls here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "sci_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sci_[A-Za-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    lib = ctypes.CDLL(os.path.join(ROOT, "adaptivepnp_sci_b200", "libsci_b200.so"))
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "declared in sci_b200.h but not exported: " + n


def test_binding_table_matches_header():
    from adaptivepnp_sci_b200 import _lib
    bound = set(_lib.PROTOTYPES) | set(_lib._SPECIAL_RESTYPE)
    assert bound == set(_declared())


def test_version_and_error_string():
    from adaptivepnp_sci_b200 import _lib
    assert _lib.lib.sci_version() >= 1000
    # argument validation happens before any CUDA call, so this is safe without a GPU
    rc = _lib.lib.sci_pixlast_to_planar(None, None, 0, 0, 0, None)
    assert rc == -1 and b"invalid argument" in _lib.lib.sci_last_error()
