"""The C-ABI library loads and exports every symbol include/pointrix_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pointrix_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pxb_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from pointrix_b200 import _lib

    names = _declared()
    assert len(names) >= 19
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_binding_table_matches_header():
    from pointrix_b200 import _lib

    assert sorted(_lib.SIGNATURES) == _declared()


def test_argument_counts_match_header():
    from pointrix_b200 import _lib

    src = open(os.path.join(ROOT, "include", "pointrix_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for name, params in re.findall(r"\b(pxb_\w+)\s*\(([^)]*)\)\s*;", src):
        n = 0 if params.strip() in ("", "void") else params.count(",") + 1
        assert n == len(_lib.SIGNATURES[name][1]), name


def test_pure_host_entry_points():
    """Entry points that never touch the device can be called without a GPU."""
    from pointrix_b200 import _lib

    assert [_lib.lib.pxb_record_stride(c) for c in (1, 2, 3, 6, 7, 10, 18, 26, 27)] == [8, 8, 12, 12, 16, 16, 24, 32, -1]
    a = _lib.lib.pxb_bin_sort_workspace_bytes(10000, 1920, 1080)
    b = _lib.lib.pxb_bin_sort_workspace_bytes(20000, 1920, 1080)
    assert 0 < a < b
    assert 0 < _lib.lib.pxb_bin_prepare_workspace_bytes(1000) < _lib.lib.pxb_bin_prepare_workspace_bytes(100000)
