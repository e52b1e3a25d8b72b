"""CPU: the C-ABI shared library loads and exports every symbol include/d3feat_b200.h declares, and the
Python binding table covers exactly that set (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "d3feat_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(d3f_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built_lib():
    from d3feat_b200 import build
    path = build.build()
    return ctypes.CDLL(path)


def test_header_declares_the_expected_entry_points():
    syms = header_symbols()
    for must in ("d3f_grid_subsample", "d3f_radius_neighbors_build", "d3f_radius_neighbors_count",
                 "d3f_radius_neighbors_fill", "d3f_kpconv_forward", "d3f_kpconv_deform_forward",
                 "d3f_unary_forward", "d3f_ind_max_pool", "d3f_closest_pool", "d3f_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol(built_lib):
    for s in header_symbols():
        assert hasattr(built_lib, s), "libd3feat_b200.so does not export %s" % s


def test_binding_table_matches_header(built_lib):
    from d3feat_b200 import _lib
    bound = sorted(n for n, _, _ in _lib.SYMBOLS)
    assert bound == header_symbols()
    _lib.lib()          # binds argtypes / restypes for all of them
    assert _lib.lib().d3f_version() >= 100
    assert _lib.launch_count() == 0


def test_invalid_arguments_return_codes_without_a_gpu(built_lib):
    """Argument validation happens before any CUDA call, so it can be exercised on the CPU box."""
    built_lib.d3f_last_error.restype = ctypes.c_char_p
    built_lib.d3f_kpconv_forward.restype = ctypes.c_int
    rc = built_lib.d3f_unary_forward(None, None, None, ctypes.c_int(-1), ctypes.c_int(4), ctypes.c_int(4), None, None,
                                     None, None, ctypes.c_float(-1.0), None, None, None)
    assert rc == -1
    assert b"bad shape" in built_lib.d3f_last_error()
    built_lib.d3f_radius_neighbors_workspace_bytes.restype = ctypes.c_size_t
    assert built_lib.d3f_radius_neighbors_workspace_bytes(ctypes.c_int(10), ctypes.c_int(1), ctypes.c_float(0.1),
                                                          None) == 0


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from d3feat_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.D3FError):
        _lib.lib()
