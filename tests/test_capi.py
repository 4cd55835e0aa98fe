"""The C-ABI library loads without a GPU and exports every symbol include/spr_b200.h declares."""
import ctypes
import os
import re

import pytest

from superpoints_registration_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "spr_b200.h")).read()
    return sorted(set(re.findall(r"SPR_API[^;(]*?\b(spr_\w+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    syms = _declared_symbols()
    for must in ("spr_grid_subsample_batch", "spr_cell_grid_build", "spr_radius_query", "spr_kpconv_forward",
                 "spr_dual_softmax_match", "spr_weighted_procrustes", "spr_sinkhorn_weighted_targets"):
        assert must in syms
    assert len(syms) >= 15


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        pytest.fail(f"{_lib.LIB_PATH} not built: run __graft_entry__.build()")
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared_symbols():
        assert hasattr(handle, name), f"{name} declared in include/spr_b200.h but not exported"
    # and the Python binding covers the same set
    assert sorted(_lib.EXPORTED_SYMBOLS) == _declared_symbols()


def _declared_arity():
    """name -> number of parameters of every prototype in include/spr_b200.h."""
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "spr_b200.h")).read(), flags=re.S)
    out = {}
    for name, params in re.findall(r"SPR_API[^;(]*?\b(spr_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        params = params.strip()
        out[name] = 0 if params in ("", "void") else params.count(",") + 1
    return out


def test_binding_argument_counts_match_the_header():
    """A prototype that gains a parameter must gain it in the ctypes table too (a mismatch would pass garbage)."""
    arity = _declared_arity()
    assert sorted(arity) == _declared_symbols()
    for name, (_restype, argtypes) in _lib._SIGNATURES.items():
        assert len(argtypes) == arity[name], (name, len(argtypes), arity[name])


def test_binding_loads_and_reports_version():
    L = _lib.lib()
    assert L.spr_version() >= 100
    assert isinstance(_lib.launch_count(), int)


def test_workspace_queries_need_no_gpu():
    L = _lib.lib()
    assert L.spr_grid_subsample_workspace_bytes(1000, 2) > 1000 * 8
    assert L.spr_cell_grid_workspace_bytes(1000, 2) > 1000 * 16
    assert L.spr_kpconv_workspace_bytes(10, 1000, 32, 32, 15) >= 1001
    assert L.spr_instance_norm_workspace_bytes(1000, 2, 64) > 0
    assert L.spr_match_workspace_bytes(100, 100, 1) > 0
    assert L.spr_sinkhorn_workspace_bytes(100, 100, 1) > 0


def test_argument_errors_are_reported_without_touching_the_device():
    """Validation happens before any CUDA call, so it can be exercised on a CPU-only machine."""
    L = _lib.lib()
    rc = L.spr_grid_subsample_batch(None, None, 0, 0, 0.1, None, None, None, None, 0, None)
    assert rc == -1 and b"empty input" in L.spr_last_error()
    rc = L.spr_radius_query(None, None, 10, 1, None, 10, 0.1, 500, None, 1, 500, None, None)
    assert rc == -1 and b"limit" in L.spr_last_error()
    rc = L.spr_kpconv_forward(None, None, None, 1, 4, 4, None, 32, None, 32, None, 15, 0.1, None, 0, 10, 0, None, 0, None, None)
    assert rc == -1
    with pytest.raises(RuntimeError):
        _lib.check(rc, "spr_kpconv_forward")
