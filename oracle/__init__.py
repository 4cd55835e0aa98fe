"""oracle/ -- CPU checkers for the hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product
(``superpoints_registration_b200``) never imports it and has no CPU fallback.

Two libraries sit behind it:

* ``oracle/_build/libspr_oracle.so`` -- our plain-C restatement (``spr_oracle.c``), kind ``"port"``.
* ``oracle/_ref/libspr_ref.so``      -- the reference's own C++ core (cloud.cpp, neighbors.cpp,
  grid_subsampling.cpp) compiled in place from ``/root/reference`` behind ``ref_shim.cpp``,
  kind ``"reference"``.  It is built in the authoring container (the GPU box has no reference tree)
  and travels as a git-ignored binary.

Parity status: pinned (see ``tests/test_oracle_vs_ref.py`` and ``tests/golden/``).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import List, Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PORT_SO = os.path.join(_HERE, "_build", "libspr_oracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libspr_ref.so")

_port = None
_ref = None


def build(force: bool = False) -> None:
    """Compile the C restatement, and the in-place reference build when /root/reference exists."""
    if force or not os.path.exists(_PORT_SO) or os.path.getmtime(_PORT_SO) < os.path.getmtime(
            os.path.join(_HERE, "spr_oracle.c")):
        subprocess.run(["make", "-s", "-C", _HERE, "port"], check=True)
    ref_root = "/root/reference/src/models/backbone_kpconv/cpp_wrappers"
    if os.path.isdir(ref_root) and (force or not os.path.exists(_REF_SO)):
        subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def _ptr(a: Optional[np.ndarray], ctype):
    if a is None:
        return ctypes.cast(None, ctypes.POINTER(ctype))
    return a.ctypes.data_as(ctypes.POINTER(ctype))


def port_lib():
    global _port
    if _port is None:
        build()
        lib = ctypes.CDLL(_PORT_SO)
        lib.orc_grid_subsample_batch.restype = ctypes.c_int
        lib.orc_radius_neighbors_batch.restype = ctypes.c_int
        lib.orc_kpconv_forward.restype = ctypes.c_int
        _port = lib
    return _port


def have_ref() -> bool:
    return os.path.exists(_REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        if not have_ref():
            build()
        lib = ctypes.CDLL(_REF_SO)
        lib.ref_batch_neighbors.restype = ctypes.c_int
        lib.ref_batch_grid_subsampling.restype = ctypes.c_int
        _ref = lib
    return _ref


def set_threads(n: int) -> None:
    port_lib().orc_set_threads(ctypes.c_int(n))


# ------------------------------------------------------------------------------------------------
# C restatement ("port")
# ------------------------------------------------------------------------------------------------

def grid_subsample_batch(points, lengths, dl: float, max_p: int = 0, return_keys: bool = False):
    """Canonical-order barycentre grid subsampling. -> (points f32[M,3], lengths i32[B][, keys u64[M]])"""
    pts, lens = _f32(points).reshape(-1, 3), _i32(lengths)
    n = pts.shape[0]
    assert int(lens.sum()) == n
    out = np.empty((max(n, 1), 3), np.float32)
    out_l = np.empty(lens.shape[0], np.int32)
    keys = np.empty(max(n, 1), np.uint64)
    m = port_lib().orc_grid_subsample_batch(_ptr(pts, ctypes.c_float), _ptr(lens, ctypes.c_int), ctypes.c_int(len(lens)),
                                            ctypes.c_float(dl), ctypes.c_int(max_p), _ptr(out, ctypes.c_float),
                                            _ptr(out_l, ctypes.c_int), _ptr(keys, ctypes.c_uint64))
    if m < 0:
        raise RuntimeError("orc_grid_subsample_batch failed")
    if return_keys:
        return out[:m].copy(), out_l, keys[:m].copy()
    return out[:m].copy(), out_l


def radius_neighbors_batch(queries, supports, q_lengths, s_lengths, radius: float, limit: int, details: bool = False):
    """-> idx i32[Nq, limit] padded with Ns (already truncated to `limit` columns), max_count
    (, d2 f32[Nq,limit], cut f32[Nq], count i32[Nq] when details)."""
    q, s = _f32(queries).reshape(-1, 3), _f32(supports).reshape(-1, 3)
    ql, sl = _i32(q_lengths), _i32(s_lengths)
    nq, ns = q.shape[0], s.shape[0]
    idx = np.empty((nq, limit), np.int32)
    d2 = np.empty((nq, limit), np.float32) if details else None
    cut = np.empty(nq, np.float32) if details else None
    cnt = np.empty(nq, np.int32) if details else None
    mc = port_lib().orc_radius_neighbors_batch(
        _ptr(q, ctypes.c_float), ctypes.c_int(nq), _ptr(s, ctypes.c_float), ctypes.c_int(ns), _ptr(ql, ctypes.c_int),
        _ptr(sl, ctypes.c_int), ctypes.c_int(len(ql)), ctypes.c_float(radius), ctypes.c_int(limit),
        _ptr(idx, ctypes.c_int32), _ptr(d2, ctypes.c_float), _ptr(cut, ctypes.c_float), _ptr(cnt, ctypes.c_int32))
    if mc < 0:
        raise RuntimeError(f"orc_radius_neighbors_batch failed ({mc})")
    if details:
        return idx, mc, d2, cut, cnt
    return idx, mc


def kpconv_forward(q_pts, s_pts, neighb_inds, x, weights, kernel_points, extent: float, f32_acc: bool = False):
    q, s = _f32(q_pts).reshape(-1, 3), _f32(s_pts).reshape(-1, 3)
    idx = np.ascontiguousarray(neighb_inds, dtype=np.int64)
    xx, w, kp = _f32(x), _f32(weights), _f32(kernel_points)
    K, cin, cout = w.shape
    out = np.empty((q.shape[0], cout), np.float32)
    rc = port_lib().orc_kpconv_forward(
        _ptr(q, ctypes.c_float), ctypes.c_int(q.shape[0]), _ptr(s, ctypes.c_float), ctypes.c_int(s.shape[0]),
        _ptr(idx, ctypes.c_int64), ctypes.c_int(idx.shape[1]), _ptr(xx, ctypes.c_float), ctypes.c_int(cin),
        _ptr(w, ctypes.c_float), ctypes.c_int(cout), _ptr(kp, ctypes.c_float), ctypes.c_int(K), ctypes.c_float(extent),
        ctypes.c_int(1 if f32_acc else 0), _ptr(out, ctypes.c_float))
    if rc != 0:
        raise RuntimeError(f"orc_kpconv_forward failed ({rc})")
    return out


# ------------------------------------------------------------------------------------------------
# the reference's own C++ ("reference"), through oracle/_ref
# ------------------------------------------------------------------------------------------------

def ref_batch_query(queries, supports, q_batches, s_batches, radius: float, brute: bool = False) -> np.ndarray:
    """What cpp_neighbors.batch_query returns (cpp_neighbors/wrapper.cpp:58-238): i32[Nq, max_count]."""
    q, s = _f32(queries).reshape(-1, 3), _f32(supports).reshape(-1, 3)
    ql, sl = _i32(q_batches), _i32(s_batches)
    lib = ref_lib()
    mc = lib.ref_batch_neighbors(_ptr(q, ctypes.c_float), ctypes.c_int(q.shape[0]), _ptr(s, ctypes.c_float),
                                 ctypes.c_int(s.shape[0]), _ptr(ql, ctypes.c_int), _ptr(sl, ctypes.c_int),
                                 ctypes.c_int(len(ql)), ctypes.c_float(radius), ctypes.c_int(1 if brute else 0))
    out = np.empty((q.shape[0], mc), np.int32)
    lib.ref_neighbors_fetch(_ptr(out, ctypes.c_int))
    return out


def ref_subsample_batch(points, batches, sampleDl: float = 0.1, max_p: int = 0):
    """What cpp_subsampling.subsample_batch returns (cpp_subsampling/wrapper.cpp:62-333) without
    features/classes: (f32[M,3], i32[B]) in the reference's unordered_map order."""
    pts, lens = _f32(points).reshape(-1, 3), _i32(batches)
    lib = ref_lib()
    m = lib.ref_batch_grid_subsampling(_ptr(pts, ctypes.c_float), ctypes.c_int(pts.shape[0]), _ptr(lens, ctypes.c_int),
                                       ctypes.c_int(len(lens)), ctypes.c_float(sampleDl), ctypes.c_int(max_p))
    out = np.empty((m, 3), np.float32)
    out_l = np.empty(len(lens), np.int32)
    lib.ref_subsample_fetch(_ptr(out, ctypes.c_float), _ptr(out_l, ctypes.c_int))
    return out, out_l


class RefNeighborsModule:
    """Duck-typed stand-in for the reference's `radius_neighbors` extension (kpconv.py:258)."""

    @staticmethod
    def batch_query(queries, supports, q_batches, s_batches, radius=0.1):
        return ref_batch_query(np.asarray(queries), np.asarray(supports), np.asarray(q_batches),
                               np.asarray(s_batches), float(radius))


class RefSubsamplingModule:
    """Duck-typed stand-in for the reference's `grid_subsampling` extension (kpconv.py:180)."""

    @staticmethod
    def subsample_batch(points, batches, sampleDl=0.1, max_p=0, verbose=0):
        return ref_subsample_batch(np.asarray(points), np.asarray(batches), float(sampleDl), int(max_p))
