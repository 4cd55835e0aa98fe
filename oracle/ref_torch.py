"""oracle/ref_torch.py -- import the reference's *unmodified* PyTorch modules.  TEST INFRASTRUCTURE ONLY.

Works only where ``/root/reference`` exists (the authoring container).  It is used by
``tests/golden/make_golden.py`` to generate the committed golden vectors and by the
``needs_reference`` tests; nothing on the GPU box touches it.

Recipe:
  1. namespace-style stubs for the ``models`` packages so that ``models/__init__.py`` (which eagerly
     imports every model and its un-installed dependencies) never runs;
  2. empty stub modules for MinkowskiEngine, pytorch3d(.ops), nibabel, coloredlogs, easydict, h5py ...;
  3. ``kpconv.cpp_neighbors`` / ``kpconv.cpp_subsampling`` (commented-out imports at kpconv.py:12-15) are
     injected with the reference's own C++ core through ``oracle/_ref`` (``RefNeighborsModule`` ...);
  4. the kernel-point ``.ply`` is resolved relative to cwd (kernel_points.py:390), so module
     construction happens with cwd = /root/reference/src.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types

REF_SRC = "/root/reference/src"


def available() -> bool:
    return os.path.isdir(REF_SRC)


class AttrDict(dict):
    """dict with attribute access -- what the reference gets from EasyDict (train.py / test.py)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _stub(name: str, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _pkg_stub(name: str, path: str):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__path__ = [path]
    sys.modules[name] = m
    return m


_loaded = {}


def load():
    """Returns a namespace with the reference modules: kpconv, kpconv_blocks, se3_torch, regtr, misc."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError("reference tree not present")
    import oracle

    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    _pkg_stub("models", os.path.join(REF_SRC, "models"))
    for sub in ("backbone_kpconv", "losses", "transformer", "scheduler"):
        _pkg_stub(f"models.{sub}", os.path.join(REF_SRC, "models", sub))
    _pkg_stub("models.backbone_kpconv.kernels", os.path.join(REF_SRC, "models", "backbone_kpconv", "kernels"))
    me = _stub("MinkowskiEngine")
    me.utils = types.SimpleNamespace()
    _stub("pytorch3d")
    _stub("pytorch3d.ops", packed_to_padded=None, ball_query=None, knn_points=None)
    _stub("pytorch3d.ops.knn", knn_gather=None, knn_points=None)
    _stub("pytorch3d.loss", chamfer_distance=None)
    nib = _stub("nibabel")
    nib.quaternions = _stub("nibabel.quaternions")
    _stub("coloredlogs")
    _stub("easydict", EasyDict=AttrDict)
    _stub("h5py")
    _stub("open3d")
    _stub("git")
    _stub("tensorboardX")

    kpconv = importlib.import_module("models.backbone_kpconv.kpconv")
    kpconv_blocks = importlib.import_module("models.backbone_kpconv.kpconv_blocks")
    kpconv.cpp_neighbors = oracle.RefNeighborsModule
    kpconv.cpp_subsampling = oracle.RefSubsamplingModule
    se3_torch = importlib.import_module("utils.se3_torch")
    misc = importlib.import_module("utils.misc")
    seq = importlib.import_module("utils.seq_manipulation")
    regtr = importlib.import_module("models.qk_regtr_full")
    _loaded.update(kpconv=kpconv, kpconv_blocks=kpconv_blocks, se3_torch=se3_torch, misc=misc, regtr=regtr, seq=seq)
    return types.SimpleNamespace(**_loaded)


@contextlib.contextmanager
def ref_cwd():
    old = os.getcwd()
    os.chdir(REF_SRC)
    try:
        yield
    finally:
        os.chdir(old)


def load_cfg(name: str, **overrides) -> AttrDict:
    """Flat-merged yaml exactly as utils/misc.py:10-29 does it, with attribute access."""
    ref = load()
    cfg = AttrDict(ref.misc.load_config(os.path.join(REF_SRC, "conf", name)))
    cfg.update(overrides)
    return cfg


def build_model(cfg, seed: int = 0):
    """Instantiate the reference RegTR with the CPU Preprocessor (the oracle north_star names)."""
    import numpy as np
    import torch

    ref = load()
    torch.manual_seed(seed)
    np.random.seed(seed)
    with ref_cwd():
        model = ref.regtr.RegTR(cfg)
    model.preprocessor = ref.kpconv.Preprocessor(cfg)
    model.eval()
    return model
