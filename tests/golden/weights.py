"""Deterministic weight filler shared by tests/golden/make_golden.py (which writes them INTO the reference
model before recording its outputs) and by the tests (which write the same values into our modules).
Keeps multi-megabyte state_dicts out of the repository: a fixture only stores a seed.

Kernel points are never generated here -- they are copied from the reference module and stored in the fixture
(the reference perturbs them with the global NumPy RNG at construction time).
"""
from __future__ import annotations

from typing import Dict, Mapping, Tuple

import numpy as np


def filled_state(shapes: Mapping[str, Tuple[int, ...]], seed: int) -> Dict[str, np.ndarray]:
    """name -> float32 array, generated in sorted-name order from one PCG64 stream."""
    rng = np.random.default_rng(seed)
    out = {}
    for name in sorted(shapes):
        shape = tuple(shapes[name])
        if name.endswith("kernel_points"):
            continue
        if len(shape) == 0:
            out[name] = np.asarray(rng.uniform(0.5, 1.5), np.float32)
        elif len(shape) == 1:
            if ".norm" in name and name.endswith("weight"):
                out[name] = (1.0 + 0.1 * rng.standard_normal(shape)).astype(np.float32)
            else:
                out[name] = (0.05 * rng.standard_normal(shape)).astype(np.float32)
        else:
            fan_in = int(np.prod(shape[:-1])) if name.endswith("KPConv.weights") else int(np.prod(shape[1:]))
            bound = (3.0 / max(fan_in, 1)) ** 0.5
            out[name] = rng.uniform(-bound, bound, size=shape).astype(np.float32)
    return out


def reference_shapes(own_shapes: Mapping[str, Tuple[int, ...]], d_embed: int) -> Dict[str, Tuple[int, ...]]:
    """Key -> shape of the REFERENCE model's state_dict: ours plus the two (d_embed, d_embed) matrices of the
    training-only InfoNCE loss modules (`feature_criterion.W`, `feature_criterion_un.W`).  The filler draws one
    stream over the sorted key list, so tests must hand it the same key set make_golden.py saw."""
    shapes = dict(own_shapes)
    shapes["feature_criterion.W"] = (d_embed, d_embed)
    shapes["feature_criterion_un.W"] = (d_embed, d_embed)
    return shapes


def damp_transformer(values: Dict[str, np.ndarray], scale: float) -> Dict[str, np.ndarray]:
    """Scale the residual branches of the cross-encoder (attention output projections and the second FFN layer) by
    `scale`.  With scale < 1 the conditioned features stay close to the (translation-invariant) KPConv descriptors,
    so corresponding superpoints of a moved copy of a cloud are each other's best match: the filler weights then
    behave like a trained network on such a pair and the end-to-end pose is WELL-CONDITIONED (see
    tests/golden/make_golden.py:gen_forward_wellcond)."""
    out = dict(values)
    for name, v in values.items():
        if name.startswith("transformer_encoder.") and (".out_proj." in name or ".linear2." in name):
            out[name] = (v * np.float32(scale)).astype(np.float32)
    return out
