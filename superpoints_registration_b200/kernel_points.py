"""Kernel-point dispositions for KPConv initialisation.

Reference: models/backbone_kpconv/kernels/kernel_points.py:387-469 (`load_kernels`): a cached disposition is
read from `kernels/dispositions/k_{K:03d}_{fixed}_{dim}D.ply` relative to the working directory (or produced
by a repulsion optimisation), then rotated about z by a random angle, jittered with N(0, 0.01) noise and scaled
by the convolution radius -- all from the global NumPy RNG.

Trained / reference weights always carry their own `kernel_points` in the state_dict
(`...KPConv.kernel_points`), so parity never depends on this initialiser; it exists so that a freshly
constructed module is usable (bench.py, smoke()).  The base disposition here is our own: the centre plus
K-1 points spread over a sphere by a short electrostatic-repulsion descent, rescaled so that the mean radius
of the non-centre points is 0.66 (the reference's `ratio`, kernel_points.py:381).
"""
from __future__ import annotations

import numpy as np

_BASE_CACHE = {}


def _repulsion_sphere(n: int, iters: int = 400) -> np.ndarray:
    """n points on the unit sphere pushed apart by 1/d^2 forces (deterministic start: Fibonacci lattice)."""
    i = np.arange(n, dtype=np.float64) + 0.5
    phi = np.arccos(1.0 - 2.0 * i / n)
    theta = np.pi * (1.0 + 5.0 ** 0.5) * i
    p = np.stack([np.cos(theta) * np.sin(phi), np.sin(theta) * np.sin(phi), np.cos(phi)], axis=1)
    step = 0.05
    for _ in range(iters):
        d = p[:, None, :] - p[None, :, :]
        r2 = (d ** 2).sum(-1) + np.eye(n)
        f = (d / r2[..., None] ** 1.5).sum(1)
        f -= (f * p).sum(1, keepdims=True) * p  # tangential component only
        p = p + step * f
        p /= np.linalg.norm(p, axis=1, keepdims=True)
        step *= 0.995
    return p


def base_disposition(num_kpoints: int, fixed: str = "center") -> np.ndarray:
    key = (num_kpoints, fixed)
    if key not in _BASE_CACHE:
        if fixed != "center":
            raise NotImplementedError("only fixed_kernel_points='center' is used by the shipped configs")
        shell = _repulsion_sphere(num_kpoints - 1) * 0.66
        _BASE_CACHE[key] = np.concatenate([np.zeros((1, 3)), shell], axis=0)
    return _BASE_CACHE[key].copy()


def load_kernels(radius: float, num_kpoints: int, dimension: int = 3, fixed: str = "center") -> np.ndarray:
    if dimension != 3:
        raise NotImplementedError("3-D point clouds only")
    pts = base_disposition(num_kpoints, fixed)
    theta = np.random.rand() * 2 * np.pi
    c, s = np.cos(theta), np.sin(theta)
    rot = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]], dtype=np.float32)
    pts = pts + np.random.normal(scale=0.01, size=pts.shape)
    pts = radius * pts
    return (pts @ rot).astype(np.float32)
