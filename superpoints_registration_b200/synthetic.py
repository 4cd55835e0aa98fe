"""Seeded synthetic point-cloud pairs with the shapes of the reference's datasets (there is no network, so the
real 3DMatch / KITTI / ModelNet40 files are not available).

Every generator returns `src` (N,3) f32, `tgt` (M,3) f32 and the ground-truth `pose` (3,4) f32 with
tgt ~= R src + t on the overlap.  Recipes follow SURVEY.md section 8(d):
  threedmatch_pair  planar patches in a 3 m room, 4 mm plane noise, thinned to one point per voxel (0.025 m)
  modelnet_pair     717 points (data_loaders/modelnet_transforms.py:92-93) from a surface-biased unit-sphere
                    shape, partial crop 0.7, rotation <= 45 deg, translation <= 0.5
  kitti_pair        ground plane + vertical structures seen from a spinning scanner, range <= 80 m, thinned at
                    the first subsampling voxel (0.2 m in conf/qk_regtr_full_kitti.yaml:36)
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np


def _random_rotation(rng: np.random.Generator, max_deg: float) -> np.ndarray:
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    ang = np.deg2rad(rng.uniform(-max_deg, max_deg))
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * K + (1 - np.cos(ang)) * (K @ K)


def voxel_thin(points: np.ndarray, voxel: float) -> np.ndarray:
    """Keep the first point of every voxel (input order preserved)."""
    keys = np.floor(points / voxel).astype(np.int64)
    keys -= keys.min(0)
    dims = keys.max(0) + 1
    flat = (keys[:, 0] * dims[1] + keys[:, 1]) * dims[2] + keys[:, 2]
    _, first = np.unique(flat, return_index=True)
    return points[np.sort(first)]


def _room(rng: np.random.Generator, n_points: int, voxel: float, n_patches: int = 6, extent: float = 3.0) -> np.ndarray:
    """Axis-aligned planar rectangles whose total area is n_points * voxel^2."""
    area = n_points * voxel * voxel / n_patches
    pts = []
    for _ in range(n_patches):
        axis = int(rng.integers(0, 3))
        aspect = rng.uniform(0.6, 1.6)
        w, h = min(np.sqrt(area * aspect), extent), min(np.sqrt(area / aspect), extent)
        dense = int(4.5 * w * h / (voxel * voxel)) + 16
        uv = rng.uniform(0, 1, size=(dense, 2)) * np.array([w, h])
        origin = np.array([rng.uniform(0, extent - w), rng.uniform(0, extent - h)])
        p = np.empty((dense, 3))
        others = [a for a in range(3) if a != axis]
        p[:, others[0]] = origin[0] + uv[:, 0]
        p[:, others[1]] = origin[1] + uv[:, 1]
        p[:, axis] = rng.uniform(0, extent) + rng.normal(0, 0.004, size=dense)
        pts.append(p)
    return np.concatenate(pts, axis=0)


def threedmatch_pair(seed: int, n_points: int = 20000, voxel: float = 0.025, overlap: Tuple[float, float] = (0.5, 0.9)
                     ) -> Dict[str, np.ndarray]:
    rng = np.random.default_rng(seed)
    ov = rng.uniform(*overlap)
    # scene large enough that each cropped fragment has ~n_points points
    # 0.62: plane noise makes a patch occupy ~1.6 voxel layers, so the occupied-voxel count overshoots the area
    scene = _room(rng, int(n_points * (2.0 - ov) * 0.62), voxel)
    direction = rng.normal(size=3)
    direction /= np.linalg.norm(direction)
    proj = scene @ direction
    lo, hi = np.quantile(proj, [0.0, 1.0])
    span = hi - lo
    frac = 1.0 / (2.0 - ov)  # each fragment covers this fraction of the scene; they share `ov` of a fragment
    src_raw = scene[proj <= lo + frac * span]
    tgt_raw = scene[proj >= hi - frac * span]
    src = voxel_thin(src_raw, voxel)
    R = _random_rotation(rng, 45.0)
    t = rng.uniform(-0.5, 0.5, size=3)
    tgt_local = voxel_thin(tgt_raw + rng.normal(0, 0.1 * voxel, size=tgt_raw.shape), voxel)
    tgt = tgt_local @ R.T + t
    rng.shuffle(tgt, axis=0)
    pose = np.concatenate([R, t[:, None]], axis=1)
    return {"src": src.astype(np.float32), "tgt": tgt.astype(np.float32), "pose": pose.astype(np.float32),
            "overlap": float(ov)}


def threedlomatch_pair(seed: int, n_points: int = 20000, voxel: float = 0.025) -> Dict[str, np.ndarray]:
    """Low-overlap pairs: true overlap drawn from U[0.10, 0.30] (range of datasets/3dmatch/test_3DLoMatch_info.pkl)."""
    return threedmatch_pair(seed, n_points, voxel, overlap=(0.10, 0.30))


def modelnet_pair(seed: int, n_points: int = 717, partial: float = 0.7) -> Dict[str, np.ndarray]:
    rng = np.random.default_rng(seed)
    raw = int(n_points / partial) + 8
    d = rng.normal(size=(raw * 2, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    shape = d * (1.0 - 0.25 * np.abs(np.sin(3 * d[:, :1]) * np.cos(2 * d[:, 1:2]))) * rng.uniform(0.85, 1.0, (raw * 2, 1))
    shape *= rng.uniform(0.5, 1.0, size=3)  # anisotropic object

    def crop(points):
        direction = rng.normal(size=3)
        direction /= np.linalg.norm(direction)
        proj = points @ direction
        keep = proj <= np.quantile(proj, partial)
        sel = points[keep]
        idx = rng.permutation(sel.shape[0])[:n_points]
        return sel[idx]

    src = crop(shape[:raw])
    tgt_local = crop(shape[raw:] if False else shape[:raw] + rng.normal(0, 0.002, size=(raw, 3)))
    R = _random_rotation(rng, 45.0)
    t = rng.uniform(-0.5, 0.5, size=3)
    tgt = tgt_local @ R.T + t
    pose = np.concatenate([R, t[:, None]], axis=1)
    return {"src": src.astype(np.float32), "tgt": tgt.astype(np.float32), "pose": pose.astype(np.float32)}


def kitti_pair(seed: int, n_points: int = 30000, voxel: float = 0.2) -> Dict[str, np.ndarray]:
    rng = np.random.default_rng(seed)

    def scan(origin_xy, yaw):
        n_raw = int(n_points * 2.2)
        az = rng.uniform(0, 2 * np.pi, n_raw)
        # ground returns: range skewed towards the sensor
        rg = 3.0 + 77.0 * rng.beta(1.2, 4.0, n_raw)
        ground = np.stack([rg * np.cos(az), rg * np.sin(az), -1.7 + rng.normal(0, 0.03, n_raw)], 1)
        # vertical structures: walls along two street sides + poles
        m = n_raw // 2
        wx = rng.uniform(-70, 70, m)
        side = rng.choice([-1.0, 1.0], m)
        wall = np.stack([wx, side * (8.0 + 0.3 * np.sin(wx * 0.2)), rng.uniform(-1.7, 4.0, m)], 1)
        world = np.concatenate([ground, wall], 0)
        world[:, :2] -= origin_xy
        c, s = np.cos(yaw), np.sin(yaw)
        Rz = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
        local = world @ Rz
        local = local[np.linalg.norm(local[:, :2], axis=1) <= 80.0]
        thinned = voxel_thin(local + rng.normal(0, 0.02, local.shape), voxel)
        if thinned.shape[0] > n_points:
            thinned = thinned[np.sort(rng.permutation(thinned.shape[0])[:n_points])]
        return thinned, Rz, origin_xy

    src, Rs, os_ = scan(np.zeros(2), 0.0)
    move = np.array([rng.uniform(5, 10), rng.uniform(-0.5, 0.5)])
    yaw = np.deg2rad(rng.uniform(-5, 5))
    tgt, Rt, ot = scan(move, yaw)
    # p_world = Rs p_src + os ; p_tgt = Rt^T (p_world - ot)
    R = Rt.T @ Rs
    t = Rt.T @ (np.array([os_[0], os_[1], 0.0]) - np.array([ot[0], ot[1], 0.0]))
    pose = np.concatenate([R, t[:, None]], axis=1)
    return {"src": src.astype(np.float32), "tgt": tgt.astype(np.float32), "pose": pose.astype(np.float32)}


def make_batch(kind: str, n_pairs: int, seed: int = 0, **kw) -> Dict[str, List[np.ndarray]]:
    gen = {"3dmatch": threedmatch_pair, "3dlomatch": threedlomatch_pair, "modelnet": modelnet_pair,
           "kitti": kitti_pair}[kind]
    pairs = [gen(seed * 1000 + i, **kw) for i in range(n_pairs)]
    return {"src_xyz": [p["src"] for p in pairs], "tgt_xyz": [p["tgt"] for p in pairs],
            "pose": np.stack([p["pose"] for p in pairs])}
