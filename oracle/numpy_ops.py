"""oracle/numpy_ops.py -- NumPy restatement of the floating-point stages of the hot path.  TEST INFRASTRUCTURE ONLY.

Each function cites the reference lines it follows (paths relative to /root/reference/src).  Parity status:
pinned by tests/test_oracle_golden.py against vectors produced by the imported reference
(tests/golden/make_golden.py).  dtype=np.float32 reproduces the reference's working precision;
dtype=np.float64 gives the exact-arithmetic value used to put fp32 differences into perspective.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np


def _lse(x: np.ndarray, axis: int) -> np.ndarray:
    m = x.max(axis=axis, keepdims=True)
    return m + np.log(np.exp(x - m).sum(axis=axis, keepdims=True, dtype=x.dtype))


def _softmax(x: np.ndarray, axis: int) -> np.ndarray:
    e = np.exp(x - x.max(axis=axis, keepdims=True))
    return e / e.sum(axis=axis, keepdims=True, dtype=x.dtype)


def instance_norm_lrelu(x, lengths, eps=1e-5, slope=1.0, residual=None):
    """models/backbone_kpconv/kpconv_blocks.py:510-519 (per-cloud nn.InstanceNorm1d: biased variance, eps inside
    the sqrt, no affine) followed by the LeakyReLU of :556-561 / the residual sum + LeakyReLU of :741."""
    x = np.asarray(x, np.float32)
    out = np.empty_like(x)
    o = 0
    for n in np.asarray(lengths).tolist():
        seg = x[o:o + n].astype(np.float64)
        mean = seg.mean(0)
        var = seg.var(0)
        out[o:o + n] = ((seg - mean) / np.sqrt(var + eps)).astype(np.float32)
        o += n
    if residual is not None:
        out = out + np.asarray(residual, np.float32)
    return np.where(out >= 0, out, out * np.float32(slope)).astype(np.float32)


def max_pool(x, inds):
    """kpconv_blocks.py:127-143: a zero row is appended for the shadow index, then max over the neighbourhood."""
    x = np.asarray(x, np.float32)
    xp = np.concatenate([x, np.zeros_like(x[:1])], 0)
    return xp[np.asarray(inds)].max(1)


def kpconv_forward_numpy(q_pts, s_pts, neighb_inds, x, weights, kernel_points, extent, dtype=np.float64):
    """kpconv_blocks.py:305-414 (non-deformable) written with dense arrays, for SMALL inputs (it materialises
    [N,H,K] like the reference does).  The C version (spr_oracle.c:orc_kpconv_forward) is the fast one."""
    q, s = np.asarray(q_pts, np.float32), np.asarray(s_pts, np.float32)
    idx = np.asarray(neighb_inds)
    sp = np.concatenate([s, np.full((1, 3), 1e6, np.float32)], 0)                  # :309
    nb = sp[idx] - q[:, None, :]                                                   # :312-315
    diff = nb[:, :, None, :] - np.asarray(kernel_points, np.float32)[None, None]   # :324-325
    sq = (diff.astype(np.float32) ** 2).sum(-1, dtype=np.float32)                  # :328
    w = np.clip(1 - np.sqrt(sq) / np.float32(extent), 0, None).astype(np.float32)  # :368
    w = w.transpose(0, 2, 1).astype(dtype)                                         # [N,K,H]
    xp = np.concatenate([np.asarray(x, np.float32), np.zeros((1, np.shape(x)[1]), np.float32)], 0)  # :388
    nx = xp[idx]                                                                   # :391
    wf = w @ nx.astype(dtype)                                                      # :394  [N,K,Cin]
    out = np.einsum('nkc,kco->no', wf, np.asarray(weights).astype(dtype))          # :401-406
    nn = (nx.sum(-1, dtype=np.float32) > 0).sum(-1)                                # :409-410
    nn = np.maximum(nn, 1)                                                         # :411
    return (out / nn[:, None]).astype(np.float32)


def dual_softmax_match(src_feat, tgt_feat, dtype=np.float32):
    """models/qk_regtr_full.py:453-468 / :565-576 for one pair.
    -> corr (N,M), attn (N,M), val, ind  (per target if N > M else per source)."""
    S, T = np.asarray(src_feat, dtype), np.asarray(tgt_feat, dtype)
    N, D = S.shape
    M = T.shape[0]
    corr = (S @ T.T) / dtype(D ** 0.5)                          # :453
    attn = _softmax(corr, axis=0) * _softmax(corr, axis=1)      # :457-459 (dim=-2 is axis 0, dim=-1 is axis 1)
    if N > M:
        ind = attn.argmax(axis=0)                               # :468 max over dim=1 of (1,N,M)
        val = attn.max(axis=0)
    else:
        ind = attn.argmax(axis=1)                               # :576
        val = attn.max(axis=1)
    return corr, attn, val, ind.astype(np.int64)


def sinkhorn_log(log_alpha, n_iters: int):
    """utils/se3_torch.py:166-202 for one (J,K) matrix: zero-pad a slack row and column, then alternately
    normalise rows (all but the slack row) and columns (all but the slack column) in log space."""
    la = np.pad(log_alpha, ((0, 1), (0, 1)))                                        # :182-184
    for _ in range(n_iters):
        la = np.concatenate([la[:-1, :] - _lse(la[:-1, :], axis=1), la[-1:, :]], 0)  # :188-192
        la = np.concatenate([la[:, :-1] - _lse(la[:, :-1], axis=0), la[:, -1:]], 1)  # :194-198
    return la[:-1, :-1]                                                              # :201


def sinkhorn_weighted_targets(corr, tgt_xyz, softplus_alpha, exp_beta, n_iters, dtype=np.float32):
    """qk_regtr_full.py:641-647 (score = clamp(corr, 0); affinity) + se3_torch.py:209-231
    (perm = exp(sinkhorn); weighted_t = perm @ tgt / (rowsum + 1e-6); weights = rowsum)."""
    c = np.asarray(corr, dtype)
    score = np.clip(c, 0, None)
    aff = -(score - dtype(softplus_alpha)) / (dtype(exp_beta) + dtype(0.02))
    perm = np.exp(sinkhorn_log(aff, n_iters))
    rows = perm.sum(1, keepdims=True, dtype=dtype)
    wt = perm @ np.asarray(tgt_xyz, dtype) / (rows + dtype(1e-6))
    return wt, rows[:, 0]


def compute_rigid_transform(a, b, weights=None, dtype=np.float32):
    """utils/se3_torch.py:109-163 for one pair."""
    a, b = np.asarray(a, dtype), np.asarray(b, dtype)
    if weights is not None:
        w = np.asarray(weights, dtype)
        wn = (w / np.maximum(w.sum(dtype=dtype), dtype(1e-6)))[:, None]      # :137-138
        ca = (a * wn).sum(0, dtype=dtype)                                    # :139-140
        cb = (b * wn).sum(0, dtype=dtype)
        cov = (a - ca).T @ ((b - cb) * wn)                                   # :141-143
    else:
        ca, cb = a.mean(0, dtype=dtype), b.mean(0, dtype=dtype)              # :145-150
        cov = (a - ca).T @ (b - cb)
    u, s, vt = np.linalg.svd(cov)                                            # :152
    v = vt.T
    rot = v @ u.T                                                            # :153
    if not np.linalg.det(rot) > 0:                                           # :154-158
        v2 = v.copy()
        v2[:, 2] *= -1
        rot = v2 @ u.T
    t = -rot @ ca + cb                                                       # :161
    return np.concatenate([rot, t[:, None]], 1).astype(np.float32)


def ratio_test(attn, axis: int, lowe_thres: float):
    """RegTR.ratio_test (models/qk_regtr_full.py:370-384): the largest value along `axis` where second/first is below
    the threshold, else 0, and the position of the largest."""
    order = np.argsort(-attn, axis=axis, kind="stable")
    first = np.take(order, 0, axis=axis)
    top = np.take_along_axis(attn, order, axis=axis)
    v1, v2 = np.take(top, 0, axis=axis), np.take(top, 1, axis=axis)
    with np.errstate(divide="ignore", invalid="ignore"):
        keep = (v2 / v1) < lowe_thres
    return np.where(keep, v1, 0).astype(attn.dtype), first


def recompute_weights(src, tgt, weights, pose, acceptance_radius: float, dtype=np.float32):
    """RegTR.recompute_weights (:386-391)."""
    src, tgt, pose = src.astype(dtype), tgt.astype(dtype), pose.astype(dtype)
    res = np.linalg.norm(tgt - (src @ pose[:, :3].T + pose[:, 3]), axis=1)
    return weights.astype(dtype) * (res < acceptance_radius)


def local_global_registration(src, tgt, weights, pose, acceptance_radius: float, steps: int, dtype=np.float32):
    """RegTR.local_global_registration (:393-398)."""
    for _ in range(steps):
        weights = recompute_weights(src, tgt, weights, pose, acceptance_radius, dtype)
        pose = compute_rigid_transform(src, tgt, weights, dtype=dtype)
    return pose


def ransac(src, tgt, weights, sample_idx, dtype=np.float32):
    """RegTR.ransac (:400-421) with the index draws given (sample_idx [hypotheses, sample_size]); the reference draws
    them from the CUDA generator, so only this restatement can be compared number for number."""
    best, best_loss, losses = None, None, []
    for idx in sample_idx:
        T = compute_rigid_transform(src[idx], tgt[idx], weights[idx], dtype=dtype)
        loss = np.linalg.norm(tgt.astype(dtype) - (src.astype(dtype) @ T[:, :3].T + T[:, 3]), axis=1).mean()
        losses.append(loss)
        if best is None or loss < best_loss:
            best, best_loss = T, loss
    return best, np.asarray(losses)


def pose_error(pred, gt) -> Tuple[float, float]:
    """Chordal rotation error (deg) and translation error in fp64 -- see SURVEY.md row a13 for why this is used
    instead of the reference's se3_compare."""
    p, g = np.asarray(pred, np.float64), np.asarray(gt, np.float64)
    fro = np.linalg.norm(p[..., :3, :3] - g[..., :3, :3], axis=(-2, -1))
    rot = 2 * np.degrees(np.arcsin(np.clip(fro / (2 * np.sqrt(2)), 0, 1)))
    tr = np.linalg.norm(p[..., :3, 3] - g[..., :3, 3], axis=-1)
    return rot, tr


# ------------------------------------------------------------------------------------------------
# PreprocessorGPU-compatible mode (models/backbone_kpconv/kpconv.py:421-549) and the KITTI front end.
# PARITY UNPINNED: the arithmetic lives in pytorch3d (ball_query), MinkowskiEngine (sparse quantisation) and kiss_icp
# (voxel down-sampling), none of which is vendored, pinned or installed; the functions below restate their DOCUMENTED
# behaviour as the reference calls them, and the CUDA path is tested against these restatements only.
# ------------------------------------------------------------------------------------------------

def ball_query_first_k(queries, supports, q_lens, s_lens, radius, K):
    """kpconv.py:265-292 (batch_neighbors_kpconv_gpu): pytorch3d.ops.ball_query keeps, per query, the FIRST K supports of
    the same cloud in index order with d2 < r2 (not the K nearest); the matrix is always K wide, missing entries are
    Ns_total.  d2 with the same fp32 rounding sequence as the CPU path (three products, two sums)."""
    q, s = np.asarray(queries, np.float32), np.asarray(supports, np.float32)
    out = np.full((len(q), K), len(s), np.int64)
    r2 = np.float32(radius) * np.float32(radius)
    qo = so = 0
    for nq, ns in zip(q_lens, s_lens):
        sc = s[so:so + ns]
        for i in range(qo, qo + nq):
            d = q[i] - sc
            d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
            hit = np.nonzero(d2 < r2)[0][:K]
            out[i, :len(hit)] = hit + so
        qo += nq
        so += ns
    return out


def voxel_mean_me(points, lens, dl):
    """kpconv.py:221-243 (batch_grid_subsampling_kpconv_gpu): MinkowskiEngine quantises coordinates floor(p / dl) (global
    origin, per batch element) and averages the points of a voxel (UNWEIGHTED_AVERAGE).  ME's output order and
    summation order are unspecified ("not deterministic", :219-220); here: voxels in first-occurrence order per cloud,
    members summed in input order in fp32, divided by the count."""
    pts = np.asarray(points, np.float32)
    dl32 = np.float32(dl)
    out, out_lens, o = [], [], 0
    for n in lens:
        seg = pts[o:o + n]
        keys = np.floor(seg / dl32).astype(np.int64)
        order, sums, counts = {}, [], []
        for i, k in enumerate(map(tuple, keys)):
            v = order.get(k)
            if v is None:
                order[k] = len(sums)
                sums.append(seg[i].copy())
                counts.append(1)
            else:
                sums[v] = sums[v] + seg[i]          # fp32, input order
                counts[v] += 1
        out.append(np.stack([sm / np.float32(c) for sm, c in zip(sums, counts)]) if sums else np.zeros((0, 3), np.float32))
        out_lens.append(len(sums))
        o += n
    return np.concatenate(out, 0).astype(np.float32), np.asarray(out_lens, np.int32)


def voxel_first_point(points, voxel):
    """data_loaders/kitti_pred.py:12-14 (kiss_icp VoxelDownsample): the first point, in input order, of every voxel
    floor(p / voxel); emitted here in input order (kiss_icp emits its hash map's order)."""
    pts = np.asarray(points, np.float32)
    keys = np.floor(pts / np.float32(voxel)).astype(np.int64)
    seen, keep = set(), []
    for i, k in enumerate(map(tuple, keys)):
        if k not in seen:
            seen.add(k)
            keep.append(i)
    return pts[np.asarray(keep, np.int64)]
