"""Comparison utilities shared by the parity tests ("what to compare, and how")."""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np


def row_permutation(ref_rows: np.ndarray, our_rows: np.ndarray) -> np.ndarray:
    """perm with our_rows[perm[i]] == ref_rows[i] bit for bit.  Raises when the two row SETS differ."""
    assert ref_rows.shape == our_rows.shape, (ref_rows.shape, our_rows.shape)
    a = np.ascontiguousarray(ref_rows, np.float32).view(np.uint32).reshape(len(ref_rows), -1)
    b = np.ascontiguousarray(our_rows, np.float32).view(np.uint32).reshape(len(our_rows), -1)
    table = {}
    for j, row in enumerate(map(bytes, b)):
        if row in table:
            raise AssertionError("duplicate rows: permutation not unique")
        table[row] = j
    perm = np.empty(len(a), np.int64)
    for i, row in enumerate(map(bytes, a)):
        if row not in table:
            raise AssertionError(f"row {i} of the reference has no bit-identical counterpart")
        perm[i] = table[row]
    return perm


def cloud_of_rows(lengths) -> np.ndarray:
    return np.repeat(np.arange(len(lengths)), np.asarray(lengths))


def trim_shadow_columns(idx: np.ndarray, shadow: int) -> np.ndarray:
    """Drop trailing columns that are shadow in every row (width = min(max_count, limit) in the reference)."""
    if idx.size == 0:
        return idx
    keep = idx.shape[1]
    while keep > 0 and np.all(idx[:, keep - 1] == shadow):
        keep -= 1
    return idx[:, :keep]


def sqdist32(q: np.ndarray, s: np.ndarray) -> np.ndarray:
    d = (q.astype(np.float32) - s.astype(np.float32))
    xx, yy, zz = d[..., 0] * d[..., 0], d[..., 1] * d[..., 1], d[..., 2] * d[..., 2]
    return (xx + yy) + zz


def compare_neighbor_matrices(ours: np.ndarray, ref: np.ndarray, n_support: int, q_pts_ours: np.ndarray,
                              s_pts_ours: np.ndarray, query_perm: Optional[np.ndarray] = None,
                              support_perm: Optional[np.ndarray] = None) -> Dict[str, int]:
    """Row-wise sorted-set comparison of two neighbour matrices.

    ours / ref: [Nq, W] index matrices with shadow value n_support.  query_perm[i] = our row of reference row i;
    support_perm[j] = our index of reference support j (None = identity).  Rows that differ are accepted only
    when the difference lies inside an exact-distance tie at the truncation cut: the fp32 distance multisets
    of the two rows must then be identical.
    Returns counts; raises AssertionError on a real mismatch.
    """
    ours = trim_shadow_columns(np.asarray(ours, np.int64), n_support)
    ref = trim_shadow_columns(np.asarray(ref, np.int64), n_support)
    assert ours.shape[0] == ref.shape[0], (ours.shape, ref.shape)
    if support_perm is not None:
        lut = np.concatenate([support_perm, [n_support]])
        ref = lut[ref]
    if query_perm is not None:
        aligned = np.empty_like(ref)
        aligned[query_perm] = ref
        ref = aligned
    assert ours.shape[1] == ref.shape[1], f"widths differ: ours {ours.shape[1]} vs reference {ref.shape[1]}"
    so, sr = np.sort(ours, 1), np.sort(ref, 1)
    bad = np.nonzero((so != sr).any(1))[0]
    ties = 0
    sp = np.concatenate([s_pts_ours, np.full((1, 3), np.inf, np.float32)], 0)
    for r in bad:
        d_o = np.sort(sqdist32(q_pts_ours[r][None], sp[ours[r]]))
        d_r = np.sort(sqdist32(q_pts_ours[r][None], sp[ref[r]]))
        if not np.array_equal(d_o, d_r):
            raise AssertionError(f"row {r}: neighbour sets differ beyond a distance tie\n ours {so[r]}\n ref  {sr[r]}")
        ties += 1
    order_equal = int((ours == ref).all(1).sum())
    return {"rows": int(ours.shape[0]), "tie_rows": ties, "order_equal_rows": order_equal, "width": int(ours.shape[1])}


def compare_pyramids(ours: Dict[str, List[np.ndarray]], ref: Dict[str, List[np.ndarray]]) -> List[Dict[str, int]]:
    """Full pyramid comparison: lengths exact, points bit-identical as row sets (per cloud), every index matrix
    equal as sorted sets after relabelling through the per-level permutations."""
    L = len(ref["points"])
    assert len(ours["points"]) == L
    perms = []
    for l in range(L):
        lo, lr = np.asarray(ours["stack_lengths"][l]), np.asarray(ref["stack_lengths"][l])
        assert np.array_equal(lo.astype(np.int64), lr.astype(np.int64)), f"level {l}: stack_lengths differ"
        po, pr = np.asarray(ours["points"][l], np.float32), np.asarray(ref["points"][l], np.float32)
        perm = row_permutation(pr, po)
        assert np.array_equal(cloud_of_rows(lo)[perm], cloud_of_rows(lr)), f"level {l}: a point changed cloud"
        perms.append(perm)
    reports = []
    for l in range(L):
        po = np.asarray(ours["points"][l], np.float32)
        n_l = len(po)
        if ref["neighbors"][l].shape[0] > 0:
            rep = compare_neighbor_matrices(ours["neighbors"][l], ref["neighbors"][l], n_l, po, po, perms[l], perms[l])
            rep.update(level=l, kind="neighbors")
            reports.append(rep)
        else:
            assert ours["neighbors"][l].shape[0] == 0
        if l + 1 < L and ref["pools"][l].shape[0] > 0:
            pn = np.asarray(ours["points"][l + 1], np.float32)
            rep = compare_neighbor_matrices(ours["pools"][l], ref["pools"][l], n_l, pn, po, perms[l + 1], perms[l])
            rep.update(level=l, kind="pools")
            reports.append(rep)
            rep = compare_neighbor_matrices(ours["upsamples"][l], ref["upsamples"][l], len(pn), po, pn, perms[l],
                                            perms[l + 1])
            rep.update(level=l, kind="upsamples")
            reports.append(rep)
        else:
            assert ours["pools"][l].shape[0] == 0 and ours["upsamples"][l].shape[0] == 0
    return reports


def load_pyramid(npz, prefix: str = "") -> Dict[str, List[np.ndarray]]:
    L = int(npz[f"{prefix}n_levels"])
    return {k: [npz[f"{prefix}{k}_{l}"] for l in range(L)]
            for k in ("points", "neighbors", "pools", "upsamples", "stack_lengths")}


def pose_error(pred, gt) -> Tuple[np.ndarray, np.ndarray]:
    p, g = np.asarray(pred, np.float64), np.asarray(gt, np.float64)
    fro = np.linalg.norm(p[..., :3, :3] - g[..., :3, :3], axis=(-2, -1))
    rot = 2 * np.degrees(np.arcsin(np.clip(fro / (2 * np.sqrt(2)), 0, 1)))
    return rot, np.linalg.norm(p[..., :3, 3] - g[..., :3, 3], axis=-1)


# ------------------------------------------------------------------------------------------------------
# Summation-order coupling between pyramid levels
# ------------------------------------------------------------------------------------------------------
# A barycentre is a SEQUENTIAL fp32 sum over the voxel's members in the order the previous level emitted
# them (grid_subsampling.h:74-78).  The reference emits a level in std::unordered_map iteration order, we
# emit it in first-occurrence order, so from level 2 on the same voxel can be summed in a different order
# and differ in the last ulp.  Parity is therefore established in two complementary ways:
#   stage-wise  every operator is fed the REFERENCE's own arrays (its order) -> results must be bit-exact;
#   end-to-end  our whole pyramid against the reference's: bit-exact through level 1, and from level 2 on
#               points equal to a few ulp with index rows allowed to differ only by neighbours whose d2 sits
#               within rounding distance of r2 (counted and reported).

def approx_row_permutation(ref_rows: np.ndarray, our_rows: np.ndarray, ref_lens, tol: float) -> np.ndarray:
    """Nearest-row matching inside each cloud; every match must be closer than tol and the map a bijection."""
    from scipy.spatial import cKDTree
    perm = np.empty(len(ref_rows), np.int64)
    o = 0
    for n in np.asarray(ref_lens).tolist():
        tree = cKDTree(our_rows[o:o + n].astype(np.float64))
        d, j = tree.query(ref_rows[o:o + n].astype(np.float64))
        assert d.max() <= tol, f"a reference point has no counterpart within {tol}: {d.max()}"
        assert len(np.unique(j)) == n, "matching is not a bijection"
        perm[o:o + n] = j + o
        o += n
    return perm


def compare_neighbor_matrices_loose(ours, ref, n_support, q_pts, s_pts, radius, query_perm, support_perm,
                                    rel_band: float = 1e-5):
    """Like compare_neighbor_matrices, but a row may also differ by members whose d2 is within rel_band*r2 of
    r2 (in-radius flips caused by last-ulp coordinate differences) or within rel_band of the row's cut distance."""
    ours = np.asarray(ours, np.int64)
    ref = np.asarray(ref, np.int64)
    lut = np.concatenate([support_perm, [n_support]])
    ref = lut[ref]
    aligned = np.empty_like(ref)
    aligned[query_perm] = ref
    ref = aligned
    r2 = np.float64(np.float32(radius)) ** 2
    sp = np.concatenate([s_pts, np.full((1, 3), np.inf, np.float32)], 0).astype(np.float64)
    flips = 0
    for r in range(ours.shape[0]):
        a, b = set(ours[r].tolist()) - {n_support}, set(ref[r].tolist()) - {n_support}
        if a == b:
            continue
        diff = np.asarray(sorted(a ^ b))
        d2 = ((sp[diff] - q_pts[r].astype(np.float64)) ** 2).sum(1)
        common = np.asarray(sorted(a & b))
        cut = ((sp[common] - q_pts[r].astype(np.float64)) ** 2).sum(1).max() if len(common) else r2
        near_r = np.abs(d2 - r2) <= rel_band * r2
        near_cut = np.abs(d2 - cut) <= rel_band * max(cut, 1e-12)
        assert np.all(near_r | near_cut), f"row {r}: difference not explained by rounding: d2={d2}, r2={r2}, cut={cut}"
        flips += 1
    return {"rows": int(ours.shape[0]), "flip_rows": flips}


def compare_pyramids_e2e(ours, ref, cfg, exact_levels: int = 2):
    """End-to-end comparison (see the block comment above)."""
    L = len(ref["points"])
    assert len(ours["points"]) == L
    perms, exact = [], []
    for l in range(L):
        lo, lr = np.asarray(ours["stack_lengths"][l]), np.asarray(ref["stack_lengths"][l])
        assert np.array_equal(lo.astype(np.int64), lr.astype(np.int64)), f"level {l}: stack_lengths differ"
        po, pr = np.asarray(ours["points"][l], np.float32), np.asarray(ref["points"][l], np.float32)
        if l < exact_levels:
            perms.append(row_permutation(pr, po))
            exact.append(True)
        else:
            scale = float(np.abs(pr).max()) if len(pr) else 1.0
            perms.append(approx_row_permutation(pr, po, lr, tol=scale * 1e-5))
            exact.append(bool(np.array_equal(po[perms[-1]].view(np.uint32), pr.view(np.uint32))))
    r = float(cfg["first_subsampling_dl"]) * float(cfg["conv_radius"])
    report = {"levels": L, "points_bit_exact": exact, "flip_rows": 0, "rows": 0, "tie_rows": 0}
    for l in range(L):
        po = np.asarray(ours["points"][l], np.float32)
        n_l = len(po)
        jobs = []
        if ref["neighbors"][l].shape[0] > 0:
            jobs.append((ours["neighbors"][l], ref["neighbors"][l], n_l, po, po, r, perms[l], perms[l], l))
        if l + 1 < L and ref["pools"][l].shape[0] > 0:
            pn = np.asarray(ours["points"][l + 1], np.float32)
            jobs.append((ours["pools"][l], ref["pools"][l], n_l, pn, po, r, perms[l + 1], perms[l], l + 1))
            jobs.append((ours["upsamples"][l], ref["upsamples"][l], len(pn), po, pn, 2 * r, perms[l], perms[l + 1],
                         l + 1))
        for (o, rf, ns, q, s, rad, qp, spm, max_level) in jobs:
            if max_level < exact_levels:
                rep = compare_neighbor_matrices(o, rf, ns, q, s, qp, spm)
                report["tie_rows"] += rep["tie_rows"]
                report["rows"] += rep["rows"]
            else:
                o_t, r_t = trim_shadow_columns(np.asarray(o, np.int64), ns), trim_shadow_columns(np.asarray(rf, np.int64), ns)
                w = max(o_t.shape[1], r_t.shape[1])
                pad = lambda m: np.concatenate([m, np.full((m.shape[0], w - m.shape[1]), ns, np.int64)], 1)
                rep = compare_neighbor_matrices_loose(pad(o_t), pad(r_t), ns, q, s, rad, qp, spm)
                report["flip_rows"] += rep["flip_rows"]
                report["rows"] += rep["rows"]
        r *= 2
    return report


def compare_pyramid_stagewise(ref, cfg, neighbors_fn, subsample_fn):
    """Feed every operator the reference's own arrays; results must be bit-exact (index rows as sorted sets).

    neighbors_fn(q, s, q_lens, s_lens, radius, limit) -> index matrix [Nq, <=limit] (shadow = len(s))
    subsample_fn(points, lens, dl) -> (points, lens)
    """
    L = len(ref["points"])
    limits = cfg["neighborhood_limits"]
    r = float(cfg["first_subsampling_dl"]) * float(cfg["conv_radius"])
    reports = []
    for l in range(L):
        p, ln = np.asarray(ref["points"][l], np.float32), np.asarray(ref["stack_lengths"][l], np.int32)
        if ref["neighbors"][l].shape[0] > 0:
            ours = neighbors_fn(p, p, ln, ln, r, limits[l])
            reports.append(compare_neighbor_matrices(ours, ref["neighbors"][l], len(p), p, p))
        if l + 1 < L and ref["pools"][l].shape[0] > 0:
            pn, lnn = np.asarray(ref["points"][l + 1], np.float32), np.asarray(ref["stack_lengths"][l + 1], np.int32)
            sp, sl = subsample_fn(p, ln, 2 * r / float(cfg["conv_radius"]))
            assert np.array_equal(np.asarray(sl, np.int64), lnn.astype(np.int64))
            o = 0
            for n in lnn.tolist():
                row_permutation(pn[o:o + n], np.asarray(sp, np.float32)[o:o + n])  # bit-identical row sets per cloud
                o += n
            ours = neighbors_fn(pn, p, lnn, ln, r, limits[l])
            reports.append(compare_neighbor_matrices(ours, ref["pools"][l], len(p), pn, p))
            ours = neighbors_fn(p, pn, ln, lnn, 2 * r, limits[l])
            reports.append(compare_neighbor_matrices(ours, ref["upsamples"][l], len(pn), p, pn))
        r *= 2
    return reports
