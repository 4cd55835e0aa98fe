"""CUDA preprocessing (grid subsampling, radius neighbours, pyramid) against the oracle and the golden vectors.
Everything goes through the C ABI (libspr_b200.so)."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import pipeline
from parity import (compare_neighbor_matrices, compare_pyramid_stagewise, compare_pyramids_e2e, load_pyramid,
                    row_permutation, sqdist32)
from superpoints_registration_b200 import config as cfgs
from superpoints_registration_b200 import ops, synthetic
from superpoints_registration_b200.kpconv import Preprocessor, batch_grid_subsampling_kpconv, batch_neighbors_kpconv

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _t(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def _clouds(kind, n_pairs, seed, **kw):
    b = synthetic.make_batch(kind, n_pairs, seed=seed, **kw)
    return b["src_xyz"] + b["tgt_xyz"]


def _stack(clouds):
    return np.concatenate(clouds).astype(np.float32), np.asarray([len(c) for c in clouds], np.int32)


# ---------------------------------------------------------------------------------------------------------
# grid subsampling: bit-identical to the oracle, same (first-occurrence) order
# ---------------------------------------------------------------------------------------------------------

def _check_subsample(pts, lens, dl):
    op, ol = oracle.grid_subsample_batch(pts, lens, dl)
    gp, gl = ops.grid_subsample_batch(_t(pts), _t(lens), dl)
    assert np.array_equal(gl.cpu().numpy(), ol)
    assert np.array_equal(gp.cpu().numpy().view(np.uint32), op.view(np.uint32))
    return op


@pytest.mark.parametrize("dl", [0.05, 0.1, 0.4])
def test_subsample_3dmatch(dl):
    pts, lens = _stack(_clouds("3dmatch", 2, 3, n_points=5000))
    _check_subsample(pts, lens, dl)


def test_subsample_kitti_scale_negative_coordinates():
    pts, lens = _stack(_clouds("kitti", 1, 4, n_points=8000))
    _check_subsample(pts, lens, 0.4)
    _check_subsample(pts, lens, 1.6)


def test_subsample_crowded_voxels_and_tiny_clouds():
    rng = np.random.default_rng(0)
    # > 16 members per voxel exercises the large-voxel path; single-point and duplicate-point clouds too
    a = rng.uniform(0, 1, size=(4000, 3)).astype(np.float32)
    b = np.asarray([[3, 3, 3]], np.float32)
    c = np.repeat(np.asarray([[7.5, -2, 1]], np.float32), 40, 0)
    pts, lens = _stack([a, b, c])
    out = _check_subsample(pts, lens, 0.25)
    assert len(out) < 200


def test_subsample_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "preprocess_3dmatch.npz"))
    ref = load_pyramid(g)
    cfg = cfgs.threedmatch_config()
    r = cfg.first_subsampling_dl * cfg.conv_radius
    for l in range(len(ref["points"]) - 1):
        gp, gl = batch_grid_subsampling_kpconv(_t(ref["points"][l]), _t(ref["stack_lengths"][l]),
                                               sampleDl=2 * r / cfg.conv_radius)
        assert np.array_equal(gl.cpu().numpy(), ref["stack_lengths"][l + 1])
        o = 0
        for n in ref["stack_lengths"][l + 1].tolist():
            row_permutation(ref["points"][l + 1][o:o + n], gp.cpu().numpy()[o:o + n])
            o += n
        r *= 2


def test_subsample_rejects_empty_and_cpu_input():
    with pytest.raises(RuntimeError):
        ops.grid_subsample_batch(torch.zeros((0, 3), device=DEV), torch.zeros((1,), dtype=torch.int32, device=DEV), 0.1)
    with pytest.raises(RuntimeError):
        ops.grid_subsample_batch(torch.zeros((10, 3)), torch.tensor([10], dtype=torch.int32), 0.1)


# ---------------------------------------------------------------------------------------------------------
# radius neighbours: identical matrices (same (d2, index) order) and identical max_count
# ---------------------------------------------------------------------------------------------------------

def _check_neighbors(q, s, ql, sl, radius, limit, dtype=torch.int64):
    oi, omc = oracle.radius_neighbors_batch(q, s, ql, sl, radius, limit)
    gi, gmc = ops.radius_neighbors_batch(_t(q), _t(s), _t(ql), _t(sl), radius, limit, index_dtype=dtype)
    assert int(gmc.item()) == omc
    assert gi.dtype == dtype
    assert np.array_equal(gi.cpu().numpy().astype(np.int64), oi.astype(np.int64))


@pytest.mark.parametrize("radius,limit", [(0.0625, 40), (0.125, 40), (0.125, 7), (0.25, 128)])
def test_neighbors_3dmatch(radius, limit):
    pts, lens = _stack(_clouds("3dmatch", 2, 5, n_points=4000))
    _check_neighbors(pts, pts, lens, lens, radius, limit)


def test_neighbors_int32_indices():
    pts, lens = _stack(_clouds("3dmatch", 1, 6, n_points=2000))
    _check_neighbors(pts, pts, lens, lens, 0.0625, 40, dtype=torch.int32)


def test_neighbors_dense_rows_overflow_the_stage():
    """> 320 in-radius supports per query forces the stage compaction path."""
    rng = np.random.default_rng(1)
    s = rng.uniform(0, 1, size=(6000, 3)).astype(np.float32)
    lens = np.asarray([6000], np.int32)
    _check_neighbors(s[:700], s, np.asarray([700], np.int32), lens, 0.33, 128)
    _check_neighbors(s[:300], s, np.asarray([300], np.int32), lens, 0.5, 33)


def test_neighbors_ragged_and_outside_queries():
    rng = np.random.default_rng(2)
    s = rng.uniform(0, 2, size=(3000, 3)).astype(np.float32)
    q = rng.uniform(-0.7, 2.7, size=(1000, 3)).astype(np.float32)
    _check_neighbors(q, s, np.asarray([300, 5, 695], np.int32), np.asarray([1, 1999, 1000], np.int32), 0.2, 50)


def test_neighbors_lattice_ties():
    g = np.stack(np.meshgrid(*[np.arange(9)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    lens = np.asarray([len(g)], np.int32)
    _check_neighbors(g, g, lens, lens, 1.5, 10)
    _check_neighbors(g, g, lens, lens, 2.0, 64)


def test_neighbors_kitti_scale():
    pts, lens = _stack(_clouds("kitti", 1, 7, n_points=6000))
    _check_neighbors(pts, pts, lens, lens, 0.85, 39)
    _check_neighbors(pts, pts, lens, lens, 6.8, 74)


def test_neighbors_wrapper_width_and_errors():
    pts, lens = _stack(_clouds("3dmatch", 1, 8, n_points=1500))
    idx = batch_neighbors_kpconv(_t(pts), _t(pts), _t(lens), _t(lens), 0.0625, 40)
    ref = oracle.ref_batch_query(pts, pts, lens, lens, 0.0625)[:, :40] if oracle.have_ref() else None
    if ref is not None:
        assert idx.shape == ref.shape
        compare_neighbor_matrices(idx.cpu().numpy(), ref, len(pts), pts, pts)
    assert idx.dtype == torch.int64
    with pytest.raises(RuntimeError):
        ops.radius_neighbors_batch(_t(pts), _t(pts), _t(lens), _t(lens), 0.1, 129)  # limit too large
    with pytest.raises(RuntimeError):
        ops.radius_neighbors_batch(_t(pts), _t(pts), _t(lens), _t(lens[:1]), 0.1, 10)  # batch count mismatch


# ---------------------------------------------------------------------------------------------------------
# pyramid
# ---------------------------------------------------------------------------------------------------------

def _to_numpy_pyramid(meta):
    return {k: [t.cpu().numpy() for t in v] for k, v in meta.items()}


@pytest.mark.parametrize("name,cfg", [("3dmatch", cfgs.threedmatch_config()), ("kitti", cfgs.kitti_config()),
                                      ("modelnet", cfgs.modelnet_config())])
def test_pyramid_against_golden_reference(golden_dir, name, cfg):
    g = np.load(os.path.join(golden_dir, f"preprocess_{name}.npz"))
    ref = load_pyramid(g)
    clouds = [g[f"cloud_{i}"] for i in range(int(g["n_clouds"]))]
    # stage-wise on the reference's own arrays: bit-exact
    def nb(q, s, ql, sl, radius, limit):
        return batch_neighbors_kpconv(_t(q), _t(s), _t(ql), _t(sl), radius, limit).cpu().numpy()

    def sub(p, l, dl):
        a, b = batch_grid_subsampling_kpconv(_t(p), _t(l), sampleDl=dl)
        return a.cpu().numpy(), b.cpu().numpy()
    assert compare_pyramid_stagewise(ref, cfg, nb, sub)
    # end to end
    meta = Preprocessor(cfg)([_t(c) for c in clouds])
    ours = _to_numpy_pyramid(meta)
    for l in range(len(ours["points"])):
        assert ours["neighbors"][l].dtype == np.int64 and ours["stack_lengths"][l].dtype in (np.int32, np.int64)
        assert ours["neighbors"][l].shape == ref["neighbors"][l].shape  # same trimmed width as the reference
        assert ours["pools"][l].shape == ref["pools"][l].shape
        assert ours["upsamples"][l].shape == ref["upsamples"][l].shape
    rep = compare_pyramids_e2e(ours, ref, cfg)
    assert rep["points_bit_exact"][0] and (rep["levels"] < 2 or rep["points_bit_exact"][1])
    assert rep["flip_rows"] <= max(2, rep["rows"] // 1000), rep


@pytest.mark.parametrize("cfg,kind,kw", [(cfgs.threedmatch_config(), "3dmatch", dict(n_points=6000)),
                                         (cfgs.threedmatch_4stage_config(), "3dmatch", dict(n_points=6000)),
                                         (cfgs.kitti_config(), "kitti", dict(n_points=6000)),
                                         (cfgs.modelnet_config(), "modelnet", {})])
def test_pyramid_identical_to_oracle(cfg, kind, kw):
    """Same canonical order on both sides -> every array of the pyramid must be identical."""
    clouds = _clouds(kind, 3, 9, **kw)
    want = pipeline.preprocess(cfg, clouds, backend="port")
    got = _to_numpy_pyramid(Preprocessor(cfg)([_t(c) for c in clouds]))
    for key in ("points", "neighbors", "pools", "upsamples", "stack_lengths"):
        for l, (a, b) in enumerate(zip(got[key], want[key])):
            assert a.shape == b.shape, (key, l, a.shape, b.shape)
            if key == "points":
                assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (key, l)
            else:
                assert np.array_equal(a.astype(np.int64), b.astype(np.int64)), (key, l)


def test_pyramid_full_size_properties():
    """BASELINE-size input (8 pairs x ~20k points): size-independent properties instead of a CPU comparison."""
    cfg = cfgs.threedmatch_config()
    clouds = _clouds("3dmatch", 8, 2, n_points=20000)
    meta = Preprocessor(cfg)([_t(c) for c in clouds])
    r = cfg.first_subsampling_dl * cfg.conv_radius
    rng = np.random.default_rng(0)
    for l, pts_t in enumerate(meta["points"]):
        pts = pts_t.cpu().numpy()
        lens = meta["stack_lengths"][l].cpu().numpy()
        assert lens.sum() == len(pts)
        idx = meta["neighbors"][l].cpu().numpy()
        n = len(pts)
        starts = np.concatenate([[0], np.cumsum(lens)])
        cloud = np.repeat(np.arange(len(lens)), lens)
        rows = rng.choice(n, size=min(n, 4000), replace=False)
        sp = np.concatenate([pts, np.full((1, 3), np.inf, np.float32)])
        d2 = sqdist32(pts[rows][:, None, :], sp[idx[rows]])
        valid = idx[rows] < n
        r2 = np.float32(r) * np.float32(r)
        assert np.all(d2[valid] < r2)                                     # every neighbour is inside the radius
        assert np.all(idx[rows][:, 0] == rows)                            # nearest neighbour of a point is itself
        assert np.all(np.diff(np.where(valid, d2, np.float32(1e30)), axis=1) >= 0)  # ascending distances, padding last
        same_cloud = cloud[np.minimum(idx[rows], n - 1)] == cloud[rows][:, None]
        assert np.all(same_cloud | ~valid)                                # neighbours never cross clouds
        # exact counts against a brute force on a few rows
        for rr in rows[:50]:
            b = cloud[rr]
            seg = pts[starts[b]:starts[b + 1]]
            cnt = int((sqdist32(pts[rr][None], seg) < r2).sum())
            assert int(valid[list(rows).index(rr)].sum()) == min(cnt, idx.shape[1])
        r *= 2


def test_pyramid_is_int32_inside_and_lazy_where_nobody_looks():
    """The dict hands out the reference's int64 matrices; the kernels read the int32 ones behind them (Pyramid.index),
    and 'upsamples' is searched on first access.  Same values whichever way the pyramid is built."""
    from superpoints_registration_b200 import _lib
    cfg = cfgs.threedmatch_4stage_config()
    clouds = [_t(c) for c in _clouds("3dmatch", 2, 4, n_points=5000)]
    n0 = _lib.launch_count()
    meta = Preprocessor(cfg)(clouds)
    lazy_launches = _lib.launch_count() - n0
    n0 = _lib.launch_count()
    eager = Preprocessor(cfg, lazy_upsamples=False)(clouds)
    eager_launches = _lib.launch_count() - n0
    assert lazy_launches < eager_launches                       # three searches were not run
    assert set(meta.keys()) == {"points", "neighbors", "pools", "upsamples", "stack_lengths"}
    n0 = _lib.launch_count()
    for key in ("neighbors", "pools"):
        for l in range(len(meta["points"])):
            wide, raw = meta[key][l], meta.index(key, l)
            assert wide.dtype == torch.int64 and raw.dtype == torch.int32 and wide.shape == raw.shape
            assert torch.equal(wide, raw.long()) and torch.equal(wide, eager[key][l])
            assert meta[key][l] is wide                          # converted once
    assert _lib.launch_count() == n0                             # nothing searched so far
    ups = list(meta["upsamples"])                                # first access runs the three searches
    assert _lib.launch_count() > n0
    assert len(ups) == len(meta["points"]) and tuple(ups[-1].shape) == (0, 1)
    for l, u in enumerate(ups):
        assert u.dtype == torch.int64 and torch.equal(u, eager["upsamples"][l])
    m32 = Preprocessor(cfg, index_dtype=torch.int32)(clouds)
    assert m32["neighbors"][0].dtype == torch.int32 and torch.equal(m32["neighbors"][0].long(), meta["neighbors"][0])


# ---------------------------------------------------------------------------------------------------------
# PreprocessorGPU-compatible mode and the KITTI front end.  PARITY UNPINNED: pytorch3d,
# MinkowskiEngine and kiss_icp are absent; the oracle is the NumPy restatement of their documented behaviour.
# ---------------------------------------------------------------------------------------------------------

def test_ball_query_rows_keep_the_first_k_in_index_order():
    from oracle import numpy_ops
    rng = np.random.default_rng(31)
    ql, sl = np.array([300, 0, 150, 41], np.int32), np.array([400, 0, 220, 3], np.int32)
    q = rng.uniform(0, 1, size=(int(ql.sum()), 3)).astype(np.float32)
    s = rng.uniform(0, 1, size=(int(sl.sum()), 3)).astype(np.float32)
    for radius, K in ((0.2, 12), (0.45, 40), (0.05, 8)):
        grid = ops.CellGrid(_t(s), _t(sl), radius)
        for dtype in (torch.int32, torch.int64):
            got, mc = grid.query(_t(q), _t(ql), K, index_dtype=dtype, by_index=True)
            want = numpy_ops.ball_query_first_k(q, s, ql, sl, radius, K)
            assert got.dtype == dtype and np.array_equal(got.cpu().numpy().astype(np.int64), want), (radius, K)
        near, _ = grid.query(_t(q), _t(ql), K, by_index=False)
        # same neighbour SETS whenever nothing is truncated
        full = (want < len(s)).sum(1) < K
        assert np.array_equal(np.sort(near.cpu().numpy()[full], 1), np.sort(want[full], 1))


def test_voxel_mean_and_first_point_modes():
    from oracle import numpy_ops
    from superpoints_registration_b200 import frontends
    rng = np.random.default_rng(32)
    lens = np.array([5000, 1, 0, 2500], np.int32)
    pts = (rng.normal(size=(int(lens.sum()), 3)) * np.array([2.0, 1.5, 0.3])).astype(np.float32)   # both signs: global lattice
    for dl in (0.11, 0.4):
        got_p, got_l = ops.grid_subsample_batch(_t(pts), _t(lens), dl, mode="mean")
        want_p, want_l = numpy_ops.voxel_mean_me(pts, lens, dl)
        assert np.array_equal(got_l.cpu().numpy(), want_l)
        assert np.array_equal(got_p.cpu().numpy().view(np.uint32), want_p.view(np.uint32))       # bit-identical means
    cloud = pts[:5000]
    got = frontends.voxel_down_sample(_t(cloud), 0.3).cpu().numpy()
    want = numpy_ops.voxel_first_point(cloud, 0.3)
    assert np.array_equal(got, want)                              # a subset of the input, in input order
    with pytest.raises(RuntimeError):
        frontends.voxel_down_sample(torch.from_numpy(cloud), 0.3)  # CPU tensors are refused


@pytest.mark.parametrize("cfg,kind,kw", [(cfgs.threedmatch_4stage_config(), "3dmatch", dict(n_points=5000)),
                                         (cfgs.kitti_config(), "kitti", dict(n_points=5000))])
def test_gpu_compat_pyramid_against_restatement(cfg, kind, kw):
    """Preprocessor(mode='gpu_compat') / PreprocessorGPU level by level against the NumPy restatement of
    kpconv.py:421-549: ball_query rows (first K by index, always `limit` wide), voxel means on the global lattice,
    int64 lengths, the reference's radius / voxel schedule and last-level placeholders."""
    from oracle import numpy_ops
    from superpoints_registration_b200 import PreprocessorGPU
    clouds = _clouds(kind, 2, 6, **kw)
    meta = PreprocessorGPU(cfg)([_t(c) for c in clouds])
    assert set(meta.keys()) == {"points", "neighbors", "pools", "upsamples", "stack_lengths"}
    pts = np.concatenate(clouds).astype(np.float32)
    lens = np.array([len(c) for c in clouds], np.int32)
    r = cfg.first_subsampling_dl * cfg.conv_radius
    L = len(meta["points"])
    for l in range(L):
        limit = cfg.neighborhood_limits[l]
        assert meta["stack_lengths"][l].dtype == torch.int64
        assert np.array_equal(meta["stack_lengths"][l].cpu().numpy(), lens)
        assert np.array_equal(meta["points"][l].cpu().numpy().view(np.uint32), pts.view(np.uint32))
        conv = meta["neighbors"][l].cpu().numpy()
        assert conv.dtype == np.int64 and conv.shape == (len(pts), limit)
        assert np.array_equal(conv, numpy_ops.ball_query_first_k(pts, pts, lens, lens, r, limit))
        if l == L - 1:
            assert tuple(meta["pools"][l].shape) == (0, 1) and tuple(meta["upsamples"][l].shape) == (0, 1)
            break
        nxt, nxt_lens = numpy_ops.voxel_mean_me(pts, lens, 2 * r / cfg.conv_radius)
        assert np.array_equal(meta["pools"][l].cpu().numpy(), numpy_ops.ball_query_first_k(nxt, pts, nxt_lens, lens, r, limit))
        assert np.array_equal(meta["upsamples"][l].cpu().numpy(),
                              numpy_ops.ball_query_first_k(pts, nxt, lens, nxt_lens, 2 * r, limit))
        pts, lens, r = nxt, nxt_lens, 2 * r
    # and the whole forward runs on it (cfg.preprocessor selects the mode inside RegTR)
    from superpoints_registration_b200.model import RegTR
    torch.manual_seed(1)
    np.random.seed(1)
    model = RegTR(cfgs.Config(cfg, preprocessor="gpu_compat")).to(DEV).eval()
    out = model({"src_xyz": [_t(c) for c in clouds[:2]], "tgt_xyz": [_t(c) for c in clouds[2:]]})
    assert tuple(out["pose"].shape) == (2, 3, 4) and torch.isfinite(out["pose"]).all()
