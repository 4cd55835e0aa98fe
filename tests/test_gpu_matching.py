"""CUDA superpoint matching, Sinkhorn and pose solve against the reference's golden vectors and the oracle."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import numpy_ops
from parity import load_pyramid, pose_error
from superpoints_registration_b200 import config as cfgs
from superpoints_registration_b200 import ops
from superpoints_registration_b200.model import RegTR
from superpoints_registration_b200.se3 import compute_rigid_transform
from weights import filled_state, reference_shapes

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# north_star tolerance for poses, measured with the chordal metric in fp64 (never se3_compare)
ROT_TOL_DEG = 1e-3
TRANS_TOL = 1e-5


def _t(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def test_procrustes_against_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "procrustes.npz"))
    for i, kind in enumerate(map(str, g["kinds"])):
        a, b, w, T = g[f"a_{i}"], g[f"b_{i}"], g[f"w_{i}"], g[f"T_{i}"]
        ours = compute_rigid_transform(_t(a), _t(b), None if kind == "unweighted" else _t(w)).cpu().numpy()
        exact = numpy_ops.compute_rigid_transform(a, b, None if kind == "unweighted" else w, dtype=np.float64)
        scale = max(1.0, float(np.abs(b).max()))
        rot, tr = pose_error(ours, exact)
        # against exact arithmetic our fp64-moment solver must be at fp32 output rounding
        assert rot < 2e-5 and tr < 4e-7 * scale, (kind, rot, tr)
        rot, tr = pose_error(ours, T)
        # against the fp32 reference the bar is north_star's, scaled by coordinate magnitude for KITTI-size inputs
        assert rot < ROT_TOL_DEG and tr < TRANS_TOL * scale, (kind, rot, tr)
        R = ours[:, :3].astype(np.float64)
        assert abs(np.linalg.det(R) - 1) < 1e-5 and np.abs(R @ R.T - np.eye(3)).max() < 1e-5
    T = compute_rigid_transform(_t(g["batched_a"]), _t(g["batched_b"]), _t(g["batched_w"])).cpu().numpy()
    assert T.shape == (4, 3, 4)
    rot, tr = pose_error(T, g["batched_T"])
    assert rot.max() < ROT_TOL_DEG and tr.max() < TRANS_TOL


def test_procrustes_assertions_match_reference_behaviour():
    a = torch.rand(10, 3, device=DEV)
    with pytest.raises(AssertionError):
        compute_rigid_transform(a, a[:5])
    with pytest.raises(AssertionError):
        compute_rigid_transform(a, a, torch.full((10,), 2.0, device=DEV))  # weights outside [0, 1]


def test_procrustes_identity_and_degenerate_inputs():
    rng = np.random.default_rng(0)
    a = rng.normal(size=(50, 3)).astype(np.float32)
    T = compute_rigid_transform(_t(a), _t(a), _t(np.full(50, 0.5, np.float32))).cpu().numpy()
    assert np.abs(T - np.concatenate([np.eye(3), np.zeros((3, 1))], 1)).max() < 1e-6
    # all-zero weights (sum clamped to 1e-6 in the reference): must stay finite and orthonormal
    T0 = compute_rigid_transform(_t(a), _t(a + 1), _t(np.zeros(50, np.float32))).cpu().numpy()
    assert np.isfinite(T0).all()
    # collinear points: rank-1 covariance
    line = np.outer(np.linspace(-1, 1, 30), [1, 2, 3]).astype(np.float32)
    T1 = compute_rigid_transform(_t(line), _t(line), None).cpu().numpy()
    R = T1[:, :3].astype(np.float64)
    assert np.isfinite(T1).all() and abs(np.linalg.det(R) - 1) < 1e-5


def _model(cfg, alpha, beta):
    m = RegTR(cfg).to(DEV)
    m.alpha.data.fill_(float(alpha))
    m.beta.data.fill_(float(beta))
    return m


@pytest.mark.parametrize("tag,cfg", [("sinkhorn", cfgs.threedmatch_config()), ("argmax", cfgs.kitti_config())])
def test_softmax_correlation_against_golden(golden_dir, tag, cfg):
    g = np.load(os.path.join(golden_dir, "matching.npz"))
    P = int(g[f"{tag}_n_pairs"])
    model = _model(cfg, g[f"{tag}_alpha"], g[f"{tag}_beta"])
    src_f = [_t(g[f"{tag}_src_f_{i}"])[None] for i in range(P)]
    tgt_f = [_t(g[f"{tag}_tgt_f_{i}"])[None] for i in range(P)]
    src_x = [_t(g[f"{tag}_src_xyz_{i}"]) for i in range(P)]
    tgt_x = [_t(g[f"{tag}_tgt_xyz_{i}"]) for i in range(P)]
    pose, attn, val, ind, sp, tp = model.softmax_correlation(src_f, tgt_f, src_x, tgt_x, None, None)
    for i in range(P):
        assert np.array_equal(ind[i].cpu().numpy(), g[f"{tag}_ind_{i}"])           # correspondences: exact
        assert np.allclose(val[i].cpu().numpy(), g[f"{tag}_val_{i}"], rtol=2e-4, atol=1e-9)
        assert tuple(attn[i].shape) == (1,) + g[f"{tag}_attn_{i}"].shape
        assert np.allclose(attn[i][0].cpu().numpy(), g[f"{tag}_attn_{i}"], rtol=2e-4, atol=1e-9)
    rot, tr = pose_error(pose.cpu().numpy(), g[f"{tag}_pose"])
    assert rot.max() < ROT_TOL_DEG and tr.max() < TRANS_TOL, (rot, tr)


REFINEMENTS = {
    "ratio": dict(use_ratio_test=True, lowe_thres=2e-4),
    "median": dict(threshold_corr=True),
    "overlap": dict(remove_outliers_overlap=True),
    "overlap_w": dict(remove_outliers_overlap=True, use_overlap_as_weights=True),
    "lgr": dict(use_lgr=True, acceptance_radius=0.3, num_refinement_steps=4),
    "topk": dict(remove_points_from_val=True, val_threshold=0.25),
    "combo": dict(use_ratio_test=True, lowe_thres=5e-4, threshold_corr=True, remove_outliers_overlap=True, use_lgr=True,
                  acceptance_radius=0.3, num_refinement_steps=3),
}


def _refinement_inputs(g):
    P = int(g["n_pairs"])
    return ([_t(g[f"src_f_{i}"])[None] for i in range(P)], [_t(g[f"tgt_f_{i}"])[None] for i in range(P)],
            [_t(g[f"src_xyz_{i}"]) for i in range(P)], [_t(g[f"tgt_xyz_{i}"]) for i in range(P)],
            [_t(g[f"src_ov_{i}"]) for i in range(P)], [_t(g[f"tgt_ov_{i}"]) for i in range(P)])


@pytest.mark.parametrize("tag", sorted(REFINEMENTS))
def test_optional_refinements_against_golden(golden_dir, tag):
    """Each optional refinement of softmax_correlation (qk_regtr_full.py:370-398,465-502) against the reference run
    with the same switches (same settings as tests/golden/make_golden.py:REFINEMENT_VARIANTS)."""
    g = np.load(os.path.join(golden_dir, "refinements.npz"))
    P = int(g["n_pairs"])
    model = _model(cfgs.kitti_config(**REFINEMENTS[tag]), 0.0, 0.0)
    pose, attn, val, ind, sp, tp = model.softmax_correlation(*_refinement_inputs(g))
    for i in range(P):
        assert np.array_equal(ind[i].cpu().numpy(), g[f"{tag}_ind_{i}"]), (tag, i)
        v, v_ref = val[i].cpu().numpy(), g[f"{tag}_val_{i}"]
        assert np.array_equal(v > 0, v_ref > 0), (tag, i)          # the same rows are rejected
        assert np.allclose(v, v_ref, rtol=2e-4, atol=1e-9)
        assert np.array_equal(sp[i].cpu().numpy(), g[f"{tag}_src_pts_{i}"])
        assert np.array_equal(tp[i].cpu().numpy(), g[f"{tag}_tgt_pts_{i}"])
    rot, tr = pose_error(pose.cpu().numpy(), g[f"{tag}_pose"])
    assert rot.max() < ROT_TOL_DEG and tr.max() < TRANS_TOL, (tag, rot, tr)


def test_refinement_switches_the_reference_cannot_run_are_refused():
    for bad in (dict(use_attn_affinity=True), dict(use_corr_affinity=True), dict(use_overlap_as_weights=True)):
        with pytest.raises(NotImplementedError):
            RegTR(cfgs.kitti_config(**bad))
    with pytest.raises(NotImplementedError):
        RegTR(cfgs.threedmatch_config(use_lgr=True))
    model = _model(cfgs.kitti_config(use_ratio_test=True), 0.0, 0.0)
    one = [torch.randn(1, 1, 256, device=DEV)]
    with pytest.raises(RuntimeError):                                   # torch.topk(k=2) over one entry
        model.softmax_correlation(one, one, [torch.zeros(1, 3, device=DEV)], [torch.zeros(1, 3, device=DEV)])


def test_ransac_against_oracle_with_given_draws(golden_dir):
    """RegTR.ransac (:400-421) draws from the CUDA generator, so the comparison feeds both sides the same draws."""
    g = np.load(os.path.join(golden_dir, "refinements.npz"))
    P = int(g["n_pairs"])
    rng = np.random.default_rng(3)
    a = [g[f"lgr_src_pts_{i}"] for i in range(P)]
    b = [g[f"lgr_tgt_pts_{i}"] for i in range(P)]
    w = [np.clip(g[f"lgr_val_{i}"], 0, 1) for i in range(P)]
    H, S = 40, 12
    idx = np.stack([rng.integers(0, len(a[i]), size=(H, S)) for i in range(P)])
    offs = np.concatenate([[0], np.cumsum([len(x) for x in a])]).astype(np.int32)
    pose, loss, best = ops.ransac(_t(np.concatenate(a)), _t(np.concatenate(b)), _t(np.concatenate(w)), _t(offs),
                                  _t(idx))
    for i in range(P):
        T_o, loss_o = numpy_ops.ransac(a[i], b[i], w[i], idx[i], dtype=np.float64)
        assert np.allclose(loss[i].cpu().numpy(), loss_o, rtol=1e-4, atol=1e-6)
        assert int(best[i]) == int(np.argmin(loss_o))                  # first strict minimum
        rot, tr = pose_error(pose[i].cpu().numpy(), T_o)
        assert rot < ROT_TOL_DEG and tr < TRANS_TOL
    # through the model: hypotheses drawn on the device from a seeded generator, result reproducible and at least
    # as good (mean residual) as the plain weighted solve restricted to the hypotheses' own measure
    gen = torch.Generator(device=DEV)
    model = _model(cfgs.kitti_config(use_ransac=True), 0.0, 0.0)
    model.ransac_generator, model.ransac_hypotheses, model.ransac_sample_size = gen, 64, 16
    gen.manual_seed(11)
    p1 = model.softmax_correlation(*_refinement_inputs(g))[0]
    gen.manual_seed(11)
    p2 = model.softmax_correlation(*_refinement_inputs(g))[0]
    assert torch.equal(p1, p2) and tuple(p1.shape) == (P, 3, 4) and torch.isfinite(p1).all()


def test_matching_matches_oracle_on_ragged_batch():
    rng = np.random.default_rng(5)
    shapes = [(300, 17), (16, 500), (1, 1), (257, 256)]
    S = [rng.normal(size=(n, 256)).astype(np.float32) for n, _ in shapes]
    T = [rng.normal(size=(m, 256)).astype(np.float32) for _, m in shapes]
    pairs = ops.PackedPairs([n for n, _ in shapes], [m for _, m in shapes], DEV)
    corr, attn, val, ind = ops.dual_softmax_match(_t(np.concatenate(S)), _t(np.concatenate(T)), pairs, want_attn=True)
    for p, (n, m) in enumerate(shapes):
        c_o, a_o, v_o, i_o = numpy_ops.dual_softmax_match(S[p], T[p], dtype=np.float64)
        c = corr[pairs.h_co[p]:pairs.h_co[p + 1]].view(n, m).cpu().numpy()
        assert np.abs(c - c_o).max() < 1e-4
        a = attn[pairs.h_co[p]:pairs.h_co[p + 1]].view(n, m).cpu().numpy()
        assert np.allclose(a, a_o, rtol=1e-3, atol=1e-12)
        got_i = ind[pairs.h_oo[p]:pairs.h_oo[p + 1]].cpu().numpy()
        # random features: accept a differing argmax only when the two candidates are within fp32 noise
        diff = got_i != i_o
        if diff.any():
            col = n > m
            for k in np.nonzero(diff)[0]:
                x, y = (a_o[got_i[k], k], a_o[i_o[k], k]) if col else (a_o[k, got_i[k]], a_o[k, i_o[k]])
                assert abs(x - y) <= 1e-5 * abs(y)


def test_matching_with_non_finite_features_keeps_indices_in_range():
    """torch.max semantics (qk_regtr_full.py:468,576): a row / column of NaN attention yields a NaN value and an
    IN-RANGE index (the first NaN), so the gather that follows never leaves the cloud."""
    rng = np.random.default_rng(8)
    shapes = [(40, 60), (70, 30)]
    S = [rng.normal(size=(n, 256)).astype(np.float32) for n, _ in shapes]
    T = [rng.normal(size=(m, 256)).astype(np.float32) for _, m in shapes]
    S[0][3] = np.nan                                   # N <= M: source row 3 is all NaN
    T[1][5] = np.inf                                   # N > M: target column 5 has infinite correlations
    pairs = ops.PackedPairs([n for n, _ in shapes], [m for _, m in shapes], DEV)
    corr, attn, val, ind = ops.dual_softmax_match(_t(np.concatenate(S)), _t(np.concatenate(T)), pairs, want_attn=True)
    ind_h, val_h = ind.cpu().numpy(), val.cpu().numpy()
    for p, (n, m) in enumerate(shapes):
        got = ind_h[pairs.h_oo[p]:pairs.h_oo[p + 1]]
        assert got.min() >= 0 and got.max() < (m if n <= m else n), (p, got.min(), got.max())
    assert np.isnan(val_h[3]) and ind_h[3] == 0        # first index, like torch.max
    want = torch.max(attn[:40 * 60].view(40, 60), dim=1)
    ok = ~torch.isnan(want.values)
    assert torch.equal(ind[:40][ok], want.indices[ok])
    pts = _t(rng.normal(size=(200, 3)).astype(np.float32))
    base = torch.zeros(ind.shape[0], dtype=torch.int32, device=DEV)
    out = ops.gather_rows3(pts, torch.full_like(ind, 2 ** 31 - 1), base)    # a sentinel index is clamped, not read
    assert torch.equal(out, pts[-1].expand_as(out))


def test_sinkhorn_surface_against_golden(golden_dir):
    """The reference's callable surface utils/se3_torch.py:166-239 (`sinkhorn`, `compute_rigid_transform_with_sinkhorn`,
    callers qk_regtr_full.py:536,647) on its own outputs."""
    from superpoints_registration_b200 import compute_rigid_transform_with_sinkhorn, sinkhorn
    g = np.load(os.path.join(golden_dir, "sinkhorn.npz"))
    aff, xs, xt = _t(g["affinity"]), _t(g["xyz_s"]), _t(g["xyz_t"])
    for it in (1, 3, 5):
        got = sinkhorn(aff, n_iters=it, slack=True)
        assert tuple(got.shape) == tuple(aff.shape)
        assert np.allclose(got.cpu().numpy(), g[f"log_perm_{it}"], rtol=0, atol=2e-5)
    T = compute_rigid_transform_with_sinkhorn(xs, xt, aff, True, 3)
    rot, tr = pose_error(T.cpu().numpy(), g["transform_3"])
    assert tuple(T.shape) == (3, 3, 4) and rot.max() < ROT_TOL_DEG and tr.max() < TRANS_TOL, (rot, tr)
    T1 = compute_rigid_transform_with_sinkhorn(xs[:1], xt[:1], aff[:1], True, 3)
    assert tuple(T1.shape) == (3, 4)                                   # squeezed like the reference (:231)
    rot, tr = pose_error(T1.cpu().numpy(), g["transform_single"])
    assert rot < ROT_TOL_DEG and tr < TRANS_TOL


def _v64(corr, axis):
    a = numpy_ops._softmax(corr, 0) * numpy_ops._softmax(corr, 1)
    return a.max(axis=axis)


def test_full_forward_against_golden(golden_dir):
    """End to end on the reference's inputs and weights: pyramid, encoder, transformer, matching, pose.
    Stage-wise pose parity is asserted at north_star's bar; end to end the Sinkhorn fixture (ill-conditioned by its filler
    weights) is held to 0.05 deg, the arg-max fixture to the plain bar unless a correspondence flips in a provable tie."""
    for tag, cfg in (("3dmatch", cfgs.threedmatch_config()), ("modelnet", cfgs.modelnet_config())):
        g = np.load(os.path.join(golden_dir, f"forward_{tag}.npz"))
        model = RegTR(cfg).to(DEV).eval()
        own = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        vals = filled_state(reference_shapes(own, cfg.d_embed), int(g["weight_seed"]))
        sd = {k: (_t(g[f"kp::{k}"]) if k.endswith("kernel_points") else _t(vals[k])) for k in own}
        model.load_state_dict(sd)
        B = int(g["n_pairs"])
        # (1) stage-wise: the reference's conditioned features through our matching + pose
        src_f = [_t(g[f"src_feat_{i}"])[None] for i in range(B)]
        tgt_f = [_t(g[f"tgt_feat_{i}"])[None] for i in range(B)]
        meta = load_pyramid(g, "meta_")
        lens = meta["stack_lengths"][-1].tolist()
        pts = _t(meta["points"][-1])
        chunks = torch.split(pts, lens)
        pose, _, val, ind, _, _ = model.softmax_correlation(src_f, tgt_f, chunks[:B], chunks[B:], None, None)
        pose = pose.cpu().numpy()
        # These features come from filler weights and are nearly uniform across superpoints, so the soft assignment
        # is ill-conditioned: the reference's own fp32 result sits up to ~2e-3 deg from the exact-arithmetic (fp64)
        # evaluation of the same formulas.  The bar is therefore north_star's tolerance, widened to the
        # reference's measured self-noise when that is larger; well-conditioned inputs (matching.npz) use the
        # plain tolerance in test_softmax_correlation_against_golden.
        sp_alpha, e_beta = float(np.log1p(np.exp(vals["alpha"]))), float(np.exp(vals["beta"]))
        for i in range(B):
            S, T = g[f"src_feat_{i}"], g[f"tgt_feat_{i}"]
            sx, tx = chunks[i].cpu().numpy(), chunks[B + i].cpu().numpy()
            corr, _, _, ind_o = numpy_ops.dual_softmax_match(S, T, dtype=np.float64)
            if cfg.use_sinkhorn:
                wt, w = numpy_ops.sinkhorn_weighted_targets(corr, tx, sp_alpha, e_beta, cfg.sinkhorn_itr, dtype=np.float64)
                exact = numpy_ops.compute_rigid_transform(sx, wt, w, dtype=np.float64)
            elif len(S) > len(T):
                exact = numpy_ops.compute_rigid_transform(sx[ind_o], tx, _v64(corr, 0), dtype=np.float64)
            else:
                exact = numpy_ops.compute_rigid_transform(sx, tx[ind_o], _v64(corr, 1), dtype=np.float64)
            noise_rot, noise_tr = pose_error(exact, g["pose"][i])
            rot, tr = pose_error(pose[i], exact)
            rot_ref, tr_ref = pose_error(pose[i], g["pose"][i])
            print(f"[{tag}] pair {i}: ours vs exact {rot:.2e} deg / {tr:.2e} m; reference vs exact {noise_rot:.2e} / "
                  f"{noise_tr:.2e}; ours vs reference {rot_ref:.2e} / {tr_ref:.2e}")
            assert rot <= max(ROT_TOL_DEG, 4 * noise_rot) and tr <= max(TRANS_TOL, 4 * noise_tr), (tag, i, rot, tr)
            assert rot_ref <= max(ROT_TOL_DEG, 5 * noise_rot) and tr_ref <= max(TRANS_TOL, 5 * noise_tr)
            assert np.array_equal(ind[i].cpu().numpy(), g[f"ind_{i}"])
        # (2) end to end from raw clouds
        batch = {"src_xyz": [_t(g[f"src_{i}"]) for i in range(B)], "tgt_xyz": [_t(g[f"tgt_{i}"]) for i in range(B)]}
        out = model(batch)
        assert tuple(out["pose"].shape) == (B, 3, 4)
        for key in ("pose", "attn", "src_feat", "tgt_feat", "src_kp", "tgt_kp", "src_corr", "tgt_corr", "src_overlap",
                    "tgt_overlap", "overlap_prob_list", "ind_list"):
            assert key in out
        rot, tr = pose_error(out["pose"].cpu().numpy(), g["pose"])
        print(f"[{tag}] end-to-end pose difference vs reference: rot {rot.max():.2e} deg, trans {tr.max():.2e} m")
        if cfg.use_sinkhorn:
            # ill-conditioned by construction of the fixture (see above): the reference's own fp32 evaluation is
            # ~2e-3 deg from exact arithmetic here; the well-conditioned end-to-end case is
            # test_full_forward_end_to_end_pose_on_the_well_conditioned_fixture
            assert rot.max() < 0.05 and tr.max() < 1e-3, (tag, rot, tr)
        else:
            # arg-max correspondences: with these filler weights the features are nearly uniform across superpoints, so a
            # few rows have two candidates whose scores differ by less than the fp32 noise of the encoder.  The pose is
            # held to the plain tolerance when every correspondence equals the reference's; a differing row must be a
            # provable near-tie of the reference's OWN scores (fp64 restatement on the golden features), at most 1 % of
            # the rows may differ, and the pose must then equal the fp64 restatement evaluated on OUR features.
            # (end to end our superpoints come in canonical order, the reference's in its hash-map order: correspondences
            # are compared through the coordinates of the points they join)
            from scipy.spatial import cKDTree
            flips = 0
            for i in range(B):
                S, T = g[f"src_feat_{i}"], g[f"tgt_feat_{i}"]
                ref_src, ref_tgt = chunks[i].cpu().numpy(), chunks[B + i].cpu().numpy()
                sx, tx = out["src_kp"][i].cpu().numpy(), out["tgt_kp"][i].cpu().numpy()
                ds, s_of = cKDTree(ref_src).query(sx)      # our source row -> the reference's row of the same point
                dt, t_of = cKDTree(ref_tgt).query(tx)
                assert ds.max() < 1e-5 and dt.max() < 1e-5 and len(sx) == len(ref_src) and len(tx) == len(ref_tgt)
                ours_ind, ref_ind = out["ind_list"][i].cpu().numpy(), g[f"ind_{i}"]
                if len(S) > len(T):    # one correspondence per target point: rows = targets, candidates = sources
                    picks_ours, picks_ref = s_of[ours_ind], ref_ind[t_of]
                    rows = t_of
                else:
                    picks_ours, picks_ref = t_of[ours_ind], ref_ind[s_of]
                    rows = s_of
                diff = np.nonzero(picks_ours != picks_ref)[0]
                flips += len(diff)
                assert len(diff) <= max(1, len(ref_ind) // 100), (tag, i, len(diff))
                if len(diff):
                    _, attn64, _, _ = numpy_ops.dual_softmax_match(S, T, dtype=np.float64)
                    a = attn64.T if len(S) > len(T) else attn64        # row = the point that picks, column = candidate
                    r = rows[diff]
                    gap = np.abs(a[r, picks_ours[diff]] - a[r, picks_ref[diff]]) / a[r, picks_ref[diff]]
                    print(f"[{tag}] pair {i}: {len(diff)} of {len(ref_ind)} correspondences differ, score gaps {gap}")
                    assert gap.max() < 2e-4, (tag, i, gap)
                    last = lambda f: f.reshape(-1, f.shape[-2], f.shape[-1])[-1].cpu().numpy()   # features of the last layer
                    So, To = last(out["src_feat"][i]), last(out["tgt_feat"][i])
                    # the fp64 restatement on our features, with OUR choice in the tied rows
                    _, attn_o, _, _ = numpy_ops.dual_softmax_match(So, To, dtype=np.float64)
                    if len(So) > len(To):
                        w = attn_o[ours_ind, np.arange(len(To))]
                        exact = numpy_ops.compute_rigid_transform(sx[ours_ind], tx, w, dtype=np.float64)
                    else:
                        w = attn_o[np.arange(len(So)), ours_ind]
                        exact = numpy_ops.compute_rigid_transform(sx, tx[ours_ind], w, dtype=np.float64)
                    r1, t1 = pose_error(out["pose"][i].cpu().numpy(), exact)
                    assert r1 < ROT_TOL_DEG and t1 < TRANS_TOL, (tag, i, r1, t1)
            if flips == 0:
                assert rot.max() < ROT_TOL_DEG and tr.max() < TRANS_TOL, (tag, rot, tr)


@pytest.mark.parametrize("tag", ["argmax", "sinkhorn"])
def test_full_forward_end_to_end_pose_on_the_well_conditioned_fixture(golden_dir, tag):
    """END-TO-END pose parity at north_star's tolerance (1e-3 deg, 1e-5 m) against the unmodified reference, raw
    clouds in, pose out, on the 4-stage architecture of the bench (tests/golden/make_golden.py:gen_forward_wellcond):
    pyramid, 11 encoder blocks (tcgen05 KPConv and GEMMs), 6 cross-encoder layers, matching and pose solve all on
    our kernels."""
    from weights import damp_transformer
    g = np.load(os.path.join(golden_dir, "forward_wellcond.npz"))
    cfg = cfgs.threedmatch_4stage_config(use_sinkhorn=False) if tag == "argmax" else cfgs.threedmatch_4stage_config()
    model = RegTR(cfg).to(DEV).eval()
    own = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    vals = damp_transformer(filled_state(reference_shapes(own, cfg.d_embed), int(g["weight_seed"])), float(g["damp"]))
    model.load_state_dict({k: (_t(g[f"kp::{k}"]) if k.endswith("kernel_points") else _t(vals[k])) for k in own})
    B = int(g["n_pairs"])
    batch = {"src_xyz": [_t(g[f"src_{i}"]) for i in range(B)], "tgt_xyz": [_t(g[f"{tag}_tgt_{i}"]) for i in range(B)]}
    out = model(batch)
    for i in range(B):
        assert out["src_feat"][i].shape[1] == int(g[f"{tag}_n_src_{i}"])         # same superpoints
        assert out["tgt_feat"][i].shape[1] == int(g[f"{tag}_n_tgt_{i}"])
    rot, tr = pose_error(out["pose"].cpu().numpy(), g[f"{tag}_pose"])
    # The bar is north_star's, unless the reference ITSELF moves by more than half of it when its input points are
    # merely relabelled (recorded in the fixture): argmax 3.9e-5 deg / 1.3e-6 m -> the plain tolerance applies;
    # Sinkhorn (near-uniform assignment with untrained weights) 2.2e-4 deg / 8.9e-6 m -> twice that self-noise.
    rot_bar = max(ROT_TOL_DEG, 2 * float(g[f"{tag}_self_noise_rot_deg"]))
    tr_bar = max(TRANS_TOL, 2 * float(g[f"{tag}_self_noise_trans"]))
    print(f"[wellcond/{tag}] end-to-end pose vs reference: rot {rot.max():.2e} deg, trans {tr.max():.2e} m "
          f"(bars {rot_bar:.1e} / {tr_bar:.1e})")
    if tag == "argmax":
        assert rot_bar == ROT_TOL_DEG and tr_bar == TRANS_TOL
    assert rot.max() < rot_bar and tr.max() < tr_bar, (tag, rot, tr)
    # and the plain route (fp32 SIMT KPConv, unfused blocks, padded nn.MultiheadAttention) agrees to the same bar
    model.packed_transformer = False
    for m in model.modules():
        if m.__class__.__name__ == "KPConv":
            m.mode = 0
    plain = model(dict(batch))
    rot, tr = pose_error(plain["pose"].cpu().numpy(), g[f"{tag}_pose"])
    assert rot.max() < rot_bar and tr.max() < tr_bar, (tag, "plain", rot, tr)


@pytest.mark.parametrize("kind,cfg,kw", [("3dlomatch", cfgs.threedmatch_config(), dict(n_points=4000)),
                                         ("kitti", cfgs.kitti_config(), dict(n_points=6000)),
                                         ("3dmatch", cfgs.threedmatch_4stage_config(), dict(n_points=6000))])
def test_forward_lomatch_and_kitti_shapes_against_oracle(kind, cfg, kw):
    """BASELINE configs[3] (low-overlap 3DLoMatch-shape pairs, Sinkhorn) and configs[4] (KITTI-shape scans, 4-stage,
    wide neighbourhoods, argmax + Procrustes) through the whole CUDA forward against the CPU restatement of the
    reference network on the same clouds and weights (reduced point counts so that the CPU side runs in seconds)."""
    from oracle import pipeline
    from superpoints_registration_b200 import synthetic
    torch.manual_seed(7)
    np.random.seed(7)            # the kernel-point dispositions are optimised from NumPy's global stream
    model = RegTR(cfg).to(DEV).eval()
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    data = synthetic.make_batch(kind, 2, seed=3, **kw)
    B = len(data["src_xyz"])
    out = model({"src_xyz": [_t(c) for c in data["src_xyz"]], "tgt_xyz": [_t(c) for c in data["tgt_xyz"]]})
    exact = pipeline.forward(sd, cfg, data["src_xyz"], data["tgt_xyz"], backend="port")
    sp_alpha, e_beta = float(np.log1p(np.exp(sd["alpha"].item()))), float(np.exp(sd["beta"].item()))
    for i in range(B):
        S, T = out["src_feat"][i][0].cpu().numpy(), out["tgt_feat"][i][0].cpu().numpy()
        assert S.shape == exact["src_feat"][i].shape and T.shape == exact["tgt_feat"][i].shape   # same superpoints
        ref_scale = max(np.abs(exact["src_feat"][i]).max(), np.abs(exact["tgt_feat"][i]).max())
        assert np.abs(S - exact["src_feat"][i]).max() <= 2e-3 * ref_scale                        # conditioned features
        assert np.abs(T - exact["tgt_feat"][i]).max() <= 2e-3 * ref_scale
        # stage-wise: OUR features through the fp64 restatement of matching + pose
        sx, tx = out["src_kp"][i].cpu().numpy(), out["tgt_kp"][i].cpu().numpy()
        corr, _, _, ind_o = numpy_ops.dual_softmax_match(S, T, dtype=np.float64)
        if cfg.use_sinkhorn:
            # Untrained weights give a nearly uniform soft assignment, so the final pose is ill-conditioned (the fp32
            # NumPy evaluation of the same formulas sits 3e-4 ... 3e-3 deg from the fp64 one, depending on the draw).
            # The stages are therefore checked one by one on OUR features: Sinkhorn-weighted targets against fp64,
            # then the pose solve on our own weighted targets against fp64; the end-to-end number is printed.
            pairs = ops.PackedPairs([len(S)], [len(T)], DEV)
            corr_g, _, _, _ = ops.dual_softmax_match(_t(S), _t(T), pairs)
            wt_g, w_g = ops.sinkhorn_weighted_targets(corr_g, pairs, _t(tx), sp_alpha, e_beta, int(cfg.sinkhorn_itr),
                                                      bool(cfg.slack))
            wt, w = numpy_ops.sinkhorn_weighted_targets(corr, tx, sp_alpha, e_beta, cfg.sinkhorn_itr, dtype=np.float64)
            assert np.abs(w_g.cpu().numpy() - w).max() <= 2e-4 * np.abs(w).max()
            assert np.abs(wt_g.cpu().numpy() - wt).max() <= 2e-4 * max(1.0, np.abs(tx).max())
            pose_g = ops.weighted_procrustes(_t(sx), wt_g, w_g, pairs.so)[0]
            assert (pose_g - out["pose"][i]).abs().max().item() <= 1e-6        # the batched forward took the same route
            wt_h, w_h = wt_g.cpu().numpy(), w_g.cpu().numpy()
            pose64 = numpy_ops.compute_rigid_transform(sx, wt_h, w_h, dtype=np.float64)
            pose32 = numpy_ops.compute_rigid_transform(sx, wt_h, w_h)
            e2e_rot, e2e_tr = pose_error(out["pose"][i].cpu().numpy(),
                                         numpy_ops.compute_rigid_transform(sx, wt, w, dtype=np.float64))
            print(f"[{kind}] pair {i}: end-to-end pose vs all-fp64 matching + Sinkhorn + solve: {e2e_rot:.2e} deg / {e2e_tr:.2e} m")
        else:
            # untrained weights make the argmax a lottery between nearly equal attention values: the pose solve is
            # checked on OUR correspondences and weights, the correspondences and weights themselves below
            got = out["ind_list"][i].cpu().numpy()
            val = out["overlap_prob_list"][i].cpu().numpy()
            a, b = (sx[got], tx) if len(S) > len(T) else (sx, tx[got])
            pose64 = numpy_ops.compute_rigid_transform(a, b, val.astype(np.float64), dtype=np.float64)
            pose32 = numpy_ops.compute_rigid_transform(a, b, val)
            v_o = _v64(corr, 0 if len(S) > len(T) else 1)
            same = got == ind_o
            assert same.mean() > 0.99
            assert np.allclose(val[same], v_o[same], rtol=5e-4, atol=1e-12)
            # where the argmax differs, the two candidates are within fp32 noise of each other
            attn = numpy_ops._softmax(corr, 0) * numpy_ops._softmax(corr, 1)
            for k in np.nonzero(~same)[0]:
                x, y = (attn[got[k], k], attn[ind_o[k], k]) if len(S) > len(T) else (attn[k, got[k]], attn[k, ind_o[k]])
                assert abs(x - y) <= 1e-4 * abs(y)
        noise_rot, noise_tr = pose_error(pose32, pose64)           # what fp32 evaluation of the same formulas costs
        rot, tr = pose_error(out["pose"][i].cpu().numpy(), pose64)
        scale = max(1.0, float(np.abs(tx).max()))
        print(f"[{kind}] pair {i}: N={len(S)} M={len(T)} ours vs exact {rot:.2e} deg / {tr:.2e} m (fp32 noise "
              f"{noise_rot:.2e} / {noise_tr:.2e})")
        assert rot <= max(ROT_TOL_DEG, 4 * noise_rot) and tr <= max(TRANS_TOL * scale, 4 * noise_tr), (kind, i, rot, tr)
    assert tuple(out["pose"].shape) == (B, 3, 4) and torch.isfinite(out["pose"]).all()
