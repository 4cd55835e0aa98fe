"""Pin the oracle (C restatement + NumPy restatement) against the golden vectors the imported reference
produced (tests/golden/*.npz, generator: tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

import oracle
from oracle import numpy_ops, pipeline
from parity import compare_pyramid_stagewise, compare_pyramids_e2e, load_pyramid, pose_error
from superpoints_registration_b200 import config as cfgs
from weights import filled_state

# fp32 feature tolerance: the reference's own torch CPU GEMMs differ from an exactly-rounded (fp64-accumulated)
# evaluation by a few 1e-6 relative to the output scale (SURVEY.md hard part 5: 2.7e-6 max-abs on rms 0.57).
FEAT_RTOL = 2e-5


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


@pytest.mark.parametrize("name,cfg", [("3dmatch", cfgs.threedmatch_config()), ("kitti", cfgs.kitti_config()),
                                      ("modelnet", cfgs.modelnet_config())])
def test_pyramid_against_golden(golden_dir, name, cfg):
    g = _load(golden_dir, f"preprocess_{name}.npz")
    clouds = [g[f"cloud_{i}"] for i in range(int(g["n_clouds"]))]
    ref = load_pyramid(g)

    def nb(q, s, ql, sl, radius, limit):
        idx, mc = oracle.radius_neighbors_batch(q, s, ql, sl, radius, limit)
        return idx[:, :min(mc, limit)]
    assert compare_pyramid_stagewise(ref, cfg, nb, oracle.grid_subsample_batch)
    ours = pipeline.preprocess(cfg, clouds, backend="port")
    rep = compare_pyramids_e2e(ours, ref, cfg)
    assert rep["points_bit_exact"][0] and (rep["levels"] < 2 or rep["points_bit_exact"][1])
    assert rep["flip_rows"] <= max(2, rep["rows"] // 1000), rep


def test_level_plan_matches_reference_pyramid_depth(golden_dir):
    for name, cfg in (("3dmatch", cfgs.threedmatch_config()), ("kitti", cfgs.kitti_config()),
                      ("modelnet", cfgs.modelnet_config())):
        g = _load(golden_dir, f"preprocess_{name}.npz")
        assert len(pipeline.level_plan(cfg["architecture"])) == int(g["n_levels"])


def test_kpconv_layers_against_golden(golden_dir):
    g = _load(golden_dir, "kpconv_layers.npz")
    for tag in g["tags"]:
        tag = str(tag)
        cin = g[f"{tag}_x"].shape[1]
        cout = g[f"{tag}_out"].shape[1]
        wname = f"{tag}.KPConv.weights"
        w = filled_state({wname: (15, cin, cout)}, 100 + cin)[wname]
        out = oracle.kpconv_forward(g[f"{tag}_q"], g[f"{tag}_s"], g[f"{tag}_idx"].astype(np.int64), g[f"{tag}_x"], w,
                                    g[f"{tag}_kp"], float(g[f"{tag}_extent"]))
        ref = g[f"{tag}_out"]
        scale = np.abs(ref).max()
        assert np.abs(out - ref).max() <= FEAT_RTOL * scale, (tag, np.abs(out - ref).max(), scale)
        # the dense NumPy restatement agrees with the C one
        out_np = numpy_ops.kpconv_forward_numpy(g[f"{tag}_q"], g[f"{tag}_s"], g[f"{tag}_idx"].astype(np.int64),
                                                g[f"{tag}_x"], w, g[f"{tag}_kp"], float(g[f"{tag}_extent"]))
        assert np.abs(out_np - out).max() <= 1e-6 * scale


def test_blocks_against_golden(golden_dir):
    g = _load(golden_dir, "blocks.npz")
    y = numpy_ops.instance_norm_lrelu(g["x"], g["lens"])
    d = np.abs(y - g["inorm"])
    o = 0
    for n in g["lens"].tolist():
        # a 2-point cloud is ill-conditioned (x - mean cancels, eps matters): the fp32 reference itself is only
        # accurate to ~1e-4 there; ordinary clouds agree to fp32 rounding
        assert d[o:o + n].max() < (1e-6 if n >= 32 else 2e-4), (n, d[o:o + n].max())
        o += n
    mp = numpy_ops.max_pool(g["x"], g["pool_idx"])
    assert np.array_equal(mp, g["pool_out"])


def test_procrustes_against_golden(golden_dir):
    g = _load(golden_dir, "procrustes.npz")
    for i, kind in enumerate(g["kinds"]):
        a, b, w, T = g[f"a_{i}"], g[f"b_{i}"], g[f"w_{i}"], g[f"T_{i}"]
        ours = numpy_ops.compute_rigid_transform(a, b, None if kind == "unweighted" else w, dtype=np.float64)
        rot, tr = pose_error(ours, T)
        scale = max(1.0, float(np.abs(b).max()))
        # the fp32 reference itself is only accurate to ~1e-5 deg / ~1e-6*scale m against exact arithmetic
        assert rot < 2e-3 and tr < 2e-5 * scale, (kind, rot, tr)
    for p in range(4):
        ours = numpy_ops.compute_rigid_transform(g["batched_a"][p], g["batched_b"][p], g["batched_w"][p], np.float64)
        rot, tr = pose_error(ours, g["batched_T"][p])
        assert rot < 2e-3 and tr < 2e-5


def test_matching_against_golden(golden_dir):
    g = _load(golden_dir, "matching.npz")
    for tag in ("sinkhorn", "argmax"):
        alpha, beta = float(g[f"{tag}_alpha"]), float(g[f"{tag}_beta"])
        for i in range(int(g[f"{tag}_n_pairs"])):
            S, T = g[f"{tag}_src_f_{i}"], g[f"{tag}_tgt_f_{i}"]
            sx, tx = g[f"{tag}_src_xyz_{i}"], g[f"{tag}_tgt_xyz_{i}"]
            corr, attn, val, ind = numpy_ops.dual_softmax_match(S, T)
            assert np.allclose(attn, g[f"{tag}_attn_{i}"], rtol=2e-4, atol=1e-9)
            assert np.array_equal(ind, g[f"{tag}_ind_{i}"])
            assert np.allclose(val, g[f"{tag}_val_{i}"], rtol=2e-4, atol=1e-9)
            if tag == "sinkhorn":
                wt, w = numpy_ops.sinkhorn_weighted_targets(corr, tx, np.log1p(np.exp(alpha)), np.exp(beta), 3)
                pose = numpy_ops.compute_rigid_transform(sx, wt, w, dtype=np.float64)
            else:
                N, M = len(S), len(T)
                if N > M:
                    pose = numpy_ops.compute_rigid_transform(sx[ind], tx, val, dtype=np.float64)
                else:
                    pose = numpy_ops.compute_rigid_transform(sx, tx[ind], val, dtype=np.float64)
            rot, tr = pose_error(pose, g[f"{tag}_pose"][i])
            assert rot < 5e-3 and tr < 5e-5, (tag, i, rot, tr)


def test_refinements_against_golden(golden_dir):
    """The NumPy restatement of the optional refinements against the reference's own softmax_correlation run with
    each of them switched on (tests/golden/make_golden.py:gen_refinements)."""
    g = np.load(os.path.join(golden_dir, "refinements.npz"))
    for i in range(int(g["n_pairs"])):
        S, T, sx, tx = g[f"src_f_{i}"], g[f"tgt_f_{i}"], g[f"src_xyz_{i}"], g[f"tgt_xyz_{i}"]
        n, m = len(S), len(T)
        _, attn, val, ind = numpy_ops.dual_softmax_match(S, T, dtype=np.float64)
        axis = 0 if n > m else 1
        rv, ri = numpy_ops.ratio_test(attn, axis, 2e-4)
        assert np.array_equal(ri, g[f"ratio_ind_{i}"])
        assert np.array_equal(rv > 0, g[f"ratio_val_{i}"] > 0)
        assert np.allclose(rv, g[f"ratio_val_{i}"], rtol=2e-4, atol=1e-9)
        a, b = (sx[ind], tx) if n > m else (sx, tx[ind])
        assert np.array_equal(a, g[f"lgr_src_pts_{i}"]) and np.array_equal(b, g[f"lgr_tgt_pts_{i}"])
        pose0 = numpy_ops.compute_rigid_transform(a, b, val.astype(np.float32))
        pose = numpy_ops.local_global_registration(a, b, val.astype(np.float32), pose0, 0.3, 4)
        rot, tr = pose_error(pose, g["lgr_pose"][i])
        assert rot < 1e-3 and tr < 1e-5, (i, rot, tr)
        # the inlier loop must actually move the estimate on these inputs (25 % gross outliers)
        rot0, tr0 = pose_error(pose0, g["lgr_pose"][i])
        assert rot0 > 0.05 or tr0 > 1e-3


def test_cpu_forward_port_against_golden(golden_dir):
    """The torch-CPU restatement of the network (oracle/pipeline.py:forward) reproduces the reference model's
    encoder features, correspondences and poses on the reference's inputs and weights."""
    import torch
    from superpoints_registration_b200.model import RegTR
    from weights import reference_shapes
    torch.set_num_threads(4)
    for tag, cfg in (("3dmatch", cfgs.threedmatch_config()), ("modelnet", cfgs.modelnet_config())):
        g = _load(golden_dir, f"forward_{tag}.npz")
        own = {k: tuple(v.shape) for k, v in RegTR(cfg).state_dict().items()}
        vals = filled_state(reference_shapes(own, cfg.d_embed), int(g["weight_seed"]))
        sd = {k: torch.from_numpy(g[f"kp::{k}"] if k.endswith("kernel_points") else vals[k]) for k in own}
        B = int(g["n_pairs"])
        # feed the reference's own pyramid ordering by using the reference backend when it is available
        backend = "reference" if oracle.have_ref() else "port"
        out = pipeline.forward(sd, cfg, [g[f"src_{i}"] for i in range(B)], [g[f"tgt_{i}"] for i in range(B)],
                               backend=backend)
        if backend == "reference" and "encoder_out" in g:
            ref = g["encoder_out"]
            assert np.abs(out["encoder_out"] - ref).max() <= 1e-4 * np.abs(ref).max()
        rot, tr = pose_error(out["pose"], g["pose"])
        assert rot.max() < 0.05 and tr.max() < 1e-3, (tag, rot, tr)
        if backend == "reference":
            for i in range(B):
                assert (out["ind"][i] == g[f"ind_{i}"]).mean() > 0.99


def test_cpu_forward_port_on_the_well_conditioned_fixture(golden_dir):
    """tests/golden/forward_wellcond.npz (4-stage architecture, tgt = moved copy of src, damped cross-encoder): the
    reference's end-to-end pose is insensitive to fp32 summation order there, so the CPU restatement is held to
    north_star's 1e-3 deg / 1e-5 m END TO END (pyramid in a different point order, different GEMM blocking)."""
    import torch
    from superpoints_registration_b200.model import RegTR
    from weights import damp_transformer, reference_shapes
    torch.set_num_threads(4)
    g = _load(golden_dir, "forward_wellcond.npz")
    for tag, cfg in (("argmax", cfgs.threedmatch_4stage_config(use_sinkhorn=False)),
                     ("sinkhorn", cfgs.threedmatch_4stage_config())):
        own = {k: tuple(v.shape) for k, v in RegTR(cfg).state_dict().items()}
        vals = damp_transformer(filled_state(reference_shapes(own, cfg.d_embed), int(g["weight_seed"])), float(g["damp"]))
        sd = {k: torch.from_numpy(g[f"kp::{k}"] if k.endswith("kernel_points") else vals[k]) for k in own}
        B = int(g["n_pairs"])
        out = pipeline.forward(sd, cfg, [g[f"src_{i}"] for i in range(B)], [g[f"{tag}_tgt_{i}"] for i in range(B)],
                               backend="port")
        for i in range(B):
            assert out["src_feat"][i].shape[0] == int(g[f"{tag}_n_src_{i}"])
            assert out["tgt_feat"][i].shape[0] == int(g[f"{tag}_n_tgt_{i}"])
        rot, tr = pose_error(out["pose"], g[f"{tag}_pose"])
        # north_star's bar, or twice the reference's own movement under a relabelling of its input when that is larger
        assert rot.max() < max(1e-3, 2 * float(g[f"{tag}_self_noise_rot_deg"])), (tag, rot)
        assert tr.max() < max(1e-5, 2 * float(g[f"{tag}_self_noise_trans"])), (tag, tr)
        if tag == "argmax":
            assert float(g["argmax_self_noise_rot_deg"]) < 5e-4 and float(g["argmax_self_noise_trans"]) < 5e-6   # the reference actually registers this pair: its pose is the ground truth to ~0.6 deg
            rot_gt, tr_gt = pose_error(g[f"{tag}_pose"], np.stack([g["gt_pose"]] * B))
            assert rot_gt.max() < 1.0 and tr_gt.max() < 0.05


def test_sinkhorn_restatement_against_golden(golden_dir):
    """numpy_ops.sinkhorn_log / compute_rigid_transform against the reference's utils/se3_torch.py:sinkhorn and
    compute_rigid_transform_with_sinkhorn (tests/golden/make_golden.py:gen_sinkhorn)."""
    g = _load(golden_dir, "sinkhorn.npz")
    aff, xs, xt = g["affinity"], g["xyz_s"], g["xyz_t"]
    for it in (1, 3, 5):
        for b in range(aff.shape[0]):
            assert np.allclose(numpy_ops.sinkhorn_log(aff[b], it), g[f"log_perm_{it}"][b], rtol=0, atol=2e-5)
    for b in range(aff.shape[0]):
        perm = np.exp(numpy_ops.sinkhorn_log(aff[b].astype(np.float64), 3))
        rows = perm.sum(1, keepdims=True)
        pose = numpy_ops.compute_rigid_transform(xs[b], perm @ xt[b] / (rows + 1e-6), rows[:, 0], dtype=np.float64)
        rot, tr = pose_error(pose, g["transform_3"][b])
        assert rot < 1e-3 and tr < 1e-5, (b, rot, tr)
    assert g["transform_single"].shape == (3, 4)
