"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference):
    python tests/golden/make_golden.py
What runs: the reference's own PyTorch modules imported from /root/reference/src (oracle/ref_torch.py) with
its own C++ preprocessing core compiled in place (oracle/_ref).  Nothing from superpoints_registration_b200 or
from the C/NumPy restatement is involved, so these files pin both our oracle and our CUDA path.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))  # tests/parity.py

import oracle  # noqa: E402
from oracle import ref_torch  # noqa: E402
from superpoints_registration_b200 import synthetic  # noqa: E402  (data generator only)
from weights import damp_transformer, filled_state  # noqa: E402

torch.set_num_threads(8)


def t2n(t):
    return t.detach().cpu().numpy()


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def load_weights(model, seed):
    sd = model.state_dict()
    shapes = {k: tuple(v.shape) for k, v in sd.items()}
    vals = filled_state(shapes, seed)
    new = {k: (torch.from_numpy(vals[k]) if k in vals else v) for k, v in sd.items()}
    model.load_state_dict(new)
    return {k: t2n(v) for k, v in sd.items() if k.endswith("kernel_points")}


def pyramid_arrays(meta, prefix=""):
    out = {}
    for key in ("points", "neighbors", "pools", "upsamples", "stack_lengths"):
        for l, t in enumerate(meta[key]):
            a = t2n(t)
            if a.dtype == np.int64 and key != "stack_lengths":
                a = a.astype(np.int32)
            out[f"{prefix}{key}_{l}"] = a
    out[f"{prefix}n_levels"] = np.asarray(len(meta["points"]))
    return out


def gen_preprocess(ref):
    cases = [
        ("3dmatch", "qk_regtr_full_3dmatch.yaml", synthetic.make_batch("3dmatch", 1, seed=11, n_points=1200)),
        ("kitti", "qk_regtr_full_kitti.yaml", synthetic.make_batch("kitti", 1, seed=12, n_points=800)),
        ("modelnet", "qk_regtr_full_modelnet.yaml", synthetic.make_batch("modelnet", 2, seed=13)),
    ]
    for name, yaml_name, data in cases:
        cfg = ref_torch.load_cfg(yaml_name)
        pre = ref.kpconv.Preprocessor(cfg)
        clouds = [torch.from_numpy(c) for c in data["src_xyz"] + data["tgt_xyz"]]
        meta = pre(clouds)
        arrays = pyramid_arrays(meta)
        arrays["n_clouds"] = np.asarray(len(clouds))
        for i, c in enumerate(clouds):
            arrays[f"cloud_{i}"] = t2n(c)
        save(f"preprocess_{name}.npz", **arrays)


def gen_kpconv(ref):
    """Single KPConv layers (reference module) on a reference pyramid."""
    cfg = ref_torch.load_cfg("qk_regtr_full_3dmatch.yaml")
    data = synthetic.make_batch("3dmatch", 1, seed=21, n_points=300)
    clouds = [torch.from_numpy(c) for c in data["src_xyz"] + data["tgt_xyz"]]
    meta = ref.kpconv.Preprocessor(cfg)(clouds)
    arrays = {}
    torch.manual_seed(5)
    np.random.seed(5)
    specs = [  # (tag, level, strided, cin, cout)
        ("c1", 0, False, 1, 64), ("c32", 0, False, 32, 32), ("c32s", 0, True, 32, 32), ("c64", 1, False, 64, 64),
        ("c128", 2, False, 128, 128),
    ]
    for tag, lvl, strided, cin, cout in specs:
        r = cfg.first_subsampling_dl * cfg.conv_radius * 2 ** lvl
        extent = r * cfg.KP_extent / cfg.conv_radius
        with ref_torch.ref_cwd():
            conv = ref.kpconv_blocks.KPConv(cfg.num_kernel_points, 3, cin, cout, extent, r)
        # weights come from the shared deterministic filler (kept out of the fixture): seed = 100 + cin
        wname = f"{tag}.KPConv.weights"
        conv.weights.data.copy_(torch.from_numpy(filled_state({wname: (15, cin, cout)}, 100 + cin)[wname]))
        if strided:
            q, s, idx = meta["points"][lvl + 1], meta["points"][lvl], meta["pools"][lvl]
        else:
            q, s, idx = meta["points"][lvl], meta["points"][lvl], meta["neighbors"][lvl]
        if cin == 1:
            x = torch.ones(s.shape[0], 1)
        else:
            # leaky-relu'd normal features: mixed signs, some rows with a non-positive sum (exercises neighbor_num)
            x = torch.nn.functional.leaky_relu(torch.randn(s.shape[0], cin) + 0.15, 0.1)
        with torch.no_grad():
            out = conv(q, s, idx, x)
        arrays.update({f"{tag}_q": t2n(q), f"{tag}_s": t2n(s), f"{tag}_idx": t2n(idx).astype(np.int32), f"{tag}_x": t2n(x),
                       f"{tag}_kp": t2n(conv.kernel_points),
                       f"{tag}_extent": np.asarray(extent, np.float32), f"{tag}_out": t2n(out)})
    arrays["tags"] = np.asarray([s[0] for s in specs])
    save("kpconv_layers.npz", **arrays)


def gen_blocks(ref):
    """UnaryBlock-style instance norm and max_pool from the reference modules."""
    torch.manual_seed(6)
    lens = torch.tensor([300, 2, 211, 64], dtype=torch.int32)
    n = int(lens.sum())
    x = torch.randn(n, 64) * 2 + 0.3
    bn = ref.kpconv_blocks.BatchNormBlock(64, True, 0.02)
    with torch.no_grad():
        y = bn(x, lens)
    idx = torch.randint(0, n + 1, (150, 12))
    mp = ref.kpconv_blocks.max_pool(x, idx)
    save("blocks.npz", x=t2n(x), lens=t2n(lens), inorm=t2n(y), pool_idx=t2n(idx).astype(np.int32), pool_out=t2n(mp))


def gen_pose(ref):
    rng = np.random.default_rng(31)
    arrays = {}
    cases = []
    for i, (n, scale, kind) in enumerate([(500, 1.0, "generic"), (60, 60.0, "kitti_scale"), (300, 1.0, "planar"),
                                          (3, 1.0, "minimal"), (200, 1.0, "reflection"), (100, 1.0, "unweighted")]):
        a = rng.normal(size=(n, 3)) * scale
        if kind == "planar":
            a[:, 2] = 0.3
        R = synthetic._random_rotation(rng, 170.0)
        t = rng.normal(size=3) * scale
        b = a @ R.T + t + rng.normal(size=(n, 3)) * 0.01 * scale
        if kind == "reflection":
            b = b * np.array([1, 1, -1.0])  # best orthogonal map is a reflection -> exercises the det<=0 branch
        w = rng.uniform(0, 1, size=n)
        if kind == "kitti_scale":
            a += np.array([40.0, -25.0, 1.0])
        a32, b32, w32 = a.astype(np.float32), b.astype(np.float32), w.astype(np.float32)
        with torch.no_grad():
            T = ref.se3_torch.compute_rigid_transform(torch.from_numpy(a32), torch.from_numpy(b32),
                                                      None if kind == "unweighted" else torch.from_numpy(w32))
        arrays[f"a_{i}"], arrays[f"b_{i}"], arrays[f"w_{i}"], arrays[f"T_{i}"] = a32, b32, w32, t2n(T)
        cases.append(kind)
    arrays["kinds"] = np.asarray(cases)
    # batched call ([B, N, 3]) as used at se3_torch.py:231
    a = rng.normal(size=(4, 120, 3)).astype(np.float32)
    b = rng.normal(size=(4, 120, 3)).astype(np.float32)
    w = rng.uniform(0, 1, size=(4, 120)).astype(np.float32)
    with torch.no_grad():
        T = ref.se3_torch.compute_rigid_transform(torch.from_numpy(a), torch.from_numpy(b), torch.from_numpy(w))
    arrays.update(batched_a=a, batched_b=b, batched_w=w, batched_T=t2n(T))
    save("procrustes.npz", **arrays)


def gen_matching(ref):
    """softmax_correlation of the reference model on synthetic conditioned features, both N > M and N <= M,
    with the 3DMatch (Sinkhorn) and the KITTI/ModelNet (argmax + Procrustes) settings."""
    rng = np.random.default_rng(41)
    arrays = {}
    for tag, yaml_name in (("sinkhorn", "qk_regtr_full_3dmatch.yaml"), ("argmax", "qk_regtr_full_kitti.yaml")):
        cfg = ref_torch.load_cfg(yaml_name)
        model = ref_torch.build_model(cfg, seed=0)
        shapes = [(95, 70), (60, 101), (33, 33)]
        src_f, tgt_f, src_xyz, tgt_xyz = [], [], [], []
        for n, m in shapes:
            # structured features: tgt rows are noisy copies of a subset of src rows so that matches are meaningful
            base = rng.normal(size=(max(n, m), 256)).astype(np.float32) * 2.0
            sf = base[:n] + 0.1 * rng.normal(size=(n, 256)).astype(np.float32)
            tf = base[rng.permutation(max(n, m))[:m]] + 0.1 * rng.normal(size=(m, 256)).astype(np.float32)
            src_f.append(torch.from_numpy(sf)[None])
            tgt_f.append(torch.from_numpy(tf)[None])
            src_xyz.append(torch.from_numpy(rng.normal(size=(n, 3)).astype(np.float32)))
            tgt_xyz.append(torch.from_numpy(rng.normal(size=(m, 3)).astype(np.float32)))
        with torch.no_grad():
            pose, attn, val, ind, sp, tp = model.softmax_correlation(src_f, tgt_f, src_xyz, tgt_xyz, None, None)
        arrays[f"{tag}_alpha"] = t2n(model.alpha)
        arrays[f"{tag}_beta"] = t2n(model.beta)
        arrays[f"{tag}_pose"] = t2n(pose)
        for i in range(len(shapes)):
            arrays[f"{tag}_src_f_{i}"] = t2n(src_f[i][0])
            arrays[f"{tag}_tgt_f_{i}"] = t2n(tgt_f[i][0])
            arrays[f"{tag}_src_xyz_{i}"] = t2n(src_xyz[i])
            arrays[f"{tag}_tgt_xyz_{i}"] = t2n(tgt_xyz[i])
            arrays[f"{tag}_attn_{i}"] = t2n(attn[i][0]).astype(np.float32)
            arrays[f"{tag}_val_{i}"] = t2n(val[i])
            arrays[f"{tag}_ind_{i}"] = t2n(ind[i])
        arrays[f"{tag}_n_pairs"] = np.asarray(len(shapes))
    save("matching.npz", **arrays)


REFINEMENT_VARIANTS = {
    "ratio": dict(use_ratio_test=True, lowe_thres=2e-4),   # dual-softmax ratios are ~1e-4 on these features
    "median": dict(threshold_corr=True),
    "overlap": dict(remove_outliers_overlap=True),
    "overlap_w": dict(remove_outliers_overlap=True, use_overlap_as_weights=True),
    "lgr": dict(use_lgr=True, acceptance_radius=0.3, num_refinement_steps=4),
    "topk": dict(remove_points_from_val=True, val_threshold=0.25),
    "combo": dict(use_ratio_test=True, lowe_thres=5e-4, threshold_corr=True, remove_outliers_overlap=True, use_lgr=True,
                  acceptance_radius=0.3, num_refinement_steps=3),
}


def gen_refinements(ref):
    """softmax_correlation of the reference model with each optional refinement switched on (KITTI settings: argmax
    correspondences + Procrustes).  RegTR.ransac is absent: it moves its draws to .cuda() (:404)."""
    rng = np.random.default_rng(43)
    shapes = [(95, 70), (60, 101), (48, 48)]
    ang = 0.4
    R = np.array([[np.cos(ang), -np.sin(ang), 0], [np.sin(ang), np.cos(ang), 0], [0, 0, 1]], dtype=np.float32)
    t = np.array([0.5, -0.3, 0.2], dtype=np.float32)
    src_f, tgt_f, src_xyz, tgt_xyz, src_ov, tgt_ov = [], [], [], [], [], []
    for n, m in shapes:
        k = max(n, m)
        base = rng.normal(size=(k, 256)).astype(np.float32) * 0.6   # moderately peaked attention
        xyz = rng.uniform(-2, 2, size=(k, 3)).astype(np.float32)
        perm = rng.permutation(k)[:m]
        sf = base[:n] + 0.1 * rng.normal(size=(n, 256)).astype(np.float32)
        tf = base[perm] + 0.1 * rng.normal(size=(m, 256)).astype(np.float32)
        txyz = xyz[perm] @ R.T + t + 0.02 * rng.normal(size=(m, 3)).astype(np.float32)
        bad = rng.uniform(size=m) < 0.25                       # gross outliers for the inlier test to reject
        txyz[bad] += rng.uniform(-1.5, 1.5, size=(int(bad.sum()), 3)).astype(np.float32)
        src_f.append(torch.from_numpy(sf)[None])
        tgt_f.append(torch.from_numpy(tf)[None])
        src_xyz.append(torch.from_numpy(xyz[:n].copy()))
        tgt_xyz.append(torch.from_numpy(txyz.astype(np.float32)))
        src_ov.append(torch.from_numpy(rng.uniform(0.05, 1, size=(1, n, 1)).astype(np.float32)))
        tgt_ov.append(torch.from_numpy(rng.uniform(0.05, 1, size=(1, m, 1)).astype(np.float32)))
    arrays = {"n_pairs": np.asarray(len(shapes)), "variants": np.asarray(sorted(REFINEMENT_VARIANTS))}
    for i in range(len(shapes)):
        arrays[f"src_f_{i}"], arrays[f"tgt_f_{i}"] = t2n(src_f[i][0]), t2n(tgt_f[i][0])
        arrays[f"src_xyz_{i}"], arrays[f"tgt_xyz_{i}"] = t2n(src_xyz[i]), t2n(tgt_xyz[i])
        arrays[f"src_ov_{i}"], arrays[f"tgt_ov_{i}"] = t2n(src_ov[i]), t2n(tgt_ov[i])
    for tag, overrides in REFINEMENT_VARIANTS.items():
        cfg = ref_torch.load_cfg("qk_regtr_full_kitti.yaml", **overrides)
        model = ref_torch.build_model(cfg, seed=0)
        with torch.no_grad():
            pose, attn, val, ind, sp, tp = model.softmax_correlation(src_f, tgt_f, src_xyz, tgt_xyz, src_ov, tgt_ov)
        arrays[f"{tag}_pose"] = t2n(pose)
        for i in range(len(shapes)):
            arrays[f"{tag}_val_{i}"] = t2n(val[i])
            arrays[f"{tag}_ind_{i}"] = t2n(ind[i])
            arrays[f"{tag}_src_pts_{i}"] = t2n(sp[i])
            arrays[f"{tag}_tgt_pts_{i}"] = t2n(tp[i])
        print(tag, [int((t2n(v) > 0).sum()) for v in val], "non-zero weights of", [t2n(v).size for v in val])
    save("refinements.npz", **arrays)


def gen_forward(ref):
    """Encoder features and the full forward of the reference model (weights from tests/golden/weights.py)."""
    for tag, yaml_name, data in (
            ("3dmatch", "qk_regtr_full_3dmatch.yaml", synthetic.make_batch("3dmatch", 2, seed=51, n_points=800)),
            ("modelnet", "qk_regtr_full_modelnet.yaml", synthetic.make_batch("modelnet", 1, seed=52))):
        cfg = ref_torch.load_cfg(yaml_name)
        model = ref_torch.build_model(cfg, seed=0)
        kps = load_weights(model, seed=1234)
        batch = {"src_xyz": [torch.from_numpy(c) for c in data["src_xyz"]],
                 "tgt_xyz": [torch.from_numpy(c) for c in data["tgt_xyz"]]}
        with torch.no_grad():
            out = model(batch)
            meta = batch["kpconv_meta"]
            feats0 = torch.ones_like(meta["points"][0][:, 0:1])
            enc, _ = model.kpf_encoder(feats0, meta)
        arrays = pyramid_arrays(meta, prefix="meta_")
        arrays.update({f"kp::{k}": v for k, v in kps.items()})
        B = len(batch["src_xyz"])
        for i in range(B):
            arrays[f"src_{i}"] = data["src_xyz"][i]
            arrays[f"tgt_{i}"] = data["tgt_xyz"][i]
            arrays[f"src_feat_{i}"] = t2n(out["src_feat"][i][0])
            arrays[f"tgt_feat_{i}"] = t2n(out["tgt_feat"][i][0])
            arrays[f"val_{i}"] = t2n(out["overlap_prob_list"][i])
            arrays[f"ind_{i}"] = t2n(out["ind_list"][i])
        if tag == "3dmatch":
            arrays["encoder_out"] = t2n(enc)
        arrays["pose"] = t2n(out["pose"])
        arrays["n_pairs"] = np.asarray(B)
        arrays["weight_seed"] = np.asarray(1234)
        save(f"forward_{tag}.npz", **arrays)


def gen_sinkhorn(ref):
    """utils/se3_torch.py `sinkhorn` and `compute_rigid_transform_with_sinkhorn` of the reference on random affinities."""
    rng = np.random.default_rng(47)
    B, J, K = 3, 37, 53
    aff = (rng.normal(size=(B, J, K)) * 3.0).astype(np.float32)
    xs = rng.normal(size=(B, J, 3)).astype(np.float32)
    xt = rng.normal(size=(B, K, 3)).astype(np.float32)
    arrays = {"affinity": aff, "xyz_s": xs, "xyz_t": xt}
    with torch.no_grad():
        for it in (1, 3, 5):
            arrays[f"log_perm_{it}"] = t2n(ref.se3_torch.sinkhorn(torch.from_numpy(aff), n_iters=it, slack=True))
        arrays["transform_3"] = t2n(ref.se3_torch.compute_rigid_transform_with_sinkhorn(
            torch.from_numpy(xs), torch.from_numpy(xt), torch.from_numpy(aff), True, 3))
        arrays["transform_single"] = t2n(ref.se3_torch.compute_rigid_transform_with_sinkhorn(
            torch.from_numpy(xs[:1]), torch.from_numpy(xt[:1]), torch.from_numpy(aff[:1]), True, 3))
    save("sinkhorn.npz", **arrays)


def gen_est_log(ref):
    """The reference's 3DMatch est.log writer (models/generic_reg_model.py:382-403) on a small batch, both pose ranks."""
    import importlib
    import tempfile
    import types
    grm = importlib.import_module("models.generic_reg_model")
    rng = np.random.default_rng(5)
    B = 3
    batch = {"src_xyz": [torch.zeros(4, 3)] * B, "tgt_xyz": [torch.zeros(5, 3)] * B,
             "src_path": ["test/7-scenes-redkitchen/cloud_bin_12.pth", "test/7-scenes-redkitchen/cloud_bin_3.pth",
                          "test/sun3d-hotel_umd-maryland_hotel3/cloud_bin_40.pth"],
             "tgt_path": ["test/7-scenes-redkitchen/cloud_bin_0.pth", "test/7-scenes-redkitchen/cloud_bin_1.pth",
                          "test/sun3d-hotel_umd-maryland_hotel3/cloud_bin_7.pth"]}
    pose = torch.from_numpy(rng.normal(size=(B, 3, 4)).astype(np.float32))
    d = tempfile.mkdtemp()
    fake = types.SimpleNamespace(_log_path=d, cfg=types.SimpleNamespace(benchmark="3DMatch"))
    grm.GenericRegModel._save_3DMatch_log(fake, batch, {"pose": pose})
    grm.GenericRegModel._save_3DMatch_log(fake, batch, {"pose": pose[None].repeat(2, 1, 1, 1)})   # (iters, B, 3, 4)
    out = {}
    for root, _, files in os.walk(d):
        for f in files:
            out[os.path.relpath(os.path.join(root, f), d)] = open(os.path.join(root, f)).read()
    save("est_log.npz", pose=pose.numpy(), src_path=np.asarray(batch["src_path"]), tgt_path=np.asarray(batch["tgt_path"]),
         files=np.asarray(sorted(out)), texts=np.asarray([out[k] for k in sorted(out)]))


WELLCOND_ARCH = ["simple", "resnetb", "resnetb_strided", "resnetb", "resnetb", "resnetb_strided", "resnetb", "resnetb",
                 "resnetb_strided", "resnetb", "resnetb"]          # conf/qk_regtr_full_3dmatch.yaml:64-74 (4-stage)
WELLCOND_DAMP = 0.3


def gen_forward_wellcond(ref):
    """A WELL-CONDITIONED end-to-end case on the 4-stage architecture BASELINE.json names: tgt is a translated copy of
    src (translation = a multiple of the coarsest voxel, so both pyramids see the same geometry), and the residual
    branches of the cross-encoder are damped (weights.damp_transformer) so that conditioned features stay close to the
    translation-invariant KPConv descriptors.  Corresponding superpoints are then each other's best match and the
    reference recovers the ground-truth pose; its fp32 result is insensitive to summation order, so the end-to-end
    pose can be held to north_star's 1e-3 deg / 1e-5 m.  Two variants on the same clouds and weights:
      argmax    use_sinkhorn=False (the KITTI / ModelNet matching route), target jittered by 1 mm
      sinkhorn  the 3DMatch yaml's route, exact copy (the reference's affinity rewards LOW correlation, :535, so its
                untrained Sinkhorn assignment is close to uniform, the pose far from the ground truth and only
                moderately conditioned; recorded as is, together with the reference's own self-noise)
    Each variant also stores how far the REFERENCE's pose moves when the points of every cloud are relabelled
    (`*_self_noise_*`): the floor below which a difference from the reference means nothing.
    """
    rng = np.random.default_rng(81)
    trans = np.array([0.3, -0.2, 0.1])
    data = synthetic.make_batch("3dmatch", 2, seed=81, n_points=4000)
    srcs = data["src_xyz"]
    arrays = {"weight_seed": np.asarray(1234), "damp": np.asarray(WELLCOND_DAMP, np.float32), "n_pairs": np.asarray(2),
              "gt_pose": np.concatenate([np.eye(3), trans[:, None]], 1).astype(np.float32)}
    for tag, jitter, overrides in (("argmax", 0.001, dict(use_sinkhorn=False)), ("sinkhorn", 0.0, {})):
        tgts = [(s.astype(np.float64) + trans + rng.normal(0, jitter, s.shape)).astype(np.float32) for s in srcs]
        cfg = ref_torch.load_cfg("qk_regtr_full_3dmatch.yaml", architecture=list(WELLCOND_ARCH), **overrides)
        model = ref_torch.build_model(cfg, seed=0)
        sd = model.state_dict()
        vals = damp_transformer(filled_state({k: tuple(v.shape) for k, v in sd.items()}, 1234), WELLCOND_DAMP)
        model.load_state_dict({k: (torch.from_numpy(vals[k]) if k in vals else v) for k, v in sd.items()})
        batch = {"src_xyz": [torch.from_numpy(c) for c in srcs], "tgt_xyz": [torch.from_numpy(c) for c in tgts]}
        with torch.no_grad():
            out = model(batch)
        for k, v in sd.items():
            if k.endswith("kernel_points"):
                arrays[f"kp::{k}"] = t2n(v)
        for i in range(2):
            arrays[f"src_{i}"] = srcs[i]
            arrays[f"{tag}_tgt_{i}"] = tgts[i]
            arrays[f"{tag}_ind_{i}"] = t2n(out["ind_list"][i])
            arrays[f"{tag}_val_{i}"] = t2n(out["overlap_prob_list"][i])
            arrays[f"{tag}_n_src_{i}"] = np.asarray(out["src_feat"][i].shape[1])
            arrays[f"{tag}_n_tgt_{i}"] = np.asarray(out["tgt_feat"][i].shape[1])
        arrays[f"{tag}_pose"] = t2n(out["pose"])
        # the reference's own conditioning on this input: its pose when the points of every cloud are merely
        # relabelled (three shuffles) -- mathematically the same problem, different fp32 summation orders
        from parity import pose_error
        noise_rot, noise_tr = 0.0, 0.0
        for trial in range(3):
            prng = np.random.default_rng(100 + trial)
            shuffled = {"src_xyz": [torch.from_numpy(c[prng.permutation(len(c))]) for c in srcs],
                        "tgt_xyz": [torch.from_numpy(c[prng.permutation(len(c))]) for c in tgts]}
            with torch.no_grad():
                again = model(shuffled)
            rot, tr = pose_error(t2n(again["pose"]), t2n(out["pose"]))
            noise_rot, noise_tr = max(noise_rot, float(rot.max())), max(noise_tr, float(tr.max()))
        arrays[f"{tag}_self_noise_rot_deg"] = np.asarray(noise_rot)
        arrays[f"{tag}_self_noise_trans"] = np.asarray(noise_tr)
        print(tag, f"reference self-noise under relabelling: {noise_rot:.2e} deg / {noise_tr:.2e} m")
        print(tag, "superpoints", [(int(arrays[f"{tag}_n_src_{i}"]), int(arrays[f"{tag}_n_tgt_{i}"])) for i in range(2)])
    save("forward_wellcond.npz", **arrays)


def main():
    oracle.build()
    ref = ref_torch.load()
    gen_preprocess(ref)
    gen_kpconv(ref)
    gen_blocks(ref)
    gen_pose(ref)
    gen_matching(ref)
    gen_refinements(ref)
    gen_sinkhorn(ref)
    gen_est_log(ref)
    gen_forward(ref)
    gen_forward_wellcond(ref)


if __name__ == "__main__":
    main()
