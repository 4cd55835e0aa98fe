"""Host-side logic that needs no GPU: configuration, level planning, packing, state_dict surface, errors."""
import os

import numpy as np
import pytest
import torch
import yaml

import superpoints_registration_b200 as spr
from oracle import pipeline
from superpoints_registration_b200 import ops
from superpoints_registration_b200.config import flatten_sections
from superpoints_registration_b200.kpconv import _split_levels
from superpoints_registration_b200.kernel_points import load_kernels


def test_config_flattening_matches_reference_semantics(tmp_path):
    doc = {"general": {"expt_name": "x"}, "kpconv_options": {"conv_radius": 2.5, "architecture": ["simple"]},
           "model": {"d_embed": 256, "conv_radius": 3.0}}
    p = tmp_path / "c.yaml"
    p.write_text(yaml.safe_dump(doc))
    cfg = spr.load_config(str(p))
    assert cfg.expt_name == "x" and cfg["d_embed"] == 256
    assert cfg.conv_radius == 3.0            # later sections win, section names are discarded (utils/misc.py:24-27)
    assert cfg.get("missing", 7) == 7
    with pytest.raises(AttributeError):
        cfg.nope


@pytest.mark.parametrize("cfg,levels", [(spr.threedmatch_config(), 3), (spr.threedmatch_4stage_config(), 4),
                                        (spr.kitti_config(), 4), (spr.modelnet_config(), 2)])
def test_level_plan(cfg, levels):
    plan = _split_levels(cfg.architecture)
    assert len(plan) == levels
    assert all(has_conv for has_conv, _ in plan)
    assert [s for _, s in plan] == [True] * (levels - 1) + [False]
    assert plan == pipeline.level_plan(cfg.architecture)      # product and oracle agree on the walk


def test_level_plan_rejects_deformable():
    with pytest.raises(NotImplementedError):
        _split_levels(["simple", "resnetb_deformable"])


def test_state_dict_keys_are_the_reference_ones():
    m = spr.RegTR(spr.threedmatch_config())
    keys = set(m.state_dict())
    for k in ("alpha", "beta", "kpf_encoder.encoder_blocks.0.KPConv.weights",
              "kpf_encoder.encoder_blocks.0.KPConv.kernel_points", "kpf_encoder.encoder_blocks.1.unary1.mlp.weight",
              "kpf_encoder.encoder_blocks.1.unary2.mlp.weight", "kpf_encoder.encoder_blocks.1.unary_shortcut.mlp.weight",
              "feat_proj.weight", "feat_proj.bias", "transformer_encoder.layers.5.self_attn.in_proj_weight",
              "transformer_encoder.layers.0.multihead_attn.out_proj.bias", "transformer_encoder.layers.0.linear1.weight",
              "transformer_encoder.layers.0.norm3.weight", "transformer_encoder.norm.weight",
              "overlap_predictor.weight"):
        assert k in keys, k
    assert sum(p.numel() for p in m.parameters()) == 7797547 - 2 * 256 * 256   # reference minus its two loss matrices
    enc = m.kpf_encoder
    assert enc.encoder_skip_dims[-1] == 512 and len(enc.encoder_blocks) == 8
    k = spr.RegTR(spr.kitti_config()).kpf_encoder
    assert k.encoder_skip_dims[-1] == 1024 and len(k.encoder_blocks) == 11
    assert spr.RegTR(spr.modelnet_config()).kpf_encoder.encoder_skip_dims[-1] == 1024


def test_unsupported_options_fail_loudly():
    with pytest.raises(NotImplementedError):
        spr.RegTR(spr.threedmatch_config(use_ransac=True))
    with pytest.raises(NotImplementedError):
        spr.KPConv(15, 3, 32, 32, 0.05, 0.06, aggregation_mode="closest")
    with pytest.raises(ValueError):
        spr.block_decider("nonsense", 0.1, 32, 64, 0, spr.threedmatch_config())


def test_no_cpu_fallback():
    pts = torch.rand(100, 3)
    lens = torch.tensor([100], dtype=torch.int32)
    with pytest.raises(RuntimeError, match="CUDA"):
        spr.batch_neighbors_kpconv(pts, pts, lens, lens, 0.1, 10)
    with pytest.raises(RuntimeError, match="CUDA"):
        spr.batch_grid_subsampling_kpconv(pts, lens, sampleDl=0.1)
    with pytest.raises(RuntimeError, match="CUDA"):
        spr.Preprocessor(spr.threedmatch_config())([pts])
    with pytest.raises(RuntimeError, match="CUDA"):
        spr.compute_rigid_transform(pts, pts)


def test_packed_pairs_offsets():
    pp = ops.PackedPairs([5, 2, 4], [3, 7, 4], "cpu")
    assert pp.h_so == [0, 5, 7, 11] and pp.h_to == [0, 3, 10, 14]
    assert pp.h_co == [0, 15, 29, 45]
    assert pp.h_oo == [0, 3, 5, 9]          # N>M -> M outputs, else N outputs
    assert pp.max_n == 5 and pp.max_m == 7


def test_kernel_points_shape_and_center():
    np.random.seed(0)
    kp = load_kernels(0.0625, 15)
    assert kp.shape == (15, 3) and kp.dtype == np.float32
    r = np.linalg.norm(kp, axis=1)
    assert r[0] < 0.05 * 0.0625 * 3 and abs(r[1:].mean() / 0.0625 - 0.66) < 0.05


def test_pose_error_metric_resolves_below_reference_metric_floor():
    T = torch.tensor([[[0.6, -0.8, 0.0, 1.0], [0.8, 0.6, 0.0, 2.0], [0.0, 0.0, 1.0, 3.0]]])
    rot, tr = spr.pose_error(T, T)
    assert float(rot) == 0.0 and float(tr) == 0.0


def test_est_log_writer_matches_the_reference_byte_for_byte(tmp_path, golden_dir):
    """frontends.save_3dmatch_log against the text the reference's _save_3DMatch_log wrote for the same batch
    (models/generic_reg_model.py:382-403; tests/golden/make_golden.py:gen_est_log): two appends, both pose ranks."""
    import os

    import numpy as np
    import torch

    from superpoints_registration_b200 import frontends
    g = np.load(os.path.join(golden_dir, "est_log.npz"))
    pose = torch.from_numpy(g["pose"])
    B = pose.shape[0]
    batch = {"src_xyz": [torch.zeros(1, 3)] * B, "tgt_xyz": [torch.zeros(1, 3)] * B,
             "src_path": [str(s) for s in g["src_path"]], "tgt_path": [str(s) for s in g["tgt_path"]]}
    frontends.save_3dmatch_log(str(tmp_path), "3DMatch", batch, {"pose": pose})
    frontends.save_3dmatch_log(str(tmp_path), "3DMatch", batch, {"pose": pose[None].repeat(2, 1, 1, 1)})
    for rel, text in zip(map(str, g["files"]), map(str, g["texts"])):
        assert open(os.path.join(str(tmp_path), rel)).read() == text, rel


def test_collate_pair_contract():
    """data_loaders/collate_functions.py:4-23: variable-size fields stay lists, pose is stacked, overlap_p is a tensor."""
    import torch

    from superpoints_registration_b200 import frontends
    items = [{"src_xyz": torch.rand(5 + i, 3), "tgt_xyz": torch.rand(7, 3), "pose": torch.eye(4)[:3], "idx": i,
              "src_path": f"a/b/c_{i}.pth", "tgt_path": f"a/b/c_{i + 1}.pth", "overlap_p": 0.1 * i, "extra": 1}
             for i in range(3)]
    batch = frontends.collate_pair(items)
    assert isinstance(batch["src_xyz"], list) and [t.shape[0] for t in batch["src_xyz"]] == [5, 6, 7]
    assert tuple(batch["pose"].shape) == (3, 3, 4) and batch["idx"] == [0, 1, 2]
    assert torch.allclose(batch["overlap_p"], torch.tensor([0.0, 0.1, 0.2])) and "extra" not in batch
