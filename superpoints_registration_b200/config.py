"""Configuration: the reference's yaml files, flattened the way the reference flattens them.

Reference: utils/misc.py:10-29 (`load_config` merges every top-level section into one namespace, discarding
the section names) and the EasyDict wrapper applied in train.py / test.py.  easydict is not a dependency
here; `Config` gives the same attribute + item access.
"""
from __future__ import annotations

from typing import Any, Dict

import yaml


class Config(dict):
    """dict with attribute access (`cfg.conv_radius`, `cfg['conv_radius']`, `cfg.get(...)`)."""

    def __getattr__(self, key: str) -> Any:
        try:
            return self[key]
        except KeyError as exc:
            raise AttributeError(key) from exc

    def __setattr__(self, key: str, value: Any) -> None:
        self[key] = value


def flatten_sections(sections: Dict[str, Dict[str, Any]]) -> Config:
    flat = Config()
    for section in sections.values():
        if isinstance(section, dict):
            flat.update(section)
    return flat


def load_config(path: str) -> Config:
    with open(path, "r") as f:
        return flatten_sections(yaml.safe_load(f))


def _base(**kw) -> Config:
    cfg = Config(
        # kpconv_options
        aggregation_mode="sum", fixed_kernel_points="center", in_feats_dim=1, in_points_dim=3, deform_radius=5.0,
        KP_extent=2.0, KP_influence="linear", use_batch_norm=True, batch_norm_momentum=0.02, modulated=False,
        num_kernel_points=15,
        # model
        remove_points_from_val=False, threshold_corr=False, remove_outliers_overlap=False, use_overlap_as_weights=False,
        use_ratio_test=False, use_sinkhorn=False, sinkhorn_itr=1, slack=False, use_attn_affinity=False,
        use_corr_affinity=False, use_lgr=False, use_ransac=False, lowe_thres=0.9,
        attention_type="dot_prod", nhead=8, d_embed=256, d_feedforward=1024, dropout=0.0, pre_norm=True,
        transformer_act="relu", num_encoder_layers=6, transformer_encoder_has_pos_emb=True, sa_val_has_pos_emb=True,
        ca_val_has_pos_emb=True, pos_emb_type="sine",
    )
    cfg.update(kw)
    return cfg


def threedmatch_config(**overrides) -> Config:
    """The hot-path keys of conf/qk_regtr_full_3dmatch.yaml (3-stage shipped architecture, :56-63)."""
    cfg = _base(
        dataset="3dmatch", neighborhood_limits=[40, 40, 40, 40], first_subsampling_dl=0.025, first_feats_dim=128,
        conv_radius=2.5, num_layers=4,
        architecture=["simple", "resnetb", "resnetb_strided", "resnetb", "resnetb", "resnetb_strided", "resnetb",
                      "resnetb"],
        use_sinkhorn=True, sinkhorn_itr=3, slack=True,
        num_refinement_steps=4, acceptance_radius=0.1, val_threshold=0.15)
    cfg.update(overrides)
    return cfg


def threedmatch_4stage_config(**overrides) -> Config:
    """The commented-out 4-stage variant (conf/qk_regtr_full_3dmatch.yaml:64-74) BASELINE.json's config names."""
    cfg = threedmatch_config(
        architecture=["simple", "resnetb", "resnetb_strided", "resnetb", "resnetb", "resnetb_strided", "resnetb",
                      "resnetb", "resnetb_strided", "resnetb", "resnetb"])
    cfg.update(overrides)
    return cfg


def kitti_config(**overrides) -> Config:
    """conf/qk_regtr_full_kitti.yaml."""
    cfg = _base(
        dataset="kitti", neighborhood_limits=[39, 57, 68, 74], first_subsampling_dl=0.2, first_feats_dim=128,
        conv_radius=4.25, num_layers=4,
        architecture=["simple", "resnetb", "resnetb_strided", "resnetb", "resnetb", "resnetb_strided", "resnetb",
                      "resnetb", "resnetb_strided", "resnetb", "resnetb"],
        use_sinkhorn=False, num_refinement_steps=10, acceptance_radius=0.6, val_threshold=0.25)
    cfg.update(overrides)
    return cfg


def modelnet_config(**overrides) -> Config:
    """conf/qk_regtr_full_modelnet.yaml."""
    cfg = _base(
        dataset="modelnet", neighborhood_limits=[50, 50], first_subsampling_dl=0.03, first_feats_dim=512,
        conv_radius=2.75, num_layers=2,
        architecture=["simple", "resnetb", "resnetb", "resnetb_strided", "resnetb", "resnetb"],
        use_sinkhorn=False, sinkhorn_itr=1, slack=False, num_refinement_steps=5, acceptance_radius=0.05)
    cfg.update(overrides)
    return cfg
