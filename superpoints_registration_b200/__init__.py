"""superpoints_registration_b200 -- the data-parallel hot path of neu-vi/Superpoints_Registration
(KPConv preprocessing + KPConv backbone + superpoint matching + pose solve) as hand-written sm_100a CUDA
kernels behind a C ABI (include/spr_b200.h), with a Python host layer that mirrors the reference's
operator surface.  Importing the package does not need a GPU; calling any operator does, and fails loudly
when libspr_b200.so has not been built.
"""
from .config import Config, load_config, threedmatch_config, threedmatch_4stage_config, kitti_config, modelnet_config
from .kpconv import Preprocessor, PreprocessorGPU, KPFEncoder, batch_grid_subsampling_kpconv, batch_neighbors_kpconv
from .kpconv_blocks import KPConv, UnaryBlock, SimpleBlock, ResnetBottleneckBlock, BatchNormBlock, max_pool, block_decider
from .se3 import compute_rigid_transform, compute_rigid_transform_with_sinkhorn, sinkhorn, se3_transform, se3_inv, se3_cat, pose_error
from .model import RegTR
from . import frontends

__all__ = [
    "Config", "load_config", "threedmatch_config", "threedmatch_4stage_config", "kitti_config", "modelnet_config",
    "Preprocessor", "PreprocessorGPU", "KPFEncoder", "batch_grid_subsampling_kpconv", "batch_neighbors_kpconv",
    "KPConv", "UnaryBlock", "SimpleBlock", "ResnetBottleneckBlock", "BatchNormBlock", "max_pool", "block_decider",
    "compute_rigid_transform", "compute_rigid_transform_with_sinkhorn", "sinkhorn", "se3_transform", "se3_inv", "se3_cat", "pose_error", "RegTR", "frontends",
]
