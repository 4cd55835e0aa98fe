"""Callers and data formats either side of the hot path: what lets real 3DMatch / KITTI
data flow into `RegTR.forward` and its poses flow out to the registration-recall benchmark.

  collate_pair            data_loaders/collate_functions.py:4-23      list of per-pair dicts -> the batch dict forward() takes
  load_threedmatch_pair   data_loaders/threedmatch.py:74-98           one 3DMatch / 3DLoMatch pair from its .pth fragments
  voxel_down_sample       data_loaders/kitti_pred.py:12-14,203-204    first point of every voxel (kiss_icp's down-sampler), on the GPU
  save_3dmatch_log        models/generic_reg_model.py:382-403         append poses to <log>/<benchmark>/<scene>/est.log

The training-only parts of the loaders (overlap masks, correspondences, augmentation) are out of scope: inference
needs the clouds, the paths and -- for evaluation -- the ground-truth pose.
"""
from __future__ import annotations

import os
from typing import Dict, List, Mapping, Sequence

import numpy as np
import torch

from . import ops

_RETAIN_AS_LIST = ('src_xyz', 'tgt_xyz', 'tgt_raw', 'src_overlap', 'tgt_overlap', 'correspondences', 'src_path',
                   'tgt_path', 'idx')


def collate_pair(list_data: Sequence[Mapping]) -> Dict:
    """Same contract as the reference's collate_pair: variable-size fields stay Python lists, `pose` is stacked to
    (B, 3, 4), `overlap_p` (when present) becomes a tensor."""
    batch_sz = len(list_data)
    data = {k: [list_data[b][k] for b in range(batch_sz)] for k in _RETAIN_AS_LIST if k in list_data[0]}
    data['pose'] = torch.stack([list_data[b]['pose'] for b in range(batch_sz)], dim=0)
    if 'overlap_p' in list_data[0]:
        data['overlap_p'] = torch.tensor([list_data[b]['overlap_p'] for b in range(batch_sz)])
    return data


def se3_init(rot: np.ndarray, trans: np.ndarray) -> np.ndarray:
    """utils/se3_numpy.py: [R | t] as (3, 4)."""
    return np.concatenate([np.asarray(rot), np.asarray(trans).reshape(3, 1)], axis=-1)


def load_threedmatch_pair(base_dir: str, infos: Mapping, item: int) -> Dict:
    """One pair of the 3DMatch / 3DLoMatch info pickles (datasets/3dmatch/*.pkl: keys 'rot', 'trans', 'src', 'tgt',
    'overlap'); fragments are the .pth arrays Predator's preprocessing wrote.  pose transforms src to tgt."""
    pose = se3_init(infos['rot'][item], infos['trans'][item])
    src_path, tgt_path = infos['src'][item], infos['tgt'][item]
    src_xyz = torch.load(os.path.join(base_dir, src_path), weights_only=False)
    tgt_xyz = torch.load(os.path.join(base_dir, tgt_path), weights_only=False)
    return {
        'src_xyz': torch.as_tensor(np.asarray(src_xyz)).float(), 'tgt_xyz': torch.as_tensor(np.asarray(tgt_xyz)).float(),
        'pose': torch.from_numpy(pose).float(), 'idx': item, 'src_path': src_path, 'tgt_path': tgt_path,
        'overlap_p': infos['overlap'][item],
    }


def voxel_down_sample(points, voxel_size: float) -> torch.Tensor:
    """The KITTI loader's down-sampler: the first point (in input order) of every voxel of edge `voxel_size`, voxel =
    floor(p / voxel_size) (kiss_icp VoxelDownsample).  Runs on the GPU (spr_grid_subsample_batch, first-point mode);
    points come out in input order, where kiss_icp emits its hash map's order."""
    pts = torch.as_tensor(points)
    if not pts.is_cuda:
        raise RuntimeError("voxel_down_sample: points must be a CUDA tensor (no CPU fallback on the B200 path)")
    pts = pts.to(torch.float32).contiguous()
    lengths = torch.tensor([pts.shape[0]], dtype=torch.int32).pin_memory().to(pts.device, non_blocking=True)
    out, _ = ops.grid_subsample_batch(pts, lengths, float(voxel_size), mode="first_point")
    return out


def save_3dmatch_log(log_path: str, benchmark: str, batch: Mapping, pred: Mapping) -> List[str]:
    """Append the predicted poses of a batch to <log_path>/<benchmark>/<scene>/est.log in the format
    benchmark_predator reads: a header line `tgt_idx \\t src_idx \\t -1`, then the 4x4 pose, 12 decimals, tab separated.
    Returns the files written."""
    poses = pred['pose']
    written = []
    for b in range(len(batch['src_xyz'])):
        scene = batch['src_path'][b].split(os.path.sep)[1]
        src_idx = int(os.path.basename(batch['src_path'][b]).split('_')[-1].replace('.pth', ''))
        tgt_idx = int(os.path.basename(batch['tgt_path'][b]).split('_')[-1].replace('.pth', ''))
        pose = poses[-1][b] if poses.ndim == 4 else poses[b]
        pose_np = pose.detach().cpu().numpy() if isinstance(pose, torch.Tensor) else np.asarray(pose)
        if pose_np.shape[0] == 3:
            pose_np = np.concatenate([pose_np, [[0., 0., 0., 1.]]], axis=0)
        scene_folder = os.path.join(log_path, benchmark, scene)
        os.makedirs(scene_folder, exist_ok=True)
        est_log_path = os.path.join(scene_folder, 'est.log')
        with open(est_log_path, 'a') as fid:
            fid.write('{}\t{}\t{}\n'.format(tgt_idx, src_idx, -1))  # the frame count is unknown; the benchmark ignores it
            for i in range(4):
                fid.write('\t'.join(map('{0:.12f}'.format, pose_np[i])) + '\n')
        written.append(est_log_path)
    return written
