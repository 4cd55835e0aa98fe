"""One forward of the bench workload between cudaProfilerStart/Stop (for `ncu --profile-from-start off`):
    python tools/profile_step.py [--pairs 32] [--arch 4stage] [--kind 3dmatch]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import superpoints_registration_b200 as spr
from superpoints_registration_b200 import synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=32)
ap.add_argument("--points", type=int, default=20000)
ap.add_argument("--arch", default="4stage", choices=["3stage", "4stage"])
ap.add_argument("--kind", default="3dmatch", choices=["3dmatch", "kitti", "modelnet"])
args = ap.parse_args()
dev = "cuda:0"
if args.kind == "3dmatch":
    cfg = spr.threedmatch_config() if args.arch == "3stage" else spr.threedmatch_4stage_config()
    kw = dict(n_points=args.points)
elif args.kind == "kitti":
    cfg, kw = spr.kitti_config(first_subsampling_dl=0.3), dict(n_points=30000, voxel=0.3)
else:
    cfg, kw = spr.modelnet_config(), {}
torch.manual_seed(0)
np.random.seed(0)
model = spr.RegTR(cfg).to(dev).eval()
model.return_attn = False
data = synthetic.make_batch(args.kind, args.pairs, seed=200, **kw)
batch = {"src_xyz": [torch.from_numpy(c).to(dev) for c in data["src_xyz"]],
         "tgt_xyz": [torch.from_numpy(c).to(dev) for c in data["tgt_xyz"]]}
for _ in range(3):
    model(dict(batch))
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = model(dict(batch))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("pose finite:", bool(torch.isfinite(out["pose"]).all()))
