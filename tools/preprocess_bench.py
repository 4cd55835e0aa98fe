"""Builds the 3DMatch-shape pyramid a few times (grid subsampling + radius searches); short enough to sit under ncu.
    python tools/preprocess_bench.py [--pairs 8] [--reps 3]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import superpoints_registration_b200 as spr
from superpoints_registration_b200 import synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=8)
ap.add_argument("--reps", type=int, default=3)
args = ap.parse_args()
dev = "cuda:0"
cfg = spr.threedmatch_config()
data = synthetic.make_batch("3dmatch", args.pairs, seed=2, n_points=20000)
clouds = [torch.from_numpy(c).to(dev) for c in data["src_xyz"] + data["tgt_xyz"]]
pre = spr.Preprocessor(cfg)
best = 1e9
for _ in range(args.reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); meta = pre(clouds); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f"pyramid: {best:.3f} ms; levels {[tuple(p.shape) for p in meta['points']]}")
