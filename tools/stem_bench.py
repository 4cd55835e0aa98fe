"""Stem KPConv (Cin = 1 -> 64) at the bench size (level 0 of a 3DMatch-shape pyramid): CUDA-event timing with L2 flushed
and a sampled comparison with the fp64 oracle.  SPR_STEM_GEN=1 in the environment selects the round-1 kernel.
    python tools/stem_bench.py [--pairs 32] [--reps 5]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import superpoints_registration_b200 as spr
from superpoints_registration_b200 import ops, synthetic
from superpoints_registration_b200.kernel_points import load_kernels

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=32)
ap.add_argument("--points", type=int, default=20000)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--check", type=int, default=256)
args = ap.parse_args()
dev = "cuda:0"
cfg = spr.threedmatch_4stage_config()
data = synthetic.make_batch("3dmatch", args.pairs, seed=2, n_points=args.points)
meta = spr.Preprocessor(cfg)([torch.from_numpy(c).to(dev) for c in data["src_xyz"] + data["tgt_xyz"]])
rng = np.random.default_rng(0)
r = cfg.first_subsampling_dl * cfg.conv_radius
ext = r * cfg.KP_extent / cfg.conv_radius
s = meta["points"][0]
idx = meta.index("neighbors", 0)
n, H = idx.shape
cout = cfg.first_feats_dim // 2
x = torch.ones(n, 1, device=dev)
w = torch.from_numpy((rng.normal(size=(15, 1, cout)) / 4).astype(np.float32)).to(dev)
kp = torch.from_numpy(load_kernels(r, 15)).to(dev)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
best = 1e9
for _ in range(args.reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = ops.kpconv_forward(s, s, idx, x, w, kp, ext)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
by = n * H * (4 + 16) + n * (12 + 4 * cout) + 4 * 15 * cout + 180
print(f"stem gen={os.environ.get('SPR_STEM_GEN', '2')} Nq={n} H={H} Cout={cout}: {best:.3f} ms (pack + kernel), {by / best / 1e6:.0f} GB/s algorithmic")
if args.check:
    import oracle
    rows = rng.choice(n, size=args.check, replace=False)
    exact = oracle.kpconv_forward(s[rows].cpu().numpy(), s.cpu().numpy(), idx[rows].cpu().numpy().astype(np.int64),
                                  x.cpu().numpy(), w.cpu().numpy(), kp.cpu().numpy(), ext)
    err = np.abs(out[rows].cpu().numpy() - exact).max()
    print(f"max |out - oracle| over {args.check} rows: {err:.3e} = {err / np.abs(exact).max():.2e} x max|out|")
