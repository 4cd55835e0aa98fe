"""KPConv-only micro-benchmark: builds a 3DMatch-shape pyramid on the GPU and runs every KPConv layer shape of
the encoder a few times, printing CUDA-event timings.  Short enough to sit under ncu.
    python tools/kpconv_bench.py [--pairs 2] [--reps 3]"""
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
ap.add_argument("--pairs", type=int, default=2)
ap.add_argument("--points", type=int, default=20000)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--cfg", default="3dmatch")
ap.add_argument("--mode", type=int, default=None)
args = ap.parse_args()
dev = "cuda:0"
cfg = {"3dmatch": spr.threedmatch_config, "kitti": spr.kitti_config, "modelnet": spr.modelnet_config}[args.cfg]()
np.random.seed(0)
kw = dict(n_points=args.points) if args.cfg != "modelnet" else {}
data = synthetic.make_batch(args.cfg, args.pairs, seed=2, **kw)
meta = spr.Preprocessor(cfg)([torch.from_numpy(c).to(dev) for c in data["src_xyz"] + data["tgt_xyz"]])
rng = np.random.default_rng(0)
r0 = cfg.first_subsampling_dl * cfg.conv_radius
L = len(meta["points"])
shapes = []
for l in range(L):
    c = (cfg.first_feats_dim // 4) * 2 ** l
    shapes.append((l, False, c))
    if l + 1 < L:
        shapes.append((l, True, c))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
print("level strided   C       Nq      Ns   H   valid/row   ms(best)    GB/s(alg)  TFLOP/s(fp32, valid nbrs)")
for (l, strided, c) in shapes:
    r = r0 * 2 ** l
    ext = r * cfg.KP_extent / cfg.conv_radius
    s = meta["points"][l]
    q = meta["points"][l + 1] if strided else s
    idx = meta["pools"][l] if strided else meta["neighbors"][l]
    ns, nq, H = s.shape[0], q.shape[0], idx.shape[1]
    x = torch.from_numpy(np.where(rng.uniform(size=(ns, c)) < 0.6, rng.normal(size=(ns, c)), -0.05).astype(np.float32)).to(dev)
    w = torch.from_numpy((rng.normal(size=(15, c, c)) / np.sqrt(15 * c)).astype(np.float32)).to(dev)
    kp = torch.from_numpy(load_kernels(r, 15)).to(dev)
    valid = float((idx < ns).sum().item()) / nq
    best = 1e9
    for _ in range(args.reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.kpconv_forward(q, s, idx, x, w, kp, ext, mode=args.mode)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    bytes_alg = nq * H * (4 * c + 16) + nq * (12 + 4 * c) + 4 * 15 * c * c + 180
    flops = 2.0 * nq * 15 * c * c + 2.0 * nq * 15 * valid * c
    print(f"{l:5d} {str(strided):7s} {c:4d} {nq:8d} {ns:8d} {H:3d} {valid:9.1f} {best:10.3f} {bytes_alg / best / 1e6:12.1f} {flops / best / 1e9:10.2f}")
