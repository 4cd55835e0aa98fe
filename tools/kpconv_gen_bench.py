"""KPConv layer shapes of the 4-stage encoder on a 3DMatch-shape pyramid: generation 1 (kpconv_tc.cu) against generations 2
(kpconv_g.cu) and 3 (kpconv_s.cu) through the prepared entry point, CUDA-event timings, L2 flushed.
    python tools/kpconv_gen_bench.py [--pairs 8] [--reps 5] [--gens 1,3]"""
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
ap.add_argument("--pairs", type=int, default=8)
ap.add_argument("--points", type=int, default=20000)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--gens", default="1,2,3")
args = ap.parse_args()
dev = "cuda:0"
cfg = spr.threedmatch_4stage_config()
np.random.seed(0)
data = synthetic.make_batch("3dmatch", args.pairs, seed=2, n_points=args.points)
meta = spr.Preprocessor(cfg)([torch.from_numpy(c).to(dev) for c in data["src_xyz"] + data["tgt_xyz"]])
rng = np.random.default_rng(0)
r0 = cfg.first_subsampling_dl * cfg.conv_radius
L = len(meta["points"])
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
GENS = [int(g) for g in args.gens.split(",")]
print("level strided   C       Nq   H | " + " | ".join(f" gen{g} ms   GB/s" for g in GENS) + " | vs gen1, max|diff|/max|out|")
tot = {g: 0.0 for g in GENS}
for l in range(L):
    c = (cfg.first_feats_dim // 4) * 2 ** l
    for strided in ((False, True) if l + 1 < L else (False,)):
        r = r0 * 2 ** l
        ext = r * cfg.KP_extent / cfg.conv_radius
        s = meta["points"][l]
        q = meta["points"][l + 1] if strided else s
        idx = meta.index("pools" if strided else "neighbors", l)
        lens = meta.lengths32[l]
        order = meta.order[l + 1 if strided else l]
        ns, nq, H = s.shape[0], q.shape[0], idx.shape[1]
        x = torch.from_numpy(rng.normal(size=(ns, c)).astype(np.float32)).to(dev)
        p_il = ops.instance_norm_lrelu_ex(x, lens, slope=0.1, want_f32=False, kpconv_points=s)["kpconv"]
        p_pl = ops.instance_norm_lrelu_ex(x, lens, slope=0.1, want_f32=False, kpconv_points=s, kpconv_planar=True)["kpconv"]
        preps = {1: p_il, 2: p_pl, 3: p_pl}
        gens = [g for g in GENS if g != 2 or ops._lib.lib().spr_kpconv_gather_supported(c, H)]
        w = torch.from_numpy((rng.normal(size=(15, c, c)) / np.sqrt(15 * c)).astype(np.float32)).to(dev)
        kp = torch.from_numpy(load_kernels(r, 15)).to(dev)
        res, outs = {}, {}
        for gen in gens:
            best = 1e9
            for _ in range(args.reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                outs[gen] = ops.kpconv_forward_prepared(q, idx, preps[gen], w, kp, ext, order=order, generation=gen)
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            res[gen] = best
            tot[gen] += best * (1 if strided else 2 if l > 0 else 1)
        by = nq * H * (4 * c + 16) + nq * (12 + 4 * c) + 4 * 15 * c * c + 180
        cols = " | ".join(f"{res[g]:8.3f} {by / res[g] / 1e6:6.0f}" if g in res else " " * 15 for g in GENS)
        rel = "  ".join(f"gen{g} {res[1] / res[g]:4.2f}x {(outs[1] - outs[g]).abs().max().item() / outs[1].abs().max().item():.1e}"
                        for g in gens if g != 1 and 1 in res)
        print(f"{l:5d} {str(strided):7s} {c:4d} {nq:8d} {H:3d} | {cols} | {rel}")
print("encoder sum (two plain layers per level >= 1): " + ", ".join(f"gen{g} {tot[g]:.3f} ms" for g in GENS))
