"""Host-side cost of one forward at the latency configuration (BASELINE configs[0]: one ~5k-point pair, shipped 3-stage yaml):
wall time per forward with and without a device sync per step, device-busy time, and a cProfile of the Python side.
    python tools/host_profile.py [--pairs 1] [--points 5000] [--n 50] [--top 45]"""
import argparse
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import superpoints_registration_b200 as spr
from superpoints_registration_b200 import synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=1)
ap.add_argument("--points", type=int, default=5000)
ap.add_argument("--n", type=int, default=50)
ap.add_argument("--top", type=int, default=45)
ap.add_argument("--arch", default="3stage")
args = ap.parse_args()
dev = "cuda:0"
torch.manual_seed(0)
cfg = spr.threedmatch_config() if args.arch == "3stage" else spr.threedmatch_4stage_config()
model = spr.RegTR(cfg).to(dev).eval()
model.return_attn = False
data = synthetic.make_batch("3dmatch", args.pairs, seed=217, n_points=args.points)
batch = {k: [torch.from_numpy(c).to(dev) for c in data[k]] for k in ("src_xyz", "tgt_xyz")}
for _ in range(10):
    model(dict(batch))
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(args.n):
    model(dict(batch))
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
t1 = time.perf_counter()
for _ in range(args.n):
    model(dict(batch))["pose"].cpu()
t_sync = time.perf_counter() - t1
from superpoints_registration_b200 import _lib
c0 = _lib.lib().spr_launch_count()
model(dict(batch))
torch.cuda.synchronize()
print(f"forward of {args.pairs} pair(s), {args.points} pts, {args.arch}: host issue {1e3 * t_issue / args.n:.3f} ms, "
      f"issue + drain {1e3 * t_all / args.n:.3f} ms, with a pose read-back per step {1e3 * t_sync / args.n:.3f} ms; "
      f"{_lib.lib().spr_launch_count() - c0} library launches")
pr = cProfile.Profile()
pr.enable()
for _ in range(args.n):
    model(dict(batch))
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(args.top)
