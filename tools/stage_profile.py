"""Per-stage timing of the forward pass (host wall time with a device sync after every stage, and CUDA-event
device time).  Development aid; not part of the bench contract.  python tools/stage_profile.py [--pairs 8]"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import superpoints_registration_b200 as spr
from superpoints_registration_b200 import _lib, model as M, ops, synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--pairs", type=int, default=8)
ap.add_argument("--points", type=int, default=20000)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--cfg", default="3dmatch")
args = ap.parse_args()
dev = "cuda:0"
cfg = {"3dmatch": spr.threedmatch_config, "kitti": spr.kitti_config, "modelnet": spr.modelnet_config}[args.cfg]()
torch.manual_seed(0); np.random.seed(0)
model = spr.RegTR(cfg).to(dev).eval()
model.return_attn = False
kind = {"3dmatch": "3dmatch", "kitti": "kitti", "modelnet": "modelnet"}[args.cfg]
kw = dict(n_points=args.points) if kind != "modelnet" else {}
data = synthetic.make_batch(kind, args.pairs, seed=2, **kw)
batch = {"src_xyz": [torch.from_numpy(c).to(dev) for c in data["src_xyz"]],
         "tgt_xyz": [torch.from_numpy(c).to(dev) for c in data["tgt_xyz"]]}
B = args.pairs


def staged(rec):
    def mark(name, t0, e0):
        e1 = torch.cuda.Event(enable_timing=True); e1.record(); torch.cuda.synchronize()
        rec.setdefault(name, []).append((time.perf_counter() - t0, e0.elapsed_time(e1)))
    def begin():
        torch.cuda.synchronize(); e = torch.cuda.Event(enable_timing=True); e.record(); return time.perf_counter(), e
    with torch.no_grad():
        t, e = begin()
        meta = model.preprocessor(list(batch["src_xyz"]) + list(batch["tgt_xyz"]))
        mark("preprocess", t, e)
        t, e = begin()
        slens_c = meta["stack_lengths"][-1].tolist()
        feats0 = torch.ones_like(meta["points"][0][:, 0:1])
        feats, _ = model.kpf_encoder(feats0, meta)
        mark("encoder", t, e)
        t, e = begin()
        both = model.feat_proj(feats)
        pts_c = meta["points"][-1]
        pe = model.pos_embed(pts_c)
        mark("proj+pe", t, e)
        t, e = begin()
        cond = model.transformer_encoder.forward_packed(both, pe, slens_c)
        mark("transformer", t, e)
        t, e = begin()
        ts = sum(slens_c[:B])
        out = model._match_and_solve(cond[:ts], cond[ts:], pts_c, slens_c[:B], slens_c[B:])
        mark("match+pose", t, e)
    return meta


rec = {}
for _ in range(3):
    staged({})
l0 = _lib.launch_count()
for _ in range(args.iters):
    meta = staged(rec)
print("levels:", [tuple(p.shape) for p in meta["points"]], "widths:", [tuple(n.shape) for n in meta["neighbors"]])
print("superpoints per cloud:", meta["stack_lengths"][-1].tolist())
print(f"our kernel launches per forward: {(_lib.launch_count() - l0) / args.iters:.0f}")
tot_h = tot_d = 0
for k, v in rec.items():
    h = 1e3 * np.median([a for a, _ in v]); d = np.median([b for _, b in v])
    tot_h += h; tot_d += d
    print(f"{k:12s} host {h:8.2f} ms   device {d:8.2f} ms")
print(f"{'total':12s} host {tot_h:8.2f} ms   device {tot_d:8.2f} ms   -> {B / tot_h * 1e3:.1f} pairs/s (staged, with syncs)")
# whole forward, no intermediate syncs
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(args.iters):
    model(dict(batch))
t_host = (time.perf_counter() - t0) / args.iters
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / args.iters
print(f"unstaged forward: {1e3 * dt:.2f} ms -> {B / dt:.1f} pairs/s   (host returns after {1e3 * t_host:.2f} ms per forward)")
# host-side cost alone: profile the Python of one forward
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(3):
    model(dict(batch))
torch.cuda.synchronize(); pr.disable()
st = pstats.Stats(pr); st.sort_stats("cumulative")
import io
buf = io.StringIO(); st.stream = buf; st.print_stats(22)
print("\n".join(buf.getvalue().splitlines()[:45]))
