"""Runs the forward repeatedly on the same input and reports which stage output changes between runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import superpoints_registration_b200 as spr
from superpoints_registration_b200 import synthetic, ops
kind = sys.argv[1] if len(sys.argv) > 1 else "3dmatch"
dev = "cuda:0"
cfg = {"3dmatch": spr.threedmatch_config, "modelnet": spr.modelnet_config, "kitti": spr.kitti_config}[kind]()
torch.manual_seed(3); np.random.seed(3)
model = spr.RegTR(cfg).to(dev).eval(); model.return_attn = False
kw = {"n_points": 6000} if kind != "modelnet" else {}
data = synthetic.make_batch(kind, 2, seed=7, **kw)
clouds = [torch.from_numpy(c).to(dev) for c in data["src_xyz"] + data["tgt_xyz"]]
B = 2
ref = None
for it in range(12):
    with torch.no_grad():
        meta = model.preprocessor(list(clouds))
        feats0 = torch.ones_like(meta["points"][0][:, 0:1])
        enc, _ = model.kpf_encoder(feats0, meta)
        both = ops.linear_tc(enc, model.feat_proj.weight, model.feat_proj.bias)
        lens = meta["stack_lengths"][-1].tolist()
        pe = model.pos_embed(meta["points"][-1])
        cond = model.transformer_encoder.forward_packed(both, pe, lens)
    cur = {"nbr0": meta["neighbors"][0].clone(), "pool0": meta["pools"][0].clone(), "pts1": meta["points"][1].clone(),
           "enc": enc.clone(), "proj": both.clone(), "cond": cond.clone()}
    torch.cuda.synchronize()
    if ref is None:
        ref = cur
    else:
        msg = []
        for k in cur:
            if cur[k].shape != ref[k].shape:
                msg.append(f"{k}: SHAPE {tuple(cur[k].shape)} vs {tuple(ref[k].shape)}")
            elif not torch.equal(cur[k], ref[k]):
                d = (cur[k].double() - ref[k].double()).abs().max().item()
                msg.append(f"{k}: max diff {d:.3e}")
        print(f"run {it}: " + ("identical" if not msg else "; ".join(msg)))
