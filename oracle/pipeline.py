"""oracle/pipeline.py -- CPU restatement of the multi-stage parts of the hot path.  TEST INFRASTRUCTURE ONLY.

  preprocess(cfg, clouds, backend)   the pyramid builder, models/backbone_kpconv/kpconv.py:302-418
  encoder / transformer / forward    the torch-side forward, models/qk_regtr_full.py:126-311, as plain functions
                                     over a state_dict (torch CPU ops, fp32) -- the reference's Python cannot
                                     travel to the GPU box, so this is what bench.py's CPU baseline times there.

backend="port" uses our C restatement (canonical first-occurrence subsample order); backend="reference" uses
the reference's own C++ core through oracle/_ref (its hash-table order) -- identical index SETS up to the
relabelling tests/parity.py computes.

Parity status: pinned.  tests/test_oracle_golden.py checks every function here against tests/golden/*.npz,
which were produced by the imported reference itself.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence

import numpy as np

import oracle
from oracle import numpy_ops


# ------------------------------------------------------------------------------------------------
# pyramid
# ------------------------------------------------------------------------------------------------

def level_plan(architecture: Sequence[str]):
    """(has_conv, strided) per pyramid level -- the block walk of kpconv.py:333-346, :389-392."""
    plan, pending = [], []
    arch = list(architecture)
    for i, block in enumerate(arch):
        if "global" in block or "upsample" in block:
            break
        closes = "pool" in block or "strided" in block
        if not closes:
            pending.append(block)
            if i < len(arch) - 1 and "upsample" not in arch[i + 1]:
                continue
        plan.append((len(pending) > 0, closes))
        pending = []
    return plan


def preprocess(cfg, clouds: List[np.ndarray], backend: str = "port") -> Dict[str, List[np.ndarray]]:
    """-> dict of lists of NumPy arrays with the reference's keys; index matrices int64, trimmed to
    min(max_count, limit) columns like kpconv.py:259-260; lengths int32."""
    limits = cfg["neighborhood_limits"]
    r = np.float64(cfg["first_subsampling_dl"]) * np.float64(cfg["conv_radius"])  # Python floats in the reference
    pts = np.concatenate([np.asarray(c, np.float32) for c in clouds], 0)
    lens = np.asarray([len(c) for c in clouds], np.int32)

    def neighbors(q, s, ql, sl, radius, limit):
        if backend == "reference":
            full = oracle.ref_batch_query(q, s, ql, sl, float(radius))
            return full[:, :limit].astype(np.int64)
        idx, mc = oracle.radius_neighbors_batch(q, s, ql, sl, float(radius), int(limit))
        return idx[:, :min(mc, limit)].astype(np.int64)

    def subsample(p, l, dl):
        if backend == "reference":
            return oracle.ref_subsample_batch(p, l, float(dl))
        return oracle.grid_subsample_batch(p, l, float(dl))

    out = {k: [] for k in ("points", "neighbors", "pools", "upsamples", "stack_lengths")}
    for level, (has_conv, strided) in enumerate(level_plan(cfg["architecture"])):
        limit = limits[level]
        conv_i = neighbors(pts, pts, lens, lens, r, limit) if has_conv else np.zeros((0, 1), np.int64)   # :353
        if strided:
            dl = 2 * r / cfg["conv_radius"]                                                              # :367
            pool_p, pool_b = subsample(pts, lens, dl)                                                    # :370
            pool_i = neighbors(pool_p, pts, pool_b, lens, r, limit)                                      # :380
            up_i = neighbors(pts, pool_p, lens, pool_b, 2 * r, limit)                                    # :384
        else:
            pool_i, up_i = np.zeros((0, 1), np.int64), np.zeros((0, 1), np.int64)                        # :389-392
            pool_p, pool_b = np.zeros((0, 3), np.float32), np.zeros((0,), np.int64)
        out["points"].append(pts)
        out["neighbors"].append(conv_i)
        out["pools"].append(pool_i)
        out["upsamples"].append(up_i)
        out["stack_lengths"].append(lens)
        pts, lens = pool_p, pool_b
        r = r * 2                                                                                        # :406
    return out


# ------------------------------------------------------------------------------------------------
# torch-CPU restatement of the network forward (fp32), as plain functions over a state_dict
# ------------------------------------------------------------------------------------------------

def _torch():
    import torch
    return torch


def kpconv_dense(q_pts, s_pts, inds, x, weights, kernel_points, extent):
    """kpconv_blocks.py:305-414 with torch CPU ops, materialising the same intermediates the reference does
    ([N,H,K] influences, [N,H,Cin] gathered features) -- this is the cost profile of the reference's CPU path."""
    torch = _torch()
    s_pad = torch.cat([s_pts, torch.full((1, 3), 1e6, dtype=s_pts.dtype)], 0)             # :309
    rel = s_pad[inds] - q_pts[:, None, :]                                                 # :312-315
    sq = ((rel[:, :, None, :] - kernel_points[None, None]) ** 2).sum(-1)                  # :324-328  [N,H,K]
    infl = torch.clamp(1 - torch.sqrt(sq) / extent, min=0.0).transpose(1, 2)              # :368-369  [N,K,H]
    x_pad = torch.cat([x, torch.zeros_like(x[:1])], 0)                                    # :388
    nx = x_pad[inds]                                                                      # :391      [N,H,Cin]
    wf = torch.matmul(infl, nx)                                                           # :394      [N,K,Cin]
    out = torch.matmul(wf.permute(1, 0, 2), weights).sum(0)                               # :401-406
    nn_ = (nx.sum(-1) > 0).sum(-1).clamp(min=1)                                           # :409-411
    return out / nn_[:, None]                                                             # :412


def _inorm(x, lens):
    """kpconv_blocks.py:510-519: one nn.InstanceNorm1d call per cloud."""
    torch = _torch()
    F = torch.nn.functional
    outs, o = [], 0
    for n in lens:
        seg = x[o:o + n].t()[None]                       # (1, C, L)
        outs.append(F.instance_norm(seg, eps=1e-5)[0].t())
        o += n
    return torch.cat(outs, 0)


def _unary(sd, prefix, x, lens, relu=True):
    torch = _torch()
    y = _inorm(x @ sd[prefix + ".mlp.weight"].t(), lens)
    return torch.nn.functional.leaky_relu(y, 0.1) if relu else y


def encoder_forward(sd, cfg, meta, prefix="kpf_encoder.encoder_blocks."):
    """kpconv.py:22-92 + kpconv_blocks.py:590-741 for 'simple' / 'resnetb' / 'resnetb_strided' blocks."""
    torch = _torch()
    F = torch.nn.functional
    lens = [m.tolist() for m in meta["stack_lengths"]]
    r = cfg["first_subsampling_dl"] * cfg["conv_radius"]
    level = 0
    x = torch.ones((meta["points"][0].shape[0], 1))
    for i, name in enumerate(cfg["architecture"]):
        p = f"{prefix}{i}."
        strided = "strided" in name
        q_pts = meta["points"][level + 1] if strided else meta["points"][level]
        s_pts = meta["points"][level]
        inds = meta["pools"][level] if strided else meta["neighbors"][level]
        post = lens[level + 1] if strided else lens[level]
        extent = r * cfg["KP_extent"] / cfg["conv_radius"]
        conv = lambda feats: kpconv_dense(q_pts, s_pts, inds, feats, sd[p + "KPConv.weights"],
                                          sd[p + "KPConv.kernel_points"], extent)
        if name.startswith("simple"):
            x = F.leaky_relu(_inorm(conv(x), post), 0.1)
        elif name.startswith("resnetb"):
            feats = x
            y = _unary(sd, p + "unary1", feats, lens[level]) if (p + "unary1.mlp.weight") in sd else feats
            y = F.leaky_relu(_inorm(conv(y), post), 0.1)
            y = _unary(sd, p + "unary2", y, post, relu=False)
            if strided:
                fp = torch.cat([feats, torch.zeros_like(feats[:1])], 0)
                sc = fp[inds].max(1)[0]                                                   # max_pool :127-143
            else:
                sc = feats
            if (p + "unary_shortcut.mlp.weight") in sd:
                sc = _unary(sd, p + "unary_shortcut", sc, post, relu=False)
            x = F.leaky_relu(y + sc, 0.1)
        else:
            raise NotImplementedError(name)
        if strided:
            level += 1
            r *= 2
    return x


def _pos_embed(xyz, d_model, scale=1.0, temperature=10000):
    """models/transformer/position_embedding.py:7-50."""
    torch = _torch()
    nf = d_model // 3 // 2 * 2
    k = torch.arange(nf, dtype=torch.float32)
    freq = temperature ** (2 * torch.div(k, 2, rounding_mode="trunc") / nf)
    ang = (xyz * (scale * 2 * math.pi))[..., None] / freq
    emb = torch.stack([ang[..., 0::2].sin(), ang[..., 1::2].cos()], -1).reshape(*xyz.shape[:-1], -1)
    return torch.nn.functional.pad(emb, (0, d_model - nf * 3))


def _mha(sd, p, q, k, v, mask, nhead):
    torch = _torch()
    F = torch.nn.functional
    return F.multi_head_attention_forward(
        q, k, v, q.shape[-1], nhead, sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"], None, None, False, 0.0,
        sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"], training=False, key_padding_mask=mask,
        need_weights=True)[0]


def transformer_forward(sd, cfg, src, tgt, src_mask, tgt_mask, src_pos, tgt_pos, prefix="transformer_encoder."):
    """models/transformer/transformers.py:27-59 with the pre-norm layer :184-245 (sequence-first tensors)."""
    torch = _torch()
    F = torch.nn.functional
    d, H = cfg["d_embed"], cfg["nhead"]
    ln = lambda x, p: F.layer_norm(x, (d,), sd[p + ".weight"], sd[p + ".bias"])
    assert cfg["pre_norm"] and cfg["sa_val_has_pos_emb"] and cfg["ca_val_has_pos_emb"]
    for l in range(cfg["num_encoder_layers"]):
        p = f"{prefix}layers.{l}."
        s2 = ln(src, p + "norm1") + src_pos
        src = src + _mha(sd, p + "self_attn", s2, s2, s2, src_mask, H)
        t2 = ln(tgt, p + "norm1") + tgt_pos
        tgt = tgt + _mha(sd, p + "self_attn", t2, t2, t2, tgt_mask, H)
        s2, t2 = ln(src, p + "norm2") + src_pos, ln(tgt, p + "norm2") + tgt_pos
        s3 = _mha(sd, p + "multihead_attn", s2, t2, t2, tgt_mask, H)
        t3 = _mha(sd, p + "multihead_attn", t2, s2, s2, src_mask, H)
        src, tgt = src + s3, tgt + t3
        ffn = lambda x: F.linear(F.relu(F.linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"])),
                                 sd[p + "linear2.weight"], sd[p + "linear2.bias"])
        src = src + ffn(ln(src, p + "norm3"))
        tgt = tgt + ffn(ln(tgt, p + "norm3"))
    return ln(src, prefix + "norm"), ln(tgt, prefix + "norm")


def forward(sd, cfg, src_clouds, tgt_clouds, backend="port", timings=None):
    """models/qk_regtr_full.py:126-311 on the CPU: pyramid -> encoder -> transformer -> matching -> pose.
    sd: state_dict of torch CPU tensors (reference key names).  Returns dict with pose [B,3,4] (NumPy) etc."""
    import time
    torch = _torch()
    B = len(src_clouds)
    t0 = time.perf_counter()
    meta_np = preprocess(cfg, list(src_clouds) + list(tgt_clouds), backend=backend)
    meta = {k: [torch.from_numpy(np.ascontiguousarray(a)) for a in v] for k, v in meta_np.items()}
    t1 = time.perf_counter()
    with torch.no_grad():
        feats = encoder_forward(sd, cfg, meta)
        t2 = time.perf_counter()
        both = feats @ sd["feat_proj.weight"].t() + sd["feat_proj.bias"]
        lens = meta_np["stack_lengths"][-1].tolist()
        pts_c = meta["points"][-1]
        parts, xyzs = torch.split(both, lens), torch.split(pts_c, lens)
        pes = torch.split(_pos_embed(pts_c, cfg["d_embed"]), lens)
        pad = torch.nn.utils.rnn.pad_sequence

        def mask_of(seqs):
            n = torch.tensor([len(s) for s in seqs])
            return torch.arange(int(n.max()))[None, :] >= n[:, None]
        src, tgt = pad(parts[:B]), pad(parts[B:])
        sm, tm = mask_of(parts[:B]), mask_of(parts[B:])
        sc, tc = transformer_forward(sd, cfg, src, tgt, sm, tm, pad(pes[:B]), pad(pes[B:]))
        t3 = time.perf_counter()
        poses, inds, vals, sfeat, tfeat = [], [], [], [], []
        sp_alpha = float(torch.nn.functional.softplus(sd["alpha"]))
        e_beta = float(torch.exp(sd["beta"]))
        for b in range(B):
            S, T = sc[:lens[b], b].numpy(), tc[:lens[B + b], b].numpy()
            sx, tx = xyzs[b].numpy(), xyzs[B + b].numpy()
            corr, attn, val, ind = numpy_ops.dual_softmax_match(S, T)
            if cfg["use_sinkhorn"]:
                wt, w = numpy_ops.sinkhorn_weighted_targets(corr, tx, sp_alpha, e_beta, cfg["sinkhorn_itr"])
                pose = numpy_ops.compute_rigid_transform(sx, wt, w)
            elif len(S) > len(T):
                pose = numpy_ops.compute_rigid_transform(sx[ind], tx, val)
            else:
                pose = numpy_ops.compute_rigid_transform(sx, tx[ind], val)
            poses.append(pose)
            inds.append(ind)
            vals.append(val)
            sfeat.append(S)
            tfeat.append(T)
        t4 = time.perf_counter()
    if timings is not None:
        timings.update(preprocess=t1 - t0, encoder=t2 - t1, transformer=t3 - t2, matching_pose=t4 - t3)
    return {"pose": np.stack(poses), "ind": inds, "val": vals, "src_feat": sfeat, "tgt_feat": tfeat,
            "encoder_out": feats.numpy(), "meta": meta_np}
