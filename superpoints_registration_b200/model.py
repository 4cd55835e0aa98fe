"""The registration forward pass (hot path only) on B200.

Mirrors models/qk_regtr_full.py of the reference: `RegTR.forward` (:126-311) and
`RegTR.softmax_correlation` (:423-672), with the same sub-module names so that a reference checkpoint
loads key for key (`kpf_encoder.*`, `feat_proj.*`, `transformer_encoder.*`, `overlap_predictor.*`,
`alpha`, `beta`).  Training-only members (losses, metrics, optimiser plumbing) are out of scope.

What runs where:
  preprocessor, kpf_encoder (KPConv, unary Linear layers, InstanceNorm), feat_proj, the 6-layer cross-attention
  transformer (packed tokens: LayerNorm, projections and FFN on the tcgen05 GEMM, varlen flash-attention),
  superpoint matching, Sinkhorn, pose solve  -> our sm_100a kernels (libspr_b200.so);
  PyTorch supplies memory, streams and a few element-wise glue ops (sine position embedding, sigmoid).
  Configurations outside the shipped ones (post-norm, values without positional embedding) keep the padded
  nn.MultiheadAttention modules on the GPU.
The per-pair Python loop of the reference (:445) is gone: all pairs of the batch go through the same launches.
"""
from __future__ import annotations

import copy
import math
from typing import Dict, List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .kpconv import KPFEncoder, Preprocessor

# use_attn_affinity raises ValueError inside the reference itself (qk_regtr_full.py:509-513,619-623) and
# use_corr_affinity is shape-inconsistent there for N > M (:516-519); neither is reproduced.
_UNSUPPORTED_FLAGS = ("use_attn_affinity", "use_corr_affinity")


# ------------------------------------------------------------------------------------------------
# sequence helpers (utils/seq_manipulation.py:6-48)
# ------------------------------------------------------------------------------------------------

def split_src_tgt(feats, stack_lengths, dim=0):
    if isinstance(stack_lengths, torch.Tensor):
        stack_lengths = stack_lengths.tolist()
    half = len(stack_lengths) // 2
    parts = torch.split(feats, stack_lengths, dim=dim)
    return parts[:half], parts[half:]


def pad_sequence(sequences, require_padding_mask=False, require_lens=False, batch_first=False):
    padded = nn.utils.rnn.pad_sequence(sequences, batch_first=batch_first)
    mask = None
    if require_padding_mask:
        lens = torch.tensor([s.shape[0] for s in sequences], device=padded.device)
        steps = padded.shape[1] if batch_first else padded.shape[0]
        mask = torch.arange(steps, device=padded.device)[None, :] >= lens[:, None]
    lens_out = [s.shape[0] for s in sequences] if require_lens else None
    return padded, mask, lens_out


def unpad_sequences(padded, seq_lens):
    return [padded[..., :seq_lens[b], b, :] for b in range(len(seq_lens))]


# ------------------------------------------------------------------------------------------------
# transformer (models/transformer/position_embedding.py:7-50, transformers.py:18-259) -- PyTorch
# ------------------------------------------------------------------------------------------------

class PositionEmbeddingCoordsSine(nn.Module):
    def __init__(self, n_dim: int = 1, d_model: int = 256, temperature=10000, scale=None):
        super().__init__()
        self.n_dim = n_dim
        self.num_pos_feats = d_model // n_dim // 2 * 2
        self.temperature = temperature
        self.padding = d_model - self.num_pos_feats * self.n_dim
        self.scale = (1.0 if scale is None else scale) * 2 * math.pi

    def forward(self, xyz: torch.Tensor) -> torch.Tensor:
        assert xyz.shape[-1] == self.n_dim
        k = torch.arange(self.num_pos_feats, dtype=torch.float32, device=xyz.device)
        freq = self.temperature ** (2 * torch.div(k, 2, rounding_mode='trunc') / self.num_pos_feats)
        ang = (xyz * self.scale).unsqueeze(-1) / freq
        emb = torch.stack([ang[..., 0::2].sin(), ang[..., 1::2].cos()], dim=-1).reshape(*xyz.shape[:-1], -1)
        return F.pad(emb, (0, self.padding))


class TransformerCrossEncoderLayer(nn.Module):
    def __init__(self, d_model, nhead, dim_feedforward=2048, dropout=0.1, activation="relu", normalize_before=False,
                 sa_val_has_pos_emb=False, ca_val_has_pos_emb=False, attention_type='dot_prod', batch_first=False):
        super().__init__()
        if attention_type != 'dot_prod':
            raise NotImplementedError
        if dropout != 0.0:
            raise NotImplementedError("inference path: dropout must be 0 (as in every shipped config)")
        self.self_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout, batch_first=batch_first)
        self.multihead_attn = nn.MultiheadAttention(d_model, nhead, dropout=dropout, batch_first=batch_first)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self.norm1 = nn.LayerNorm(d_model)
        self.norm2 = nn.LayerNorm(d_model)
        self.norm3 = nn.LayerNorm(d_model)
        self.activation = {"relu": F.relu, "gelu": F.gelu}[activation]
        self.normalize_before = normalize_before
        self.sa_val_has_pos_emb = sa_val_has_pos_emb
        self.ca_val_has_pos_emb = ca_val_has_pos_emb

    @staticmethod
    def _pe(x, pos):
        return x if pos is None else x + pos

    def _sa(self, x, pos, mask):
        xp = self._pe(x, pos)
        return self.self_attn(xp, xp, value=xp if self.sa_val_has_pos_emb else x, key_padding_mask=mask,
                              need_weights=False)[0]

    def _ca(self, q, q_pos, kv, kv_pos, kv_mask):
        kp = self._pe(kv, kv_pos)
        return self.multihead_attn(query=self._pe(q, q_pos), key=kp, value=kp if self.ca_val_has_pos_emb else kv,
                                   key_padding_mask=kv_mask, need_weights=False)[0]

    def _ffn(self, x):
        return self.linear2(self.activation(self.linear1(x)))

    # ---- packed-token path: our attention kernel, no padding, src and tgt clouds in one launch -------------
    def _attend_packed(self, mha, xp, tiles):
        d = xp.shape[1]
        qkv = F.linear(xp, mha.in_proj_weight, mha.in_proj_bias)                     # [T, 3d]
        hi, lo = ops.split_f16(qkv, n_scaled=d, scale=math.log2(math.e) / math.sqrt(d // mha.num_heads))
        o = ops.attention_varlen(hi, lo, tiles, mha.num_heads, 0, d, 2 * d, d)
        return F.linear(o, mha.out_proj.weight, mha.out_proj.bias)

    def forward_packed(self, x, pos, sa_tiles, ca_tiles):
        """x, pos: [total_tokens, d] with the clouds stacked [src_0..src_B-1, tgt_0..tgt_B-1]; same arithmetic as
        forward_pre (transformers.py:184-245) with value = key (sa_val_has_pos_emb / ca_val_has_pos_emb = True)."""
        if not (self.normalize_before and self.sa_val_has_pos_emb and self.ca_val_has_pos_emb):
            raise NotImplementedError("packed cross-encoder: pre-norm with positional values only (every shipped config)")
        xn = self.norm1(x)
        x = x + self._attend_packed(self.self_attn, xn if pos is None else xn + pos, sa_tiles)
        xn = self.norm2(x)
        x = x + self._attend_packed(self.multihead_attn, xn if pos is None else xn + pos, ca_tiles)
        return x + self._ffn(self.norm3(x))

    def forward_fused(self, x, pos, sa_tiles, ca_tiles, img, img_ffn):
        """Same arithmetic as forward_packed with every dense layer on the tcgen05 GEMM (ops.gemm_tc) and the
        element-wise work folded into producers / epilogues: LayerNorm + positional embedding write the GEMM operand
        image, the QKV projection writes the fp16 planes of the attention kernel, attention writes the operand image
        of the output projection, the output projection / FFN2 add the residual, FFN1 applies ReLU and writes FFN2's
        operand image.  x is updated in place; 11 launches per layer, no intermediate fp32 activation besides x."""
        if self.activation is not F.relu:
            raise NotImplementedError("fused cross-encoder: ReLU feed-forward only (every shipped config)")
        T, d = x.shape
        for mha, norm, tiles in ((self.self_attn, self.norm1, sa_tiles), (self.multihead_attn, self.norm2, ca_tiles)):
            ops.layernorm256_prepare(x, norm.weight, norm.bias, pos, norm.eps, img)
            hi, lo = ops.gemm_tc(img, ops.weight_image(mha.in_proj_weight), mha.in_proj_bias, T, ops.OUT_PLANES,
                                 n_scaled=d, col_scale=math.log2(math.e) / math.sqrt(d // mha.num_heads))
            ops.attention_varlen(hi, lo, tiles, mha.num_heads, 0, d, 2 * d, d, out_image=img, image_scale=ops.A_SCALE)
            ops.gemm_tc(img, ops.weight_image(mha.out_proj.weight), mha.out_proj.bias, T, ops.OUT_F32, residual=x, out=x)
        ops.layernorm256_prepare(x, self.norm3.weight, self.norm3.bias, None, self.norm3.eps, img)
        ops.gemm_tc(img, ops.weight_image(self.linear1.weight), self.linear1.bias, T, ops.OUT_AIMG, relu=True, out=img_ffn)
        ops.gemm_tc(img_ffn, ops.weight_image(self.linear2.weight), self.linear2.bias, T, ops.OUT_F32, residual=x, out=x)
        return x

    def forward(self, src, tgt, src_key_padding_mask=None, tgt_key_padding_mask=None, src_pos=None, tgt_pos=None):
        sm, tm = src_key_padding_mask, tgt_key_padding_mask
        if self.normalize_before:  # transformers.py:184-245
            src = src + self._sa(self.norm1(src), src_pos, sm)
            tgt = tgt + self._sa(self.norm1(tgt), tgt_pos, tm)
            s2, t2 = self.norm2(src), self.norm2(tgt)
            src, tgt = src + self._ca(s2, src_pos, t2, tgt_pos, tm), tgt + self._ca(t2, tgt_pos, s2, src_pos, sm)
            src = src + self._ffn(self.norm3(src))
            tgt = tgt + self._ffn(self.norm3(tgt))
            return src, tgt
        # post-norm, transformers.py:117-182
        src = self.norm1(src + self._sa(src, src_pos, sm))
        tgt = self.norm1(tgt + self._sa(tgt, tgt_pos, tm))
        s3, t3 = self._ca(src, src_pos, tgt, tgt_pos, tm), self._ca(tgt, tgt_pos, src, src_pos, sm)
        src, tgt = self.norm2(src + s3), self.norm2(tgt + t3)
        src = self.norm3(src + self._ffn(src))
        tgt = self.norm3(tgt + self._ffn(tgt))
        return src, tgt


class TransformerCrossEncoder(nn.Module):
    def __init__(self, cross_encoder_layer, num_layers, norm=None, return_intermediate=False):
        super().__init__()
        if return_intermediate:
            raise NotImplementedError
        self.layers = nn.ModuleList([copy.deepcopy(cross_encoder_layer) for _ in range(num_layers)])
        self.num_layers = num_layers
        self.norm = norm
        self.fused = True  # dense layers on the tcgen05 GEMM (False: cuBLAS fp32 through torch)
        self.native_sequencer = True  # all layers from one library call (False: one Python call per launch)
        self.cuda_graphs = True       # small inputs: the sequencer's launches replayed as a CUDA graph per shape bucket

    def forward(self, src, tgt, src_key_padding_mask=None, tgt_key_padding_mask=None, src_pos=None, tgt_pos=None):
        for layer in self.layers:
            src, tgt = layer(src, tgt, src_key_padding_mask=src_key_padding_mask,
                             tgt_key_padding_mask=tgt_key_padding_mask, src_pos=src_pos, tgt_pos=tgt_pos)
        if self.norm is not None:
            src, tgt = self.norm(src), self.norm(tgt)
        return src.unsqueeze(0), tgt.unsqueeze(0)

    def supports_packed(self) -> bool:
        l0 = self.layers[0]
        return bool(l0.normalize_before and l0.sa_val_has_pos_emb and l0.ca_val_has_pos_emb
                    and l0.self_attn.embed_dim // l0.self_attn.num_heads == 32)

    def forward_packed(self, x, pos, lens):
        """x, pos: [total_tokens, d], clouds stacked [src..., tgt...]; lens: token count of every cloud (2B entries).
        Returns the conditioned features in the same packed layout."""
        B = len(lens) // 2
        offs = [0]
        for n in lens:
            offs.append(offs[-1] + n)
        partner = list(range(B, 2 * B)) + list(range(0, B))
        native = (self.fused and self.native_sequencer and x.shape[1] == 256
                  and all(l.activation is F.relu for l in self.layers))
        graphed = (native and self.cuda_graphs and x.shape[0] <= ops.ENCODER_GRAPH_MAX_ROWS
                   and not torch.cuda.is_current_stream_capturing())  # (a caller's own capture takes the eager launches)
        pad = ops.ENCODER_GRAPH_TILE_PAD if graphed else 1
        sa_tiles = ops.attention_tiles(offs[:-1], lens, offs[:-1], lens, x.device, pad_multiple=pad)
        ca_tiles = ops.attention_tiles(offs[:-1], lens, [offs[p] for p in partner], [lens[p] for p in partner], x.device,
                                       pad_multiple=pad)
        if native:
            # one library call for all layers (csrc/encoder_seq.cu): same launches, same order, no interpreter in
            # between -- and, for small inputs, replayed as one CUDA graph per shape bucket
            table = getattr(self, '_seq_table', None)
            if table is None or table.key != ops.CrossEncoderTable.signature(self.layers):
                table = self._seq_table = ops.CrossEncoderTable(self.layers)
            run = ops.cross_encoder_forward_graphed if graphed else ops.cross_encoder_forward
            return run(x, pos, table, self.layers[0].self_attn.num_heads, sa_tiles, ca_tiles, self.norm)
        if self.fused and x.shape[1] == 256:
            x = x.clone()  # updated in place layer by layer
            img = ops.gemm_a_image(x.shape[0], 256, x.device)
            img_ffn = ops.gemm_a_image(x.shape[0], self.layers[0].linear1.out_features, x.device)
            for layer in self.layers:
                x = layer.forward_fused(x, pos, sa_tiles, ca_tiles, img, img_ffn)
            if self.norm is None:
                return x
            return ops.layernorm256_prepare(x, self.norm.weight, self.norm.bias, None, self.norm.eps, None, out_f32=True)
        for layer in self.layers:
            x = layer.forward_packed(x, pos, sa_tiles, ca_tiles)
        return self.norm(x) if self.norm is not None else x


# ------------------------------------------------------------------------------------------------
# model
# ------------------------------------------------------------------------------------------------

class RegTR(nn.Module):
    """Inference-path twin of models/qk_regtr_full.RegTR."""

    def __init__(self, cfg, *args, **kwargs):
        super().__init__()
        self.cfg = cfg
        for flag in _UNSUPPORTED_FLAGS:
            if cfg.get(flag, False):
                raise NotImplementedError(f"{flag}=True: not usable in the reference either")
        if cfg.get('use_sinkhorn', False) and (cfg.get('use_lgr', False) or cfg.get('use_ransac', False)):
            # the reference hands all N source and all M target points to the refinement (:538-542), which only
            # type-checks when N == M
            raise NotImplementedError("use_lgr / use_ransac need one-to-one correspondences: not with use_sinkhorn")
        if cfg.get('use_overlap_as_weights', False) and not cfg.get('remove_outliers_overlap', False):
            raise NotImplementedError("use_overlap_as_weights needs remove_outliers_overlap (overlap_prob, :481-487)")
        if cfg.get('use_overlap_as_weights', False) and cfg.get('remove_points_from_val', False):
            raise NotImplementedError("use_overlap_as_weights with remove_points_from_val: weights and points differ "
                                      "in length in the reference (:497-500, :546)")
        # cfg.preprocessor = 'gpu_compat' selects the PreprocessorGPU-compatible pyramid (what the reference's model
        # instantiates, qk_regtr_full.py:40); the default is the CPU Preprocessor north_star names as the oracle
        self.preprocessor = Preprocessor(cfg, mode=cfg.get('preprocessor', 'reference'))
        self.kpf_encoder = KPFEncoder(cfg, cfg.d_embed)
        self.feat_proj = nn.Linear(self.kpf_encoder.encoder_skip_dims[-1], cfg.d_embed, bias=True)
        if cfg.get('pos_emb_type', 'sine') != 'sine':
            raise NotImplementedError("learned position embedding")
        self.pos_embed = PositionEmbeddingCoordsSine(3, cfg.d_embed, scale=cfg.get('pos_emb_scaling', 1.0))
        layer = TransformerCrossEncoderLayer(
            cfg.d_embed, cfg.nhead, cfg.d_feedforward, cfg.dropout, activation=cfg.transformer_act,
            normalize_before=cfg.pre_norm, sa_val_has_pos_emb=cfg.sa_val_has_pos_emb,
            ca_val_has_pos_emb=cfg.ca_val_has_pos_emb, attention_type=cfg.attention_type)
        self.transformer_encoder = TransformerCrossEncoder(
            layer, cfg.num_encoder_layers, nn.LayerNorm(cfg.d_embed) if cfg.pre_norm else None)
        self.beta = nn.Parameter(torch.tensor(1.0))
        self.alpha = nn.Parameter(torch.tensor(1.0))
        self.overlap_predictor = nn.Linear(cfg.d_embed, 1)
        self.dual_normalization = True  # hard-coded in the reference (:120)
        self.return_attn = True          # outputs['attn'] (:295); set False to skip materialising N x M matrices
        # RegTR.ransac draws 500 x 100 rows from the global CUDA generator (:402-405); pass a torch.Generator to
        # make a run reproducible
        self.ransac_hypotheses, self.ransac_sample_size, self.ransac_generator = 500, 100, None
        self.packed_transformer = True   # packed tokens + our attention kernel; False = padded PyTorch modules

    def load_reference_state_dict(self, state_dict):
        """Load a reference checkpoint; training-only keys (loss modules) are ignored."""
        own = self.state_dict()
        picked = {k: v for k, v in state_dict.items() if k in own}
        missing = [k for k in own if k not in picked]
        if missing:
            raise RuntimeError(f"reference state_dict lacks keys: {missing[:5]}...")
        self.load_state_dict(picked, strict=True)

    # -------------------------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, batch):
        cfg = self.cfg
        B = len(batch['src_xyz'])
        meta = self.preprocessor(list(batch['src_xyz']) + list(batch['tgt_xyz']))
        batch['kpconv_meta'] = meta
        host_lens = getattr(meta, 'host_lengths', None)
        slens_c = list(host_lens[-1]) if host_lens else meta['stack_lengths'][-1].tolist()
        src_slens_c, tgt_slens_c = slens_c[:B], slens_c[B:]
        pts_c = meta['points'][-1]
        feats0 = torch.ones_like(meta['points'][0][:, 0:1])

        feats_un, _ = self.kpf_encoder(feats0, meta)
        stash = meta.get('_operand_image')  # the last encoder block wrote its output as a GEMM operand image too
        if stash is not None and stash[0] is feats_un:
            both = ops.gemm_tc(stash[1], ops.weight_image(self.feat_proj.weight), self.feat_proj.bias, feats_un.shape[0])
        else:
            both = ops.linear_tc(feats_un, self.feat_proj.weight, self.feat_proj.bias)
        src_xyz_c, tgt_xyz_c = split_src_tgt(pts_c, slens_c)
        use_pe = cfg.transformer_encoder_has_pos_emb
        pe = self.pos_embed(pts_c)
        if self.packed_transformer and self.transformer_encoder.supports_packed():
            # tokens stay packed [src_0..src_B-1, tgt_0..tgt_B-1]: no padding, no masks, one launch per operator
            cond = self.transformer_encoder.forward_packed(both, pe if use_pe else None, slens_c)
            overlap = torch.sigmoid(self.overlap_predictor(cond))
            total_src_c = sum(src_slens_c)
            src_packed, tgt_packed = cond[:total_src_c], cond[total_src_c:]
            src_cond_list = [c.unsqueeze(0) for c in torch.split(src_packed, src_slens_c)]
            tgt_cond_list = [c.unsqueeze(0) for c in torch.split(tgt_packed, tgt_slens_c)]
            src_overlap_list = [c.unsqueeze(0) for c in torch.split(overlap[:total_src_c], src_slens_c)]
            tgt_overlap_list = [c.unsqueeze(0) for c in torch.split(overlap[total_src_c:], tgt_slens_c)]
        else:
            src_feats_un, tgt_feats_un = split_src_tgt(both, slens_c)
            src_pe, tgt_pe = split_src_tgt(pe, slens_c)
            src_pe_pad, _, _ = pad_sequence(src_pe)
            tgt_pe_pad, _, _ = pad_sequence(tgt_pe)
            src_pad, src_mask, _ = pad_sequence(src_feats_un, require_padding_mask=True)
            tgt_pad, tgt_mask, _ = pad_sequence(tgt_feats_un, require_padding_mask=True)
            src_cond, tgt_cond = self.transformer_encoder(
                src_pad, tgt_pad, src_key_padding_mask=src_mask, tgt_key_padding_mask=tgt_mask,
                src_pos=src_pe_pad if use_pe else None, tgt_pos=tgt_pe_pad if use_pe else None)
            src_overlap = torch.sigmoid(self.overlap_predictor(src_cond))
            tgt_overlap = torch.sigmoid(self.overlap_predictor(tgt_cond))
            src_overlap_list = unpad_sequences(src_overlap, src_slens_c)
            tgt_overlap_list = unpad_sequences(tgt_overlap, tgt_slens_c)
            src_cond_list = unpad_sequences(src_cond, src_slens_c)
            tgt_cond_list = unpad_sequences(tgt_cond, tgt_slens_c)
            # packed (sum N, D) features in pair order for the batched matching kernels
            src_packed = src_cond[0].transpose(0, 1)[~src_mask]
            tgt_packed = tgt_cond[0].transpose(0, 1)[~tgt_mask]
        overlap_packed = None
        if cfg.get('remove_outliers_overlap', False):
            overlap_packed = torch.cat([o.reshape(-1) for o in list(src_overlap_list) + list(tgt_overlap_list)])
        pose, attn_list, val_list, ind_list, src_pts_list, tgt_pts_list = self._match_and_solve(
            src_packed, tgt_packed, pts_c, src_slens_c, tgt_slens_c, overlap_packed)

        if cfg.get('check_numerics', False):
            ops.check_numerics()  # one device synchronisation: fp16 operand range of the tensor-core GEMMs
        return {
            'pose': pose, 'attn': attn_list, 'src_feat': src_cond_list, 'tgt_feat': tgt_cond_list,
            'src_kp': src_xyz_c, 'tgt_kp': tgt_xyz_c, 'src_corr': src_pts_list, 'tgt_corr': tgt_pts_list,
            'src_overlap': src_overlap_list, 'tgt_overlap': tgt_overlap_list,
            'overlap_prob_list': val_list, 'ind_list': ind_list,
        }

    # -------------------------------------------------------------------------------------------
    def softmax_correlation(self, src_feats, tgt_feats, src_xyz, tgt_xyz, src_overlap_list=None, tgt_overlap_list=None):
        """Reference signature (:423): lists of (1,N,D)/(1,M,D) features and (N,3)/(M,3) coordinates."""
        src_lens = [f.shape[-2] for f in src_feats]
        tgt_lens = [f.shape[-2] for f in tgt_feats]
        src_packed = torch.cat([f.reshape(-1, f.shape[-1]) for f in src_feats], dim=0)
        tgt_packed = torch.cat([f.reshape(-1, f.shape[-1]) for f in tgt_feats], dim=0)
        pts = torch.cat(list(src_xyz) + list(tgt_xyz), dim=0)
        overlap_packed = None
        if self.cfg.get('remove_outliers_overlap', False):
            overlap_packed = torch.cat([o.reshape(-1) for o in list(src_overlap_list) + list(tgt_overlap_list)])
        return self._match_and_solve(src_packed, tgt_packed, pts, src_lens, tgt_lens, overlap_packed)

    def _affinity_scalars(self):
        """softplus(alpha), exp(beta) as host floats, cached per parameter version (one sync per weight load)."""
        key = (self.alpha._version, self.beta._version, self.alpha.data_ptr())
        if getattr(self, '_aff_key', None) != key:
            self._aff_val = (float(F.softplus(self.alpha.detach())), float(torch.exp(self.beta.detach())))
            self._aff_key = key
        return self._aff_val

    def _match_and_solve(self, src_packed, tgt_packed, pts_c, src_lens, tgt_lens, overlap_packed=None):
        cfg = self.cfg
        dev = src_packed.device
        pairs = ops.packed_pairs(src_lens, tgt_lens, dev)
        total_src = pairs.total_src
        src_xyz_packed, tgt_xyz_packed = pts_c[:total_src], pts_c[total_src:]
        ratio_test = bool(cfg.get('use_ratio_test', False))
        corr, attn, val, ind = ops.dual_softmax_match(src_packed, tgt_packed, pairs,
                                                      want_attn=self.return_attn or ratio_test)
        P = pairs.P
        attn_list = ([attn[pairs.h_co[p]:pairs.h_co[p + 1]].view(1, src_lens[p], tgt_lens[p]) for p in range(P)]
                     if attn is not None and self.return_attn else [None] * P)
        if ratio_test:  # :465-466 / :573-574
            val, ind = ops.top2_ratio(attn, pairs, float(cfg.lowe_thres))

        offsets, h_off = pairs.oo, pairs.h_oo
        need_rows = (not cfg.use_sinkhorn) or cfg.get('threshold_corr', False) or cfg.get('remove_outliers_overlap', False)
        if need_rows:
            # one row per correspondence: which pair it belongs to, which side the argmax indexes
            # (small host lists travel through pinned staging: torch.tensor(..., device=cuda) would block the host
            # until the device has drained)
            rows_per_pair = ops.to_device_async([h_off[p + 1] - h_off[p] for p in range(P)], torch.int64, dev)
            pair_of_row = torch.repeat_interleave(torch.arange(P, device=dev), rows_per_pair,
                                                  output_size=pairs.total_out)  # (output_size: no device read-back)
            gather_from_src = ops.to_device_async([1 if n > m else 0 for n, m in zip(src_lens, tgt_lens)], torch.int64, dev)
            from_src = gather_from_src[pair_of_row].bool()
            base = torch.where(from_src, pairs.so[:-1][pair_of_row], pairs.to[:-1][pair_of_row] + total_src)
            local = torch.arange(pairs.total_out, device=dev) - pairs.oo[:-1][pair_of_row]
            other_base = torch.where(from_src, pairs.to[:-1][pair_of_row] + total_src, pairs.so[:-1][pair_of_row])
        if cfg.get('threshold_corr', False):  # :471-473 / :579-581: keep values above the pair's (lower) median
            val = torch.where(val > _segment_median(val, h_off, pair_of_row), val, torch.zeros_like(val))
        weights = val
        if cfg.get('remove_outliers_overlap', False):  # :481-492 / :596-607
            if overlap_packed is None:
                raise RuntimeError("remove_outliers_overlap needs the overlap predictions of both clouds")
            ov = overlap_packed.reshape(-1).to(torch.float32)
            overlap_prob = ov[(base + ind).long()] * ov[(other_base + local).long()]
            if cfg.get('use_overlap_as_weights', False):
                weights = overlap_prob  # :546 / :651; val stays the reported match weight
            else:
                val = val * overlap_prob
                weights = val

        if cfg.use_sinkhorn:
            # :532-536 / :641-647 -- all source points against Sinkhorn-weighted targets
            sp_alpha, e_beta = self._affinity_scalars()
            wt, w = ops.sinkhorn_weighted_targets(corr, pairs, tgt_xyz_packed, sp_alpha, e_beta, int(cfg.sinkhorn_itr),
                                                  bool(cfg.slack))
            pose = ops.weighted_procrustes(src_xyz_packed, wt, w, pairs.so)
            src_pts_list = [src_xyz_packed[pairs.h_so[p]:pairs.h_so[p + 1]] for p in range(P)]
            tgt_pts_list = [tgt_xyz_packed[pairs.h_to[p]:pairs.h_to[p + 1]] for p in range(P)]
            if cfg.get('remove_points_from_val', False):
                raise NotImplementedError("remove_points_from_val with use_sinkhorn: the reference gathers both clouds "
                                          "with one index list (:497-500), out of range whenever N != M")
        else:
            # :478,548 (N > M: one source per target) / :589,653 (one target per source)
            picked = ops.gather_rows3(pts_c, ind, base.to(torch.int32))
            # the side that is NOT gathered is simply that pair's full cloud, in order
            other = pts_c[(other_base + local).long()]
            a = torch.where(from_src[:, None], picked, other)  # source-side points
            b = torch.where(from_src[:, None], other, picked)  # target-side points
            if cfg.get('remove_points_from_val', False):  # :497-500 / :612-615: keep the top int(thr * len) rows
                keep, new_off = [], [0]
                for p in range(P):
                    k = int(float(cfg.val_threshold) * (h_off[p + 1] - h_off[p]))
                    top = torch.topk(val[h_off[p]:h_off[p + 1]], k).indices
                    keep.append(top + h_off[p])
                    new_off.append(new_off[-1] + k)
                keep = torch.cat(keep)
                a, b, val = a[keep].contiguous(), b[keep].contiguous(), val[keep].contiguous()
                ind = keep - pairs.oo[:-1][pair_of_row[keep]].long()  # the reference reports the top-k positions
                weights, h_off = val, new_off
                offsets = ops.to_device_async(new_off, torch.int32, dev)
            pose = ops.weighted_procrustes(a, b, weights, offsets)
            if cfg.get('use_lgr', False):  # :553-554 / :658-659, starting from val (not the overlap weights)
                pose = ops.local_global_registration(a, b, val, pose, offsets, float(cfg.acceptance_radius),
                                                     int(cfg.num_refinement_steps))
            if cfg.get('use_ransac', False):  # :556-557 / :661-662: 500 hypotheses from 100 rows drawn with replacement
                counts = ops.to_device_async([float(h_off[p + 1] - h_off[p]) for p in range(P)], torch.float32, dev)
                u = torch.rand((P, self.ransac_hypotheses, self.ransac_sample_size), device=dev,
                               generator=self.ransac_generator)
                idx = torch.minimum((u * counts[:, None, None]).long(), (counts[:, None, None] - 1).long())
                pose, _, _ = ops.ransac(a, b, val, offsets, idx)
            src_pts_list = [a[h_off[p]:h_off[p + 1]] for p in range(P)]
            tgt_pts_list = [b[h_off[p]:h_off[p + 1]] for p in range(P)]
        val_list = [val[h_off[p]:h_off[p + 1]] for p in range(P)]
        ind_list = [ind[h_off[p]:h_off[p + 1]] for p in range(P)]
        return pose, attn_list, val_list, ind_list, src_pts_list, tgt_pts_list


def _segment_median(val, h_off, pair_of_row):
    """torch.median of each pair's values (the lower of the two middle ones), broadcast back to the rows."""
    P = len(h_off) - 1
    lens = [h_off[p + 1] - h_off[p] for p in range(P)]
    pad = torch.full((P, max(max(lens), 1)), float('inf'), dtype=val.dtype, device=val.device)
    local = torch.arange(val.shape[0], device=val.device) - ops.to_device_async(h_off[:-1], torch.int64, val.device)[pair_of_row]
    pad[pair_of_row, local] = val
    srt, _ = torch.sort(pad, dim=1)
    mid = ops.to_device_async([max(l - 1, 0) // 2 for l in lens], torch.int64, val.device)
    return srt[torch.arange(P, device=val.device), mid][pair_of_row]
