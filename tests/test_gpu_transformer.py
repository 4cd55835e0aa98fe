"""Packed cross-encoder path (our attention kernel) against PyTorch fp32/fp64 references."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from superpoints_registration_b200 import config as cfgs
from superpoints_registration_b200 import ops
from superpoints_registration_b200.model import RegTR

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _ref_attention(qkv, tiles_spec, n_heads, d):
    """fp64 soft-max attention per segment: tiles_spec = [(q_off, q_len, kv_off, kv_len)]."""
    out = torch.zeros((qkv.shape[0], d), dtype=torch.float64)
    hd = d // n_heads
    x = qkv.double().cpu()
    for qo, qn, ko, kn in tiles_spec:
        q = x[qo:qo + qn, :d].view(qn, n_heads, hd).transpose(0, 1)
        k = x[ko:ko + kn, d:2 * d].view(kn, n_heads, hd).transpose(0, 1)
        v = x[ko:ko + kn, 2 * d:].view(kn, n_heads, hd).transpose(0, 1)
        p = torch.softmax(q @ k.transpose(1, 2) / math.sqrt(hd), dim=-1)
        out[qo:qo + qn] = (p @ v).transpose(0, 1).reshape(qn, d)
    return out


@pytest.mark.parametrize("generation", [2, 1])
@pytest.mark.parametrize("lens", [[5, 64, 65, 130], [1, 1], [700, 333, 128, 257, 64, 63]])
def test_attention_varlen_against_fp64(lens, generation):
    """generation 2 = tcgen05 kernel (attention_tc.cu, 128-query tiles), 1 = warp-level kernel (attention.cu)."""
    torch.manual_seed(len(lens))
    d, nh = 256, 8
    T = sum(lens)
    qkv = (torch.randn(T, 3 * d) * 2.0).to(DEV)
    offs = np.concatenate([[0], np.cumsum(lens)]).tolist()
    B = len(lens) // 2
    partner = list(range(B, 2 * B)) + list(range(B))
    for kind in ("self", "cross"):
        ko = offs[:-1] if kind == "self" else [offs[p] for p in partner]
        kn = lens if kind == "self" else [lens[p] for p in partner]
        tiles = ops.attention_tiles(offs[:-1], lens, ko, kn, DEV, block_q=128 if generation == 2 else 64)
        hi, lo = ops.split_f16(qkv, n_scaled=d, scale=math.log2(math.e) / math.sqrt(d // nh))
        out = ops.attention_varlen(hi, lo, tiles, nh, 0, d, 2 * d, d, generation=generation).cpu().double()
        ref = _ref_attention(qkv, list(zip(offs[:-1], lens, ko, kn)), nh, d)
        err = (out - ref).abs().max().item()
        print(f"gen {generation} {kind}: max err {err:.2e} = {err / ref.abs().max().item():.2e} x max|out|")
        # logits reach +-16 here: 22-bit operand products leave ~|S| * 2^-22 relative error on P, the same order as fp32
        assert err <= 6e-6 * ref.abs().max().item(), (kind, err, ref.abs().max().item())
        # the operand-image output (what the fused encoder layer consumes) through the output projection's GEMM
        img = ops.gemm_a_image(T, d, DEV)
        ops.attention_varlen(hi, lo, tiles, nh, 0, d, 2 * d, d, out_image=img, image_scale=ops.A_SCALE,
                             generation=generation)
        eye = ops.weight_image(torch.eye(d, device=DEV))
        back = ops.gemm_tc(img, eye, None, T).cpu().double()
        assert (back - ref).abs().max().item() <= 6e-6 * ref.abs().max().item()
        again = ops.attention_varlen(hi, lo, tiles, nh, 0, d, 2 * d, d, generation=generation).cpu().double()
        assert torch.equal(out, again)   # deterministic


def test_split_f16_reconstructs_fp32():
    x = (torch.randn(300, 768, device=DEV) * 3)
    hi, lo = ops.split_f16(x, n_scaled=256, scale=0.25)
    want = x.clone()
    want[:, :256] *= 0.25
    err = (hi.float() + lo.float() - want).abs().max().item()
    assert err <= 4e-7 * want.abs().max().item()


def test_packed_encoder_matches_padded_pytorch_modules():
    torch.manual_seed(0)
    cfg = cfgs.threedmatch_config()
    model = RegTR(cfg).to(DEV).eval()
    enc = model.transformer_encoder
    lens = [150, 97, 64, 201]     # 2 pairs
    T = sum(lens)
    x = torch.randn(T, cfg.d_embed, device=DEV)
    pos = torch.randn(T, cfg.d_embed, device=DEV) * 0.5
    with torch.no_grad():
        got = enc.forward_packed(x, pos, lens)
        # padded reference through the nn.MultiheadAttention modules
        parts, pparts = torch.split(x, lens), torch.split(pos, lens)
        pad = lambda seqs: torch.nn.utils.rnn.pad_sequence(list(seqs))
        mask = lambda ls: (torch.arange(max(ls), device=DEV)[None, :] >= torch.tensor(ls, device=DEV)[:, None])
        s, t = enc(pad(parts[:2]), pad(parts[2:]), src_key_padding_mask=mask(lens[:2]), tgt_key_padding_mask=mask(lens[2:]),
                   src_pos=pad(pparts[:2]), tgt_pos=pad(pparts[2:]))
    ref = torch.cat([s[0, :lens[0], 0], s[0, :lens[1], 1], t[0, :lens[2], 0], t[0, :lens[3], 1]])
    err = (got - ref).abs().max().item()
    assert err <= 2e-5 * ref.abs().max().item(), (err, ref.abs().max().item())


@pytest.mark.parametrize("T,N,K", [(64, 128, 64), (200, 256, 256), (1000, 768, 256), (333, 256, 1024), (70, 32, 96)])
def test_gemm_tc_against_fp64(T, N, K):
    torch.manual_seed(T + N + K)
    x = torch.randn(T, K, device=DEV) * 1.5
    w = torch.randn(N, K, device=DEV) / math.sqrt(K)
    b = torch.randn(N, device=DEV)
    res = torch.randn(T, N, device=DEV)
    ref = (x.double() @ w.double().t() + b.double())
    y = ops.linear_tc(x, w, b)
    assert (y.double() - ref).abs().max().item() <= 3e-6 * ref.abs().max().item()
    y2 = ops.linear_tc(x, w, b, residual=res)
    assert (y2.double() - (ref + res.double())).abs().max().item() <= 3e-6 * ref.abs().max().item()
    y3 = ops.linear_tc(x, w, None, relu=True)
    assert (y3.double() - (x.double() @ w.double().t()).clamp(min=0)).abs().max().item() <= 3e-6 * ref.abs().max().item()


def test_gemm_tc_plane_and_image_outputs_chain():
    """QKV-style plane output and FFN-style image chaining (GEMM -> ReLU -> image -> GEMM)."""
    torch.manual_seed(5)
    T, d, dff = 300, 256, 1024
    x = torch.randn(T, d, device=DEV)
    w1 = torch.randn(dff, d, device=DEV) / math.sqrt(d)
    b1 = torch.randn(dff, device=DEV) * 0.1
    w2 = torch.randn(d, dff, device=DEV) / math.sqrt(dff)
    b2 = torch.randn(d, device=DEV) * 0.1
    img = ops.gemm_prepare_input(x)
    hi, lo = ops.gemm_tc(img, ops.weight_image(w1), b1, T, ops.OUT_PLANES, n_scaled=256, col_scale=0.5)
    want = x.double() @ w1.double().t() + b1.double()
    want[:, :256] *= 0.5
    got = hi.double() + lo.double()
    assert (got - want).abs().max().item() <= 3e-6 * want.abs().max().item()
    h_img = ops.gemm_tc(img, ops.weight_image(w1), b1, T, ops.OUT_AIMG, relu=True)
    y = ops.gemm_tc(h_img, ops.weight_image(w2), b2, T, ops.OUT_F32, residual=x)
    ref = (x.double() @ w1.double().t() + b1.double()).clamp(min=0) @ w2.double().t() + b2.double() + x.double()
    assert (y.double() - ref).abs().max().item() <= 3e-6 * ref.abs().max().item()


def test_layernorm_prepare_matches_torch():
    torch.manual_seed(6)
    T = 130
    x = torch.randn(T, 256, device=DEV) * 2 + 0.3
    g, b = torch.randn(256, device=DEV), torch.randn(256, device=DEV)
    pos = torch.randn(T, 256, device=DEV)
    out = ops.layernorm256_prepare(x, g, b, pos, 1e-5, None, out_f32=True)
    ref = F.layer_norm(x, (256,), g, b, 1e-5) + pos
    assert (out - ref).abs().max().item() <= 2e-6 * ref.abs().max().item()
    # the image feeds a GEMM that reproduces LN(x) @ I
    img = ops.gemm_a_image(T, 256, DEV)
    ops.layernorm256_prepare(x, g, b, pos, 1e-5, img)
    eye = torch.eye(256, device=DEV)
    y = ops.gemm_tc(img, ops.weight_image(eye), None, T, ops.OUT_F32)
    assert (y - ref).abs().max().item() <= 2e-6 * ref.abs().max().item()


def test_native_sequencer_is_bit_identical_to_the_python_sequence():
    """spr_cross_encoder_forward (all layers from one library call) issues the launches of forward_fused in the same
    order: the conditioned features must be bit-identical, with and without positions, and follow a weight update."""
    torch.manual_seed(3)
    cfg = cfgs.threedmatch_config()
    model = RegTR(cfg).to(DEV).eval()
    enc = model.transformer_encoder
    lens = [300, 17, 129, 64, 250, 31, 128, 90]
    T = sum(lens)
    x = torch.randn(T, 256, device=DEV)
    pos = torch.randn(T, 256, device=DEV) * 0.5
    for p in (pos, None):
        enc.native_sequencer, enc.cuda_graphs = True, False
        a = enc.forward_packed(x, p, lens)
        enc.native_sequencer = False
        b = enc.forward_packed(x, p, lens)
        assert torch.equal(a, b)
        # ... and replayed as a CUDA graph of its shape bucket (rows padded to 64, tile lists to 8 entries): first call
        # captures, second replays; another set of lengths in the same bucket reuses the graph
        enc.native_sequencer, enc.cuda_graphs = True, True
        g1 = enc.forward_packed(x, p, lens)
        g2 = enc.forward_packed(x, p, lens)
        assert torch.equal(g1, b) and torch.equal(g2, b)
        lens2 = [lens[0] - 5, lens[1] + 3] + lens[2:]          # same bucket (T - 2 rows), different segments
        x2 = x[:sum(lens2)].contiguous()
        p2 = None if p is None else p[:sum(lens2)].contiguous()
        g3 = enc.forward_packed(x2, p2, lens2)
        enc.native_sequencer = False
        assert torch.equal(g3, enc.forward_packed(x2, p2, lens2))
    enc.cuda_graphs = False
    with torch.no_grad():
        enc.layers[2].linear1.weight.mul_(1.5)          # version bump: the pointer table and the image must follow
    enc.native_sequencer = True
    a2 = enc.forward_packed(x, pos, lens)
    enc.native_sequencer = False
    b2 = enc.forward_packed(x, pos, lens)
    assert torch.equal(a2, b2) and not torch.equal(a2, a)


def test_fused_encoder_matches_packed_torch_linears():
    torch.manual_seed(1)
    cfg = cfgs.threedmatch_config()
    model = RegTR(cfg).to(DEV).eval()
    enc = model.transformer_encoder
    lens = [150, 97, 64, 201, 130, 33]
    T = sum(lens)
    x = torch.randn(T, cfg.d_embed, device=DEV)
    pos = torch.randn(T, cfg.d_embed, device=DEV) * 0.5
    with torch.no_grad():
        enc.fused = True
        got = enc.forward_packed(x, pos, lens)
        enc.fused = False
        ref = enc.forward_packed(x, pos, lens)
    err = (got - ref).abs().max().item()
    assert err <= 2e-5 * ref.abs().max().item(), (err, ref.abs().max().item())


@pytest.mark.parametrize("kind", ["kitti", "modelnet", "3dmatch", "3dmatch_4stage"])
def test_forward_fused_paths_agree_with_plain_paths(kind):
    """Whole forward on every shipped configuration: the fused route (format-aware encoder blocks, tensor-core KPConv,
    packed cross-encoder on our GEMM / attention kernels) against the plain route (fp32 SIMT KPConv, unfused blocks,
    padded nn.MultiheadAttention modules) with the same weights."""
    from superpoints_registration_b200 import synthetic
    cfg = {"kitti": cfgs.kitti_config, "modelnet": cfgs.modelnet_config, "3dmatch": cfgs.threedmatch_config,
           "3dmatch_4stage": cfgs.threedmatch_4stage_config}[kind]()
    kind = kind.split("_")[0]
    torch.manual_seed(3)
    np.random.seed(3)
    model = RegTR(cfg).to(DEV).eval()
    model.return_attn = False
    kw = {"n_points": 6000} if kind != "modelnet" else {}
    data = synthetic.make_batch(kind, 2, seed=7, **kw)
    batch = {"src_xyz": [torch.from_numpy(c).to(DEV) for c in data["src_xyz"]],
             "tgt_xyz": [torch.from_numpy(c).to(DEV) for c in data["tgt_xyz"]]}
    with torch.no_grad():
        fused = model(dict(batch))
        # plain route
        model.packed_transformer = False
        for m in model.modules():
            if m.__class__.__name__ == "KPConv":
                m.mode = 0
        plain = model(dict(batch))
        # third route for diagnosis: packed tokens, our attention, torch linears, fp32 SIMT KPConv
        model.packed_transformer = True
        model.transformer_encoder.fused = False
        third = model(dict(batch))
    for a, b, c in zip(fused["src_feat"] + fused["tgt_feat"], plain["src_feat"] + plain["tgt_feat"],
                       third["src_feat"] + third["tgt_feat"]):
        print(f"[{kind}] fused-plain {(a - b).abs().max().item():.2e} fused-third {(a - c).abs().max().item():.2e} "
              f"plain-third {(b - c).abs().max().item():.2e}")
    for a, b in zip(fused["src_feat"] + fused["tgt_feat"], plain["src_feat"] + plain["tgt_feat"]):
        assert torch.isfinite(a).all()
        err = (a - b).abs().max().item()
        assert err <= 3e-4 * b.abs().max().item(), (kind, err, b.abs().max().item())
    R = fused["pose"][:, :, :3].double()
    assert torch.allclose(R @ R.transpose(1, 2), torch.eye(3, device=DEV, dtype=torch.float64).expand_as(R), atol=1e-5)


def test_weight_image_cache_follows_the_tensor():
    """The operand image of a weight is rebuilt when the weight changes in place and never leaks to another tensor."""
    x = torch.randn(70, 64, device=DEV)
    w = torch.nn.Parameter(torch.randn(32, 64, device=DEV))
    y1 = ops.linear_tc(x, w)
    with torch.no_grad():
        w.mul_(2.0)
    y2 = ops.linear_tc(x, w)
    assert (y2 - 2 * y1).abs().max().item() <= 1e-5 * y2.abs().max().item()
    for _ in range(20):   # fresh parameters may reuse the freed one's id / address / version
        w2 = torch.nn.Parameter(torch.randn(32, 64, device=DEV))
        ref = x.double() @ w2.double().t()
        assert (ops.linear_tc(x, w2).double() - ref).abs().max().item() <= 3e-6 * ref.abs().max().item()
        del w2


@pytest.mark.parametrize("mag", [1e-4, 1e-2, 1.0, 60.0, 3000.0])
def test_gemm_tc_accuracy_over_input_magnitudes(mag):
    """The GEMM's activations travel as fp16 (hi, lo) pairs at a fixed power-of-two scale (ops.A_SCALE): fp32-level
    accuracy must hold from 1e-4 up to the documented limit of ~4094, relative to the output scale."""
    from superpoints_registration_b200 import _lib
    g = torch.Generator(device=DEV).manual_seed(int(mag * 1e4) % 9973)
    x = (torch.rand(300, 256, device=DEV, generator=g) * 2 - 1) * mag      # bounded: |x| <= mag < 4094
    w = torch.randn(192, 256, device=DEV, generator=g) / 16
    b = torch.randn(192, device=DEV, generator=g)
    _lib.numeric_flags(reset=True)
    y = ops.linear_tc(x, w, b)
    ref = x.double() @ w.double().t() + b.double()
    assert (y.double() - ref).abs().max().item() <= 3e-6 * ref.abs().max().item()
    assert _lib.numeric_flags() == 0


def test_gemm_tc_operand_overflow_is_loud():
    """Beyond the fp16 operand range the result is NaN (not a wrong finite number) and the sticky flag is raised."""
    from superpoints_registration_b200 import _lib
    x = torch.randn(128, 64, device=DEV)
    x[5, 7] = 1.0e4                                   # 1e4 * 16 > 65504
    w = torch.randn(32, 64, device=DEV) / 8
    _lib.numeric_flags(reset=True)
    y = ops.linear_tc(x, w)
    assert torch.isnan(y[5]).all() and torch.isfinite(y[:5]).all() and torch.isfinite(y[6:]).all()
    with pytest.raises(FloatingPointError):
        ops.check_numerics()
    ops.check_numerics()                              # the check cleared the flag
    # every other producer of operand images raises it as well
    lens = torch.tensor([128], dtype=torch.int32, device=DEV)
    big = torch.randn(128, 64, device=DEV)
    res = torch.zeros_like(big)
    res[3, 3] = 1.0e4
    ops.instance_norm_lrelu_ex(big, lens, residual=res, want_f32=False, want_image=True)
    assert _lib.numeric_flags() & _lib.FLAG_FP16_OVERFLOW
    t = torch.randn(64, 256, device=DEV)
    gamma = torch.full((256,), 1.0e4, device=DEV)                    # LayerNorm output x 1e4
    ops.layernorm256_prepare(t, gamma, torch.zeros(256, device=DEV), None, 1e-5, ops.gemm_a_image(64, 256, DEV))
    assert _lib.numeric_flags() & _lib.FLAG_FP16_OVERFLOW
    assert _lib.numeric_flags() == 0
