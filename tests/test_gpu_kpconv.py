"""CUDA KPConv layer, block epilogues and the encoder against the oracle and the reference's golden vectors."""
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import numpy_ops
from parity import load_pyramid
from superpoints_registration_b200 import config as cfgs
from superpoints_registration_b200 import ops
from superpoints_registration_b200.kpconv import KPFEncoder
from superpoints_registration_b200.kpconv_blocks import KPConv, max_pool
from weights import filled_state, reference_shapes

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

# Tolerance for fp32 features, relative to the largest output magnitude of the layer.  The reference's own
# torch GEMMs sit ~5e-6 away from the exactly-rounded value (SURVEY.md hard part 5); 2e-5 leaves room for
# two independent fp32 summation orders.
FEAT_RTOL = 2e-5


def _t(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


@pytest.mark.parametrize("mode", [0, 1])
def test_kpconv_layers_against_golden_and_oracle(golden_dir, mode):
    g = np.load(os.path.join(golden_dir, "kpconv_layers.npz"))
    for tag in map(str, g["tags"]):
        q, s, idx, x, kp = g[f"{tag}_q"], g[f"{tag}_s"], g[f"{tag}_idx"], g[f"{tag}_x"], g[f"{tag}_kp"]
        cin, cout = x.shape[1], g[f"{tag}_out"].shape[1]
        wname = f"{tag}.KPConv.weights"
        w = filled_state({wname: (15, cin, cout)}, 100 + cin)[wname]
        ext = float(g[f"{tag}_extent"])
        for idx_dtype in (torch.int64, torch.int32):
            m = mode if cin == cout else 0
            out = ops.kpconv_forward(_t(q), _t(s), _t(idx, idx_dtype), _t(x), _t(w), _t(kp), ext, mode=m).cpu().numpy()
            ref = g[f"{tag}_out"]                                   # the reference module's output
            exact = oracle.kpconv_forward(q, s, idx.astype(np.int64), x, w, kp, ext)  # fp64-accumulated oracle
            scale = np.abs(ref).max()
            assert np.abs(out - exact).max() <= FEAT_RTOL * scale, (tag, np.abs(out - exact).max(), scale)
            assert np.abs(out - ref).max() <= FEAT_RTOL * scale, (tag, np.abs(out - ref).max(), scale)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("c", [32, 64, 128, 256])
def test_kpconv_random_shapes_against_oracle(c, mode):
    rng = np.random.default_rng(c)
    ns, nq, H = 700, 333, 23                       # nq not a multiple of the tile, odd H
    s = rng.uniform(0, 1, size=(ns, 3)).astype(np.float32)
    q = s[rng.permutation(ns)[:nq]] + rng.normal(0, 0.01, size=(nq, 3)).astype(np.float32)
    idx = rng.integers(0, ns + 1, size=(nq, H))    # includes shadow entries (== ns)
    idx[5] = ns                                    # a row with no neighbour at all
    x = np.where(rng.uniform(size=(ns, c)) < 0.5, rng.normal(size=(ns, c)), -0.1 * rng.uniform(size=(ns, c))).astype(np.float32)
    w = (rng.normal(size=(15, c, c)) / np.sqrt(15 * c)).astype(np.float32)
    kp = (rng.normal(size=(15, 3)) * 0.15).astype(np.float32)
    kp[0] = 0
    out = ops.kpconv_forward(_t(q), _t(s), _t(idx), _t(x), _t(w), _t(kp), 0.3, mode=mode).cpu().numpy()
    exact = oracle.kpconv_forward(q, s, idx, x, w, kp, 0.3)
    err = np.abs(out - exact).max()
    print(f"C={c} mode={mode}: max err {err:.3e} = {err / np.abs(exact).max():.2e} x max|out|")
    assert err <= FEAT_RTOL * np.abs(exact).max()
    assert np.all(out[5] == 0)


@pytest.mark.parametrize("scale_x,scale_w", [(1e-6, 1.0), (3e4, 1e-3), (1.0, 250.0)])
def test_kpconv_tensor_core_operand_scaling(scale_x, scale_w):
    """The tensor-core path splits fp32 operands into fp16 pairs after a power-of-two rescale derived from max|x| and
    max|W|: accuracy must not depend on the magnitude of the inputs."""
    rng = np.random.default_rng(11)
    ns, nq, H, c = 900, 500, 30, 64
    s = rng.uniform(0, 1, size=(ns, 3)).astype(np.float32)
    q = s[:nq]
    idx = rng.integers(0, ns + 1, size=(nq, H))
    x = (rng.normal(size=(ns, c)) * scale_x).astype(np.float32)
    w = (rng.normal(size=(15, c, c)) / np.sqrt(15 * c) * scale_w).astype(np.float32)
    kp = (rng.normal(size=(15, 3)) * 0.15).astype(np.float32)
    out = ops.kpconv_forward(_t(q), _t(s), _t(idx), _t(x), _t(w), _t(kp), 0.3, mode=1).cpu().numpy()
    exact = oracle.kpconv_forward(q, s, idx, x, w, kp, 0.3)
    assert np.abs(out - exact).max() <= FEAT_RTOL * np.abs(exact).max()


def test_kpconv_tensor_core_many_tiles_and_zero_input():
    """More tiles than SMs (persistent loop, ring wrap-around) and an all-zero feature matrix (scale exponent guard)."""
    rng = np.random.default_rng(12)
    ns, H, c = 30000, 24, 32
    nq = ns
    s = rng.uniform(0, 3, size=(ns, 3)).astype(np.float32)
    idx = rng.integers(0, ns + 1, size=(nq, H))
    x = np.maximum(rng.normal(size=(ns, c)), 0).astype(np.float32)
    w = (rng.normal(size=(15, c, c)) / np.sqrt(15 * c)).astype(np.float32)
    kp = (rng.normal(size=(15, 3)) * 0.15).astype(np.float32)
    out = ops.kpconv_forward(_t(s), _t(s), _t(idx), _t(x), _t(w), _t(kp), 0.3, mode=1)
    ref = ops.kpconv_forward(_t(s), _t(s), _t(idx), _t(x), _t(w), _t(kp), 0.3, mode=0)
    assert (out - ref).abs().max().item() <= FEAT_RTOL * ref.abs().max().item()
    z = ops.kpconv_forward(_t(s), _t(s), _t(idx), _t(np.zeros_like(x)), _t(w), _t(kp), 0.3, mode=1)
    assert torch.count_nonzero(z).item() == 0


def test_kpconv_strided_view_indices():
    """Index matrices trimmed by the Preprocessor are non-contiguous views (row stride = limit)."""
    rng = np.random.default_rng(3)
    ns, nq = 400, 200
    s = rng.uniform(0, 1, size=(ns, 3)).astype(np.float32)
    q = s[:nq]
    full = rng.integers(0, ns + 1, size=(nq, 40))
    x = rng.normal(size=(ns, 32)).astype(np.float32)
    w = (rng.normal(size=(15, 32, 32)) * 0.05).astype(np.float32)
    kp = (rng.normal(size=(15, 3)) * 0.15).astype(np.float32)
    view = _t(full)[:, :17]
    assert not view.is_contiguous()
    out = ops.kpconv_forward(_t(q), _t(s), view, _t(x), _t(w), _t(kp), 0.3, mode=1).cpu().numpy()
    out0 = ops.kpconv_forward(_t(q), _t(s), view, _t(x), _t(w), _t(kp), 0.3, mode=0).cpu().numpy()
    exact = oracle.kpconv_forward(q, s, full[:, :17].copy(), x, w, kp, 0.3)
    assert np.abs(out - exact).max() <= FEAT_RTOL * np.abs(exact).max()
    assert np.abs(out0 - exact).max() <= FEAT_RTOL * np.abs(exact).max()


def test_kpconv_module_surface_and_errors():
    conv = KPConv(15, 3, 32, 32, 0.05, 0.0625).to(DEV)
    assert tuple(conv.weights.shape) == (15, 32, 32) and tuple(conv.kernel_points.shape) == (15, 3)
    assert not conv.kernel_points.requires_grad
    assert set(conv.state_dict()) == {"weights", "kernel_points"}
    with pytest.raises(NotImplementedError):
        KPConv(15, 3, 32, 32, 0.05, 0.0625, deformable=True)
    with pytest.raises(NotImplementedError):
        KPConv(15, 3, 32, 32, 0.05, 0.0625, KP_influence="gaussian")
    pts = torch.rand(50, 3, device=DEV)
    with pytest.raises(RuntimeError):   # unsupported channel shape fails loudly, no fallback
        ops.kpconv_forward(pts, pts, torch.zeros((50, 4), dtype=torch.int64, device=DEV), torch.rand(50, 48, device=DEV),
                           torch.rand(15, 48, 48, device=DEV), torch.rand(15, 3, device=DEV), 0.1)
    with pytest.raises(RuntimeError):   # CPU tensors are rejected
        ops.kpconv_forward(pts.cpu(), pts.cpu(), torch.zeros((50, 4), dtype=torch.int64), torch.rand(50, 32),
                           torch.rand(15, 32, 32), torch.rand(15, 3), 0.1)


def test_instance_norm_and_max_pool_against_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "blocks.npz"))
    y = ops.instance_norm_lrelu(_t(g["x"]), _t(g["lens"])).cpu().numpy()
    d = np.abs(y - g["inorm"])
    o = 0
    for n in g["lens"].tolist():
        assert d[o:o + n].max() < (1e-6 if n >= 32 else 2e-4)
        o += n
    res = np.random.default_rng(0).normal(size=g["x"].shape).astype(np.float32)
    y2 = ops.instance_norm_lrelu(_t(g["x"]), _t(g["lens"]), slope=0.1, residual=_t(res)).cpu().numpy()
    want = numpy_ops.instance_norm_lrelu(g["x"], g["lens"], slope=0.1, residual=res)
    assert np.abs(y2 - want).max() < 2e-4
    mp = max_pool(_t(g["x"]), _t(g["pool_idx"], torch.int64)).cpu().numpy()
    assert np.array_equal(mp, g["pool_out"])
    # a processing order (the pyramid's cell order in the encoder) changes which CTA pools which point, not the result
    perm = torch.randperm(g["pool_idx"].shape[0], device=DEV).to(torch.int32)
    mp2 = max_pool(_t(g["x"]), _t(g["pool_idx"], torch.int32), perm).cpu().numpy()
    assert np.array_equal(mp2, g["pool_out"])
    with pytest.raises(RuntimeError):
        max_pool(_t(g["x"]), _t(g["pool_idx"], torch.int32), perm[:-1])


def test_instance_norm_is_deterministic_and_handles_many_clouds():
    rng = np.random.default_rng(1)
    lens = rng.integers(2, 700, size=37).astype(np.int32)
    x = rng.normal(size=(int(lens.sum()), 128)).astype(np.float32) * 3 + 1
    a = ops.instance_norm_lrelu(_t(x), _t(lens), slope=0.1)
    b = ops.instance_norm_lrelu(_t(x), _t(lens), slope=0.1)
    assert torch.equal(a, b)
    want = numpy_ops.instance_norm_lrelu(x, lens, slope=0.1)
    assert np.abs(a.cpu().numpy() - want).max() < 5e-5


def test_encoder_against_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "forward_3dmatch.npz"))
    cfg = cfgs.threedmatch_config()
    enc = KPFEncoder(cfg, cfg.d_embed).to(DEV)
    sd = enc.state_dict()
    # the filler draws one stream over the FULL reference key set: regenerate with that key list
    from superpoints_registration_b200.model import RegTR
    own = {k: tuple(v.shape) for k, v in RegTR(cfg).state_dict().items()}
    vals = filled_state(reference_shapes(own, cfg.d_embed), int(g["weight_seed"]))
    new = {}
    for k in sd:
        fk = f"kpf_encoder.{k}"
        new[k] = torch.from_numpy(g[f"kp::{fk}"]) if k.endswith("kernel_points") else torch.from_numpy(vals[fk])
    enc.load_state_dict(new)
    meta_np = load_pyramid(g, "meta_")
    meta = {k: [_t(a, torch.int64 if k in ("neighbors", "pools", "upsamples") else None) for a in v]
            for k, v in meta_np.items()}
    feats0 = torch.ones((meta["points"][0].shape[0], 1), device=DEV)
    out, _ = enc(feats0, meta)
    ref = g["encoder_out"]
    err = np.abs(out.cpu().numpy() - ref).max()
    assert err <= 1e-4 * np.abs(ref).max(), (err, np.abs(ref).max())   # 8 stacked blocks of fp32 rounding


def test_format_aware_instance_norm_outputs_feed_gemm_and_kpconv():
    """instance_norm_lrelu_ex writes the operand image / pre-split rows its consumers read: results must equal the
    plain fp32 route (norm -> linear_tc, norm -> kpconv_forward)."""
    rng = np.random.default_rng(21)
    lens = np.array([300, 157, 43], dtype=np.int32)
    n, c = int(lens.sum()), 64
    x = rng.normal(size=(n, c)).astype(np.float32) * 2 + 0.3
    res = rng.normal(size=(n, c)).astype(np.float32)
    pts = rng.uniform(0, 1, size=(n, 3)).astype(np.float32)
    plain = ops.instance_norm_lrelu(_t(x), _t(lens), slope=0.1, residual=_t(res))
    o = ops.instance_norm_lrelu_ex(_t(x), _t(lens), slope=0.1, residual=_t(res), want_f32=True, want_image=True,
                                   kpconv_points=_t(pts))
    assert (o["f32"] - plain).abs().max().item() <= 1e-6 * plain.abs().max().item()   # FMA contraction may differ
    plain = o["f32"]
    w = torch.randn(96, c, device=DEV) / 8
    y_img = ops.gemm_tc(o["image"], ops.weight_image(w), None, n)
    y_ref = ops.linear_tc(plain, w)
    assert (y_img - y_ref).abs().max().item() <= 2e-6 * y_ref.abs().max().item()
    H = 20
    idx = rng.integers(0, n + 1, size=(n, H))
    wk = (rng.normal(size=(15, c, c)) / np.sqrt(15 * c)).astype(np.float32)
    kp = (rng.normal(size=(15, 3)) * 0.15).astype(np.float32)
    a = ops.kpconv_forward_prepared(_t(pts), _t(idx), o["kpconv"], _t(wk), _t(kp), 0.3)
    b = ops.kpconv_forward(_t(pts), _t(pts), _t(idx), plain, _t(wk), _t(kp), 0.3, mode=1)
    assert (a - b).abs().max().item() <= 2e-6 * b.abs().max().item()
    # c = 32: the image's K is padded to one 64-wide atom
    x32 = rng.normal(size=(n, 32)).astype(np.float32)
    o32 = ops.instance_norm_lrelu_ex(_t(x32), _t(lens), slope=0.1, want_f32=True, want_image=True)
    w32 = torch.randn(128, 32, device=DEV) / 6
    assert (ops.gemm_tc(o32["image"], ops.weight_image(w32), None, n) - ops.linear_tc(o32["f32"], w32)).abs().max().item() \
        <= 1e-6 * 10


@pytest.mark.parametrize("n_out,lens", [(64, [300, 157, 43]), (96, [16, 0, 5, 33, 1, 160, 7]), (256, [1000])])
def test_producer_block_statistics_feed_instance_norm(n_out, lens):
    """spr_gemm_tc's 16-row block sums (stats16) replace the statistics pass of the instance normalisation: same
    result as the two-pass route, for clouds that start and end anywhere relative to the 16-row blocks."""
    rng = np.random.default_rng(n_out)
    lens = np.array(lens, dtype=np.int32)
    n, k = int(lens.sum()), 64
    x = _t(rng.normal(size=(n, k)).astype(np.float32))
    w = torch.randn(n_out, k, device=DEV) / 8 + 0.05              # non-zero column means
    img = ops.gemm_prepare_input(x)
    stats = ops.block_stats(n, n_out, DEV)
    stats.fill_(float("nan"))
    y = ops.gemm_tc(img, ops.weight_image(w), None, n, ops.OUT_F32, stats16=stats)
    assert torch.equal(y, ops.gemm_tc(img, ops.weight_image(w), None, n, ops.OUT_F32))
    full = n // 16
    blocks = y[:full * 16].double().view(full, 16, n_out)
    assert (stats[:full, :, 0].double() - blocks.sum(1)).abs().max().item() <= 1e-5 * blocks.abs().sum(1).max().item()
    assert (stats[:full, :, 1].double() - (blocks ** 2).sum(1)).abs().max().item() <= 1e-5 * (blocks ** 2).sum(1).max().item()
    res = _t(rng.normal(size=(n, n_out)).astype(np.float32))
    a = ops.instance_norm_lrelu_ex(y, _t(lens), slope=0.1, residual=res)["f32"]
    b = ops.instance_norm_lrelu_ex(y, _t(lens), slope=0.1, residual=res, stats16=stats)["f32"]
    assert (a - b).abs().max().item() <= 2e-6 * a.abs().max().item()
    with pytest.raises(RuntimeError):
        ops.instance_norm_lrelu_ex(y, _t(lens), stats16=stats[:-1])
    # a RAW residual (another GEMM's output + its block sums) normalised inside the same apply kernel: bit-identical to
    # normalising it in a pass of its own (the projected shortcut of a bottleneck block, kpconv_blocks.py:706-741)
    w2 = torch.randn(n_out, k, device=DEV) / 8 - 0.02
    stats2 = ops.block_stats(n, n_out, DEV)
    raw = ops.gemm_tc(img, ops.weight_image(w2), None, n, ops.OUT_F32, stats16=stats2)
    two_pass = ops.instance_norm_lrelu_ex(raw, _t(lens), slope=1.0, stats16=stats2)["f32"]
    want = ops.instance_norm_lrelu_ex(y, _t(lens), slope=0.1, residual=two_pass, stats16=stats, want_image=True)
    got = ops.instance_norm_lrelu_ex(y, _t(lens), slope=0.1, residual=raw, stats16=stats, want_image=True,
                                     residual_stats16=stats2)
    assert torch.equal(got["f32"], want["f32"])
    w3 = ops.weight_image(torch.randn(32, n_out, device=DEV) / 8)
    assert torch.equal(ops.gemm_tc(got["image"], w3, None, n, ops.OUT_F32), ops.gemm_tc(want["image"], w3, None, n, ops.OUT_F32))
    with pytest.raises(RuntimeError):
        ops.instance_norm_lrelu_ex(y, _t(lens), residual_stats16=stats2)


@pytest.mark.parametrize("c", [32, 128])
def test_cell_order_is_a_permutation_and_does_not_change_kpconv(c):
    """CellGrid.order() lists every support once, cloud by cloud, in cell order; used as the KPConv processing order
    it must leave every output row bit-identical (the per-query arithmetic does not depend on the tile it runs in)."""
    rng = np.random.default_rng(5 + c)
    lens = np.array([700, 0, 311, 90], dtype=np.int32)
    n = int(lens.sum())
    pts = rng.uniform(0, 1, size=(n, 3)).astype(np.float32)
    grid = ops.CellGrid(_t(pts), _t(lens), 0.12)
    order = grid.order()
    o = order.cpu().numpy()
    assert np.array_equal(np.sort(o), np.arange(n))
    starts = np.concatenate([[0], np.cumsum(lens)])
    for b in range(len(lens)):                                   # clouds stay contiguous
        seg = o[starts[b]:starts[b + 1]]
        assert seg.size == 0 or (seg.min() >= starts[b] and seg.max() < starts[b + 1])
    big = slice(starts[0], starts[1])                            # spatially coherent: consecutive points are close
    step = np.linalg.norm(np.diff(pts[o[big]], axis=0), axis=1).mean()
    assert step < 0.5 * np.linalg.norm(np.diff(pts[big], axis=0), axis=1).mean()
    idx, _ = grid.query(_t(pts), _t(lens), 24)
    x = rng.normal(size=(n, c)).astype(np.float32)
    prep = ops.instance_norm_lrelu_ex(_t(x), _t(lens), slope=0.1, kpconv_points=_t(pts))["kpconv"]
    wk = (rng.normal(size=(15, c, c)) / np.sqrt(15 * c)).astype(np.float32)
    kp = (rng.normal(size=(15, 3)) * 0.05).astype(np.float32)
    a = ops.kpconv_forward_prepared(_t(pts), idx, prep, _t(wk), _t(kp), 0.1)
    b = ops.kpconv_forward_prepared(_t(pts), idx, prep, _t(wk), _t(kp), 0.1, order=order)
    assert torch.equal(a, b)
    with pytest.raises(RuntimeError):
        ops.kpconv_forward_prepared(_t(pts), idx, prep, _t(wk), _t(kp), 0.1, order=order[:-1])


@pytest.mark.parametrize("cout", [32, 64, 128, 256])
def test_kpconv_stem_cin1_against_oracle(cout):
    """First encoder block (Cin = 1): every supported output width, more than 64 neighbour columns, shadow entries."""
    rng = np.random.default_rng(cout)
    ns, nq, H = 600, 257, 70
    s = rng.uniform(0, 1, size=(ns, 3)).astype(np.float32)
    q = s[:nq] + rng.normal(0, 0.01, size=(nq, 3)).astype(np.float32)
    idx = rng.integers(0, ns + 1, size=(nq, H))
    idx[3] = ns
    x = np.where(rng.uniform(size=(ns, 1)) < 0.7, 1.0, -0.25).astype(np.float32)
    w = (rng.normal(size=(15, 1, cout)) / 4).astype(np.float32)
    kp = (rng.normal(size=(15, 3)) * 0.15).astype(np.float32)
    out = ops.kpconv_forward(_t(q), _t(s), _t(idx), _t(x), _t(w), _t(kp), 0.3).cpu().numpy()
    exact = oracle.kpconv_forward(q, s, idx, x, w, kp, 0.3)
    assert np.abs(out - exact).max() <= FEAT_RTOL * np.abs(exact).max()
    assert np.all(out[3] == 0)
    # the thread-per-query kernel reads 32-bit index rows 16 bytes at a time: whole rows, a trimmed view whose width is
    # not a multiple of four (row stride 70, 17 and 40 columns), and a view that starts off the 16-byte grid
    out32 = ops.kpconv_forward(_t(q), _t(s), _t(idx, torch.int32), _t(x), _t(w), _t(kp), 0.3)
    assert torch.equal(out32, _t(out))
    # a processing order (the encoder passes the pyramid's cell order) permutes which thread serves which query only
    perm = torch.randperm(nq, device=DEV).to(torch.int32)
    assert torch.equal(ops.kpconv_forward(_t(q), _t(s), _t(idx, torch.int32), _t(x), _t(w), _t(kp), 0.3, order=perm), out32)
    idx72 = np.concatenate([idx, np.full((nq, 2), ns, dtype=idx.dtype)], axis=1)   # row stride 72: 16-byte rows
    i32 = _t(idx72, torch.int32)
    for lo, hi in ((0, 72), (0, 17), (0, 40), (4, 36), (1, 34)):
        view = i32[:, lo:hi]
        got = ops.kpconv_forward(_t(q), _t(s), view, _t(x), _t(w), _t(kp), 0.3).cpu().numpy()
        want = oracle.kpconv_forward(q, s, idx72[:, lo:hi].copy(), x, w, kp, 0.3)
        assert np.abs(got - want).max() <= FEAT_RTOL * np.abs(want).max(), (lo, hi)


def test_kpconv_tensor_core_wide_rows():
    """More than 64 neighbour columns (KITTI limits 68 / 74): three 32-slot rounds per row."""
    rng = np.random.default_rng(33)
    ns, nq, H, c = 900, 400, 74, 128
    s = rng.uniform(0, 1, size=(ns, 3)).astype(np.float32)
    q = s[:nq]
    idx = np.sort(rng.integers(0, ns + 1, size=(nq, H)), axis=1)   # shadows at the end, like the real pyramids
    x = rng.normal(size=(ns, c)).astype(np.float32)
    w = (rng.normal(size=(15, c, c)) / np.sqrt(15 * c)).astype(np.float32)
    kp = (rng.normal(size=(15, 3)) * 0.15).astype(np.float32)
    out = ops.kpconv_forward(_t(q), _t(s), _t(idx), _t(x), _t(w), _t(kp), 0.3, mode=1).cpu().numpy()
    exact = oracle.kpconv_forward(q, s, idx, x, w, kp, 0.3)
    assert np.abs(out - exact).max() <= FEAT_RTOL * np.abs(exact).max()


def test_kpconv_full_size_properties():
    """BASELINE-size layer (level 1 of the 32-pair bench step: ~300k queries, H = 40, C = 64): properties that do
    not need a CPU evaluation -- linearity in the features, invariance to the order of a row's neighbours,
    independence of rows, shadow columns contributing nothing -- plus a sampled comparison with the oracle."""
    rng = np.random.default_rng(99)
    n, H, c = 300_000, 40, 64
    # a random walk: rows close in index are close in space, so the +-60-row neighbourhoods below give kernel-point
    # influences of every size between 0 and 1
    pts = np.cumsum(rng.normal(0, 0.01, size=(n, 3)), axis=0).astype(np.float32)
    idx = (np.arange(n)[:, None] + rng.integers(-60, 60, size=(n, H))).clip(0, n - 1)
    idx[rng.uniform(size=(n, H)) < 0.2] = n                                # shadow entries anywhere in the row
    idx.sort(axis=1)                                                        # shadows last, like the searcher writes
    x1 = rng.normal(size=(n, c)).astype(np.float32)
    x2 = rng.normal(size=(n, c)).astype(np.float32)
    w = (rng.normal(size=(15, c, c)) / np.sqrt(15 * c)).astype(np.float32)
    kp = (rng.normal(size=(15, 3)) * 0.05).astype(np.float32)
    ext = 0.12
    tp, tw, tk = _t(pts), _t(w), _t(kp)
    f = lambda x, ind: ops.kpconv_forward(tp, tp, ind, x, tw, tk, ext, mode=1)
    ti = _t(idx)
    y1, y2 = f(_t(x1), ti), f(_t(x2), ti)
    scale = y1.abs().max().item()
    assert scale > 1e-3
    # linearity.  The neighbour count divides by #{h: rowsum(x) > 0} (kpconv_blocks.py:409-412), which is not
    # linear, so the check uses inputs whose row sums are all positive for x1, x2 and the combination
    p1, p2 = np.abs(x1) + 0.1, np.abs(x2) + 0.1
    z1, z2, z12 = f(_t(p1), ti), f(_t(p2), ti), f(_t(2.0 * p1 + 0.5 * p2), ti)
    assert (z12 - (2.0 * z1 + 0.5 * z2)).abs().max().item() <= 2e-5 * z12.abs().max().item()
    # neighbour order inside a row does not matter (a sum over h)
    perm = rng.permutation(H)
    yp = f(_t(x1), _t(np.ascontiguousarray(idx[:, perm])))
    assert (yp - y1).abs().max().item() <= 2e-5 * scale
    # rows are independent: changing the neighbours of the second half leaves the first half bit-identical
    idx2 = idx.copy()
    idx2[n // 2:] = n
    yh = f(_t(x1), _t(idx2))
    assert torch.equal(yh[: n // 2], y1[: n // 2]) and yh[n // 2:].abs().max().item() == 0.0
    # sampled rows against the fp64-accumulated oracle
    rows = rng.choice(n, size=512, replace=False)
    exact = oracle.kpconv_forward(pts[rows], pts, idx[rows], x1, w, kp, ext)
    assert np.abs(y1[_t(rows)].cpu().numpy() - exact).max() <= FEAT_RTOL * np.abs(exact).max()


@pytest.mark.parametrize("c", [64, 128, 256])
def test_kpconv_tc_multi_tile_parity_and_determinism(c):
    """The regime the bench runs: C >= 64 (several channel passes -> the influence-fragment cache
    is live), >= 3 tiles per CTA (deferred epilogue, tile-parity double buffer, dispenser wrap-around), H = 40 rows
    from the real searcher (shadow tails of every length), cell-order walk on and off.  Checked: mode 1 against the
    fp32 SIMT anchor (mode 0) over ALL rows, 512 sampled rows against the fp64 oracle, and five repeated launches
    bit-identical -- the race detector of this suite (compute-sanitizer is not usable on the pool)."""
    from superpoints_registration_b200.kernel_points import load_kernels
    rng = np.random.default_rng(1000 + c)
    lens = np.array([14000, 9000, 11000, 8000], dtype=np.int32)
    n, H, r = int(lens.sum()), 40, 0.09
    pts = rng.uniform(0, 1, size=(n, 3)).astype(np.float32)
    pts[:14000, 2] *= 0.3                                      # one dense cloud: full rows next to ragged ones
    tp, tl = _t(pts), _t(lens)
    grid = ops.CellGrid(tp, tl, r)
    x = rng.normal(size=(n, c)).astype(np.float32) * 1.5 + 0.2
    o = ops.instance_norm_lrelu_ex(_t(x), tl, slope=0.1, want_f32=True, kpconv_points=tp)
    f32, prep = o["f32"], o["kpconv"]
    w = _t((rng.normal(size=(15, c, c)) / np.sqrt(15 * c)).astype(np.float32))
    kp = _t(load_kernels(r, 15))
    ext = r * 2.0 / 2.5
    for dtype in (torch.int64, torch.int32):
        idx, mc = grid.query(tp, tl, H, index_dtype=dtype)
        assert int(mc) > H                                      # the dense cloud truncates at the limit
        valid = (idx < n).sum(1)
        assert int(valid.min()) < 10 and int(valid.max()) == H  # shadow tails of many lengths
        anchor = ops.kpconv_forward(tp, tp, idx, f32, w, kp, ext, mode=0)
        scale = anchor.abs().max().item()
        for order in (None, grid.order()):
            outs = [ops.kpconv_forward_prepared(tp, idx, prep, w, kp, ext, order=order) for _ in range(5)]
            for other in outs[1:]:
                assert torch.equal(outs[0], other), (c, dtype, order is not None)
            err = (outs[0] - anchor).abs().max().item()
            assert err <= FEAT_RTOL * scale, (c, dtype, order is not None, err, scale)
        alone = [ops.kpconv_forward(tp, tp, idx, f32, w, kp, ext, mode=1) for _ in range(5)]   # with its own pre-pass
        for other in alone[1:]:
            assert torch.equal(alone[0], other)
        assert (alone[0] - anchor).abs().max().item() <= FEAT_RTOL * scale
    rows = rng.choice(n, size=512, replace=False)
    idx_h = idx.cpu().numpy().astype(np.int64)
    exact = oracle.kpconv_forward(pts[rows], pts, idx_h[rows], f32.cpu().numpy(), w.cpu().numpy(), kp.cpu().numpy(), ext)
    got = outs[0][_t(rows)].cpu().numpy()
    assert np.abs(got - exact).max() <= FEAT_RTOL * np.abs(exact).max()
