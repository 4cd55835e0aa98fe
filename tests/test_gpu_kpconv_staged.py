"""Third-generation tensor-core KPConv (csrc/kpconv_s.cu: neighbour rows staged through per-warp shared-memory rings by
asynchronous copies, ldmatrix fragments) against the fp32 SIMT anchor (mode 0), the fp64 oracle, generation 1 and itself
(repeated launches bit-identical)."""
import numpy as np
import pytest
import torch

import oracle
from superpoints_registration_b200 import _lib, ops
from superpoints_registration_b200.kernel_points import load_kernels

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
FEAT_RTOL = 2e-5


def _t(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(DEV)
    return t if dtype is None else t.to(dtype)


def _case(c, lens, H, r, seed, dense_first=True):
    rng = np.random.default_rng(seed)
    lens = np.asarray(lens, dtype=np.int32)
    n = int(lens.sum())
    pts = rng.uniform(0, 1, size=(n, 3)).astype(np.float32)
    if dense_first:
        pts[:lens[0], 2] *= 0.3
    tp, tl = _t(pts), _t(lens)
    grid = ops.CellGrid(tp, tl, r)
    x = rng.normal(size=(n, c)).astype(np.float32) * 1.5 + 0.2
    o = ops.instance_norm_lrelu_ex(_t(x), tl, slope=0.1, want_f32=True, kpconv_points=tp, kpconv_planar=True)
    w = _t((rng.normal(size=(15, c, c)) / np.sqrt(15 * c)).astype(np.float32))
    kp = _t(load_kernels(r, 15))
    return pts, tp, tl, grid, o["f32"], o["kpconv"], w, kp, r * 2.0 / 2.5


@pytest.mark.parametrize("c", [32, 64, 128, 256])
@pytest.mark.parametrize("H", [24, 40, 74])
def test_staged_kernel_single_and_few_tiles(c, H):
    assert _lib.lib().spr_kpconv_staged_supported(c, H)
    pts, tp, tl, grid, f32, prep, w, kp, ext = _case(c, [700, 300, 41], H, 0.17, 7 * c + H)
    idx, _ = grid.query(tp, tl, H, index_dtype=torch.int32)
    anchor = ops.kpconv_forward(tp, tp, idx, f32, w, kp, ext, mode=0)
    got = ops.kpconv_forward_prepared(tp, idx, prep, w, kp, ext, generation=3)
    scale = anchor.abs().max().item()
    err = (got - anchor).abs().max().item()
    print(f"C={c} H={H}: max err {err:.3e} = {err / scale:.2e} x max|out|")
    assert err <= FEAT_RTOL * scale
    rows = np.random.default_rng(1).choice(len(pts), size=128, replace=False)
    exact = oracle.kpconv_forward(pts[rows], pts, idx.cpu().numpy().astype(np.int64)[rows], f32.cpu().numpy(),
                                  w.cpu().numpy(), kp.cpu().numpy(), ext)
    assert np.abs(got[_t(rows)].cpu().numpy() - exact).max() <= FEAT_RTOL * np.abs(exact).max()


@pytest.mark.parametrize("c", [32, 64, 128, 256])
def test_staged_kernel_multi_tile_parity_and_determinism(c):
    """Several tiles per CTA, every channel pass, real shadow tails, cell-order walk on and off, both index dtypes;
    five repeated launches bit-identical."""
    pts, tp, tl, grid, f32, prep, w, kp, ext = _case(c, [14000, 9000, 11000, 8000], 40, 0.09, 2000 + c)
    n = len(pts)
    for dtype in (torch.int64, torch.int32):
        idx, mc = grid.query(tp, tl, 40, index_dtype=dtype)
        anchor = ops.kpconv_forward(tp, tp, idx, f32, w, kp, ext, mode=0)
        scale = anchor.abs().max().item()
        for order in (None, grid.order()):
            outs = [ops.kpconv_forward_prepared(tp, idx, prep, w, kp, ext, order=order, generation=3) for _ in range(5)]
            for other in outs[1:]:
                assert torch.equal(outs[0], other), (c, dtype, order is not None)
            err = (outs[0] - anchor).abs().max().item()
            assert err <= FEAT_RTOL * scale, (c, dtype, order is not None, err, scale)
        prep1 = ops.instance_norm_lrelu_ex(f32, tl, slope=1.0, want_f32=False, kpconv_points=tp)["kpconv"]   # same rows, gen-1 layout
        with pytest.raises(RuntimeError):
            ops.kpconv_forward_prepared(tp, idx, prep1, w, kp, ext, generation=3)                          # wrong layout is refused


def test_staged_kernel_strided_queries_and_empty_rows():
    """Pooling layers: queries are a different (smaller) cloud than the supports; rows without any neighbour give 0."""
    rng = np.random.default_rng(5)
    c, H = 64, 40
    pts, tp, tl, grid, f32, prep, w, kp, ext = _case(c, [5000, 3000], H, 0.12, 99, dense_first=False)
    ql = np.array([900, 500], dtype=np.int32)
    qp = np.concatenate([pts[:5000][rng.permutation(5000)[:900]], pts[5000:][rng.permutation(3000)[:500]]]) \
        + rng.normal(0, 0.01, size=(1400, 3)).astype(np.float32)
    qp[7] = 5.0                                               # far away from everything: an all-shadow row
    tq = _t(qp.astype(np.float32))
    idx, _ = grid.query(tq, _t(ql), H, index_dtype=torch.int32)
    anchor = ops.kpconv_forward(tq, tp, idx, f32, w, kp, ext, mode=0)
    got = ops.kpconv_forward_prepared(tq, idx, prep, w, kp, ext, generation=3)
    assert (got - anchor).abs().max().item() <= FEAT_RTOL * anchor.abs().max().item()
    assert torch.count_nonzero(got[7]).item() == 0


def test_staged_kernel_against_generation_1_at_wide_rows_and_scales():
    """Rows wider than 64 columns (KITTI-shape levels) and inputs / weights scaled over ten orders of magnitude: the
    operand scaling is shared with generation 1, whose result is the comparison."""
    c, H = 64, 90
    pts, tp, tl, grid, f32, prep, w, kp, ext = _case(c, [6000, 2500], H, 0.2, 321)
    idx, _ = grid.query(tp, tl, H, index_dtype=torch.int32)
    for sx, sw in ((1.0, 1.0), (1e-5, 1e3), (1e4, 1e-4)):
        x = f32 * sx
        p3 = ops.instance_norm_lrelu_ex(x, tl, slope=1.0, want_f32=False, kpconv_points=tp, kpconv_planar=True)["kpconv"]
        p1 = ops.instance_norm_lrelu_ex(x, tl, slope=1.0, want_f32=True, kpconv_points=tp)
        ws = w * sw
        a = ops.kpconv_forward_prepared(tp, idx, p1["kpconv"], ws, kp, ext, generation=1)
        b = ops.kpconv_forward_prepared(tp, idx, p3, ws, kp, ext, generation=3)
        assert (a - b).abs().max().item() <= FEAT_RTOL * a.abs().max().item(), (sx, sw)
