"""Functional wrappers: torch CUDA tensors in, torch CUDA tensors out, every byte of compute in libspr_b200.so.

PyTorch is used here for device memory (allocation through its caching allocator) and for the current
stream only.  Inputs must already live on a CUDA device: there is no CPU path.
"""
from __future__ import annotations

from typing import Optional, Tuple

import os

import torch

from . import _lib


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream() -> int:
    """cudaStream_t of torch's current stream.  The raw accessor is ~20x cheaper than building a Stream object; a
    forward makes ~350 calls."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the B200 hot path has no CPU fallback")


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    _need_cuda(t, name)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _i32c(t: torch.Tensor, name: str) -> torch.Tensor:
    _need_cuda(t, name)
    if t.dtype != torch.int32:
        t = t.to(torch.int32)
    return t.contiguous()


def _ws(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# ------------------------------------------------------------------------------------------------
# preprocessing
# ------------------------------------------------------------------------------------------------

def to_device_async(values, dtype, device) -> torch.Tensor:
    """Small host list -> device tensor through pinned memory, without blocking the host (torch.tensor(...,
    device=cuda) is a pageable copy that waits for everything queued on the stream)."""
    host = torch.tensor(values, dtype=dtype)
    if torch.device(device).type != "cuda":  # host-side bookkeeping objects (tests of the offset logic)
        return host
    return host.pin_memory().to(device, non_blocking=True)


class PendingSubsample:
    """A grid subsampling whose output size is still on its way to the host (see grid_subsample_batch_async)."""

    def __init__(self, out, meta, host_meta, event, b):
        self._out, self._meta, self._host, self._event, self._b = out, meta, host_meta, event, b

    def finish(self) -> Tuple[torch.Tensor, torch.Tensor, list]:
        """-> (points f32[M,3], lengths i32[B] on device, the same lengths as a host list)."""
        self._event.synchronize()
        host = self._host.tolist()
        return self._out[:host[self._b]], self._meta[:self._b], host[:self._b]


SUBSAMPLE_MODES = {"reference": 0, "mean": 1, "first_point": 2}   # SPR_SUBSAMPLE_* of include/spr_b200.h


def grid_subsample_batch_async(points: torch.Tensor, lengths: torch.Tensor, sample_dl: float,
                               mode: str = "reference") -> PendingSubsample:
    """Launch the barycentre voxel subsampling and request its sizes; the caller queues independent work and then
    calls .finish(), so the device never waits for the host to learn M.  mode: 'reference' (the CPU Preprocessor's
    recipe), 'mean' (MinkowskiEngine-style: global lattice, sum / count), 'first_point' (first point of every voxel)."""
    L = _lib.lib()
    pts = _f32c(points, "points")
    lens = _i32c(lengths, "lengths")
    n, b = pts.shape[0], lens.shape[0]
    if pts.dim() != 2 or pts.shape[1] != 3:
        raise RuntimeError("points must have shape (N, 3)")
    out = torch.empty((max(n, 1), 3), dtype=torch.float32, device=pts.device)
    meta = torch.empty(b + 1, dtype=torch.int32, device=pts.device)  # [lengths..., total]
    wsb = L.spr_grid_subsample_workspace_bytes(n, b)
    ws = _ws(wsb, pts.device)
    rc = L.spr_grid_subsample_batch_ex(pts.data_ptr(), lens.data_ptr(), n, b, float(sample_dl), SUBSAMPLE_MODES[mode],
                                       out.data_ptr(), meta.data_ptr(), meta.data_ptr() + 4 * b, ws.data_ptr(),
                                       ws.numel(), _stream())
    _lib.check(rc, "spr_grid_subsample_batch")
    host = torch.empty(b + 1, dtype=torch.int32, pin_memory=True)
    host.copy_(meta, non_blocking=True)
    event = torch.cuda.Event()
    event.record()
    return PendingSubsample(out, meta, host, event, b)


def grid_subsample_batch(points: torch.Tensor, lengths: torch.Tensor, sample_dl: float, mode: str = "reference"
                         ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Barycentre voxel subsampling of stacked clouds. -> (points f32[M,3], lengths i32[B]) on device.

    One host synchronisation (reading M), as in the reference where the lengths come back as a NumPy array
    (cpp_subsampling/wrapper.cpp:300-322).
    """
    pts, lens, _ = grid_subsample_batch_async(points, lengths, sample_dl, mode).finish()
    return pts, lens


class CellGrid:
    """Supports binned into a uniform cell list; answers radius queries (spr_cell_grid_build / spr_radius_query)."""

    def __init__(self, supports: torch.Tensor, s_lengths: torch.Tensor, radius: float):
        L = _lib.lib()
        self.supports = _f32c(supports, "supports")
        self.s_lengths = _i32c(s_lengths, "s_lengths")
        self.radius = float(radius)
        self.ns, self.b = self.supports.shape[0], self.s_lengths.shape[0]
        self.ws = _ws(L.spr_cell_grid_workspace_bytes(self.ns, self.b), self.supports.device)
        rc = L.spr_cell_grid_build(self.supports.data_ptr(), self.s_lengths.data_ptr(), self.ns, self.b, self.radius,
                                   self.ws.data_ptr(), self.ws.numel(), _stream())
        _lib.check(rc, "spr_cell_grid_build")

    def order(self) -> torch.Tensor:
        """int32 [ns]: the supports in (cloud, z, y, x) cell order (a spatially coherent permutation)."""
        if getattr(self, "_order", None) is not None:
            return self._order
        out = torch.empty(self.ns, dtype=torch.int32, device=self.supports.device)
        rc = _lib.lib().spr_cell_grid_order(self.ws.data_ptr(), self.ns, self.b, out.data_ptr(), _stream())
        _lib.check(rc, "spr_cell_grid_order")
        self._order = out
        return out

    def query(self, queries: torch.Tensor, q_lengths: torch.Tensor, limit: int, radius: Optional[float] = None,
              index_dtype: torch.dtype = torch.int64, by_index: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (idx [Nq, limit] index_dtype, max_count i32[1] on device).  by_index: rows keep the first `limit`
        in-radius supports in index order (ball_query) instead of the `limit` nearest."""
        L = _lib.lib()
        q = _f32c(queries, "queries")
        ql = _i32c(q_lengths, "q_lengths")
        if ql.shape[0] != self.b:
            raise RuntimeError("q_batches and s_batches must have the same length")
        r = self.radius if radius is None else float(radius)
        if r > self.radius * (1 + 1e-6):
            raise RuntimeError("query radius larger than the radius the grid was built for")
        nq = q.shape[0]
        idx = torch.empty((nq, limit), dtype=index_dtype, device=q.device)
        mc = torch.empty(1, dtype=torch.int32, device=q.device)
        rc = L.spr_radius_query_ex(q.data_ptr(), ql.data_ptr(), nq, self.b, self.ws.data_ptr(), self.ns, r, int(limit),
                                   1 if by_index else 0, idx.data_ptr(), 1 if index_dtype == torch.int64 else 0,
                                   int(limit), mc.data_ptr(), _stream())
        _lib.check(rc, "spr_radius_query")
        return idx, mc


def radius_neighbors_batch(queries, supports, q_lengths, s_lengths, radius: float, limit: int,
                           index_dtype: torch.dtype = torch.int64):
    grid = CellGrid(supports, s_lengths, radius)
    return grid.query(queries, q_lengths, limit, index_dtype=index_dtype)


# ------------------------------------------------------------------------------------------------
# KPConv and the block epilogues
# ------------------------------------------------------------------------------------------------

def _idx_arg(idx: torch.Tensor):
    _need_cuda(idx, "neighb_inds")
    if idx.dtype not in (torch.int64, torch.int32):
        idx = idx.long()
    if idx.dim() != 2:
        raise RuntimeError("neighb_inds must be 2-D")
    if idx.stride(1) != 1:
        idx = idx.contiguous()
    return idx, (1 if idx.dtype == torch.int64 else 0), idx.stride(0), idx.shape[1]


_TC_CHANNELS = (32, 64, 128, 256)


def default_kpconv_mode(cin: int, cout: int) -> int:
    """1 (tcgen05 tensor-core contraction) where the layer shape has a tensor-core kernel, else 0 (fp32 CUDA-core
    contraction: the Cin=1 stem).  SPR_KPCONV_MODE=0|1 forces one path for experiments."""
    forced = os.environ.get("SPR_KPCONV_MODE")
    if forced is not None:
        return int(forced) if (cin == cout and cin in _TC_CHANNELS) else 0
    return 1 if (cin == cout and cin in _TC_CHANNELS) else 0


def kpconv_forward(q_pts, s_pts, neighb_inds, x, weights, kernel_points, extent: float, mode=None, order=None):
    """order (optional): i32 permutation of the queries = processing order (same result; the Cin = 1 kernel uses it)."""
    L = _lib.lib()
    q, s, xx = _f32c(q_pts, "q_pts"), _f32c(s_pts, "s_pts"), _f32c(x, "x")
    w, kp = _f32c(weights, "weights"), _f32c(kernel_points, "kernel_points")
    idx, is64, stride, H = _idx_arg(neighb_inds)
    nq, ns = q.shape[0], s.shape[0]
    K, cin, cout = w.shape
    if xx.shape[0] != ns or xx.shape[1] != cin:
        raise RuntimeError(f"x must have shape ({ns}, {cin}), got {tuple(xx.shape)}")
    if idx.shape[0] != nq:
        raise RuntimeError("neighb_inds must have one row per query point")
    if mode is None:
        mode = default_kpconv_mode(cin, cout)
    if order is not None:
        order = _i32c(order, "order")
        if order.numel() != nq:
            raise RuntimeError("kpconv_forward: order must have one entry per query")
    out = torch.empty((nq, cout), dtype=torch.float32, device=q.device)
    ws = _ws(L.spr_kpconv_workspace_bytes(nq, ns, cin, cout, K), q.device)
    rc = L.spr_kpconv_forward(q.data_ptr(), s.data_ptr(), idx.data_ptr(), is64, stride, H, xx.data_ptr(), cin,
                              w.data_ptr(), cout, kp.data_ptr(), K, float(extent), out.data_ptr(), nq, ns, int(mode),
                              ws.data_ptr(), ws.numel(), _ptr(order), _stream())
    _lib.check(rc, "spr_kpconv_forward")
    return out


def instance_norm_lrelu(x, lengths, eps: float = 1e-5, slope: float = 1.0, residual=None, out=None):
    """y = LeakyReLU_slope(InstanceNorm_per_cloud(x) [+ residual]);  slope=1 -> no activation."""
    L = _lib.lib()
    xx = _f32c(x, "x")
    lens = _i32c(lengths, "stack_lengths")
    n, c = xx.shape
    res = None if residual is None else _f32c(residual, "residual")
    if out is None:
        out = torch.empty_like(xx)
    ws = _ws(L.spr_instance_norm_workspace_bytes(n, lens.shape[0], c), xx.device)
    rc = L.spr_instance_norm_lrelu(xx.data_ptr(), lens.data_ptr(), n, lens.shape[0], c, float(eps), float(slope),
                                   _ptr(res), out.data_ptr(), ws.data_ptr(), ws.numel(), _stream())
    _lib.check(rc, "spr_instance_norm_lrelu")
    return out


def block_stats(n_rows: int, c: int, device) -> torch.Tensor:
    """Buffer for the (sum, sumsq) of every 16-row block that a producer (gemm_tc, kpconv_forward_prepared) writes
    for the instance normalisation that consumes its rows."""
    return torch.empty(((n_rows + 15) // 16, c, 2), dtype=torch.float32, device=device)


def instance_norm_lrelu_ex(x, lengths, eps: float = 1e-5, slope: float = 1.0, residual=None, want_f32: bool = True,
                           want_image: bool = False, kpconv_points=None, stats16=None, kpconv_planar: bool = False,
                           residual_stats16=None):
    """InstanceNorm (+ residual) + LeakyReLU with format-aware outputs.  Returns a dict with any of
    'f32' (rows), 'image' (operand image of the next tensor-core GEMM, K = c), 'kpconv' (PreparedFeatures for
    kpconv_forward_prepared; needs kpconv_points = the [n,3] points the rows belong to).
    residual_stats16: `residual` is a raw producer output (with these 16-row block sums) and is instance-normalised,
    without activation, inside the same kernel before it is added (the projected shortcut of a bottleneck block)."""
    L = _lib.lib()
    xx = _f32c(x, "x")
    lens = _i32c(lengths, "stack_lengths")
    n, c = xx.shape
    res = None if residual is None else _f32c(residual, "residual")
    out = {}
    f32 = torch.empty_like(xx) if want_f32 else None
    img = gemm_a_image(n, c, xx.device) if want_image else None
    x16 = pts4 = amax = pts = None
    if kpconv_points is not None:
        pts = _f32c(kpconv_points, "kpconv_points")
        x16 = torch.empty((n, c), dtype=torch.int32, device=xx.device)
        pts4 = torch.empty((n, 4), dtype=torch.float32, device=xx.device)
        amax = torch.empty(1, dtype=torch.int32, device=xx.device)
    if stats16 is not None and tuple(stats16.shape) != ((n + 15) // 16, c, 2):
        raise RuntimeError("instance_norm_lrelu_ex: stats16 does not match x")
    if residual_stats16 is not None and (res is None or tuple(residual_stats16.shape) != ((n + 15) // 16, c, 2)):
        raise RuntimeError("instance_norm_lrelu_ex: residual_stats16 does not match the residual")
    if res is not None and tuple(res.shape) != (n, c):
        raise RuntimeError("instance_norm_lrelu_ex: residual does not match x")
    ws = _ws(L.spr_instance_norm_workspace_bytes(n, lens.shape[0], c), xx.device)
    rc = L.spr_instance_norm_lrelu_ex(xx.data_ptr(), lens.data_ptr(), n, lens.shape[0], c, float(eps), float(slope),
                                      _ptr(res), _ptr(f32), _ptr(img), A_SCALE, _ptr(x16), _ptr(pts4), _ptr(pts),
                                      _ptr(amax), _ptr(stats16), _ptr(residual_stats16), 1 if kpconv_planar else 0,
                                      ws.data_ptr(), ws.numel(),
                                      _stream())
    _lib.check(rc, "spr_instance_norm_lrelu_ex")
    if f32 is not None:
        out["f32"] = f32
    if img is not None:
        out["image"] = img
    if x16 is not None:
        out["kpconv"] = PreparedFeatures(x16, pts4, amax, c, planar=kpconv_planar)
    return out


class PreparedFeatures:
    """Inputs of the tensor-core KPConv already in kernel format (pre-split rows, packed points, max|x|)."""

    def __init__(self, x16, pts4, amax, c, planar: bool = False):
        # planar: x16 holds per group of 32 channels [32 hi | 32 lo] halves (generation-2 kernel) instead of one
        # (hi | lo << 16) word per channel (generation 1)
        self.x16, self.pts4, self.amax, self.c, self.planar = x16, pts4, amax, c, planar


class KPConvWeightImage:
    def __init__(self, weights: torch.Tensor):
        L = _lib.lib()
        w = _f32c(weights.detach(), "weights")
        self.c = w.shape[1]
        self.img = torch.empty(L.spr_kpconv_weight_image_bytes(self.c), dtype=torch.uint8, device=w.device)
        self.amax = torch.empty(1, dtype=torch.int32, device=w.device)
        rc = L.spr_kpconv_prepare_weights(w.data_ptr(), self.c, self.img.data_ptr(), self.amax.data_ptr(), _stream())
        _lib.check(rc, "spr_kpconv_prepare_weights")
        self.key = (weights.data_ptr(), weights._version)


def kpconv_weight_image(weights: torch.Tensor) -> KPConvWeightImage:
    key = (weights.data_ptr(), weights._version)
    wi = getattr(weights, "_spr_kpconv_image", None)   # cached on the tensor, see weight_image
    if wi is None or wi.key != key:
        wi = KPConvWeightImage(weights)
        weights._spr_kpconv_image = wi
    return wi


class KPConvGatherWeightImage:
    """Weight image of the gather kernel (csrc/kpconv_g.cu: channel-major K order)."""

    def __init__(self, weights: torch.Tensor):
        L = _lib.lib()
        w = _f32c(weights.detach(), "weights")
        self.c = w.shape[1]
        self.img = torch.empty(L.spr_kpconv_gather_weight_image_bytes(self.c), dtype=torch.uint8, device=w.device)
        self.amax = torch.empty(1, dtype=torch.int32, device=w.device)
        rc = L.spr_kpconv_gather_prepare_weights(w.data_ptr(), self.c, self.img.data_ptr(), self.amax.data_ptr(), _stream())
        _lib.check(rc, "spr_kpconv_gather_prepare_weights")
        self.key = (weights.data_ptr(), weights._version)


def kpconv_gather_weight_image(weights: torch.Tensor) -> KPConvGatherWeightImage:
    key = (weights.data_ptr(), weights._version)
    wi = getattr(weights, "_spr_kpconv_gather_image", None)
    if wi is None or wi.key != key:
        wi = KPConvGatherWeightImage(weights)
        weights._spr_kpconv_gather_image = wi
    return wi


class KPConvStagedWeightImage:
    """Weight image of the staged kernel (csrc/kpconv_s.cu: channels of a pass permuted to its accumulator layout)."""

    def __init__(self, weights: torch.Tensor):
        L = _lib.lib()
        w = _f32c(weights.detach(), "weights")
        self.c = w.shape[1]
        self.img = torch.empty(L.spr_kpconv_staged_weight_image_bytes(self.c), dtype=torch.uint8, device=w.device)
        self.amax = torch.empty(1, dtype=torch.int32, device=w.device)
        rc = L.spr_kpconv_staged_prepare_weights(w.data_ptr(), self.c, self.img.data_ptr(), self.amax.data_ptr(), _stream())
        _lib.check(rc, "spr_kpconv_staged_prepare_weights")
        self.key = (weights.data_ptr(), weights._version)


def kpconv_staged_weight_image(weights: torch.Tensor) -> KPConvStagedWeightImage:
    key = (weights.data_ptr(), weights._version)
    wi = getattr(weights, "_spr_kpconv_staged_image", None)
    if wi is None or wi.key != key:
        wi = KPConvStagedWeightImage(weights)
        weights._spr_kpconv_staged_image = wi
    return wi


def kpconv_kernel_generation(c: int, H: int) -> int:
    """Which tensor-core KPConv kernel a layer of c channels and H neighbour columns runs on: 1 = csrc/kpconv_tc.cu,
    2 = csrc/kpconv_g.cu (asynchronous gather + both products on tcgen05), 3 = csrc/kpconv_s.cu (rows staged through
    per-warp shared-memory rings).  SPR_KPCONV_GEN=1|2|3 selects the preference (default 3: 1.01-1.26x of generation 1
    at the bench layer shapes, profiles/r2c_kpconv_gen_bench_32pairs.log); a shape the preferred kernel does not support
    falls back to generation 1, which supports every shape of the tensor-core path."""
    want = int(os.environ.get("SPR_KPCONV_GEN", str(DEFAULT_KPCONV_GEN)))
    if want == 3 and _lib.lib().spr_kpconv_staged_supported(int(c), int(H)):
        return 3
    if want == 2 and _lib.lib().spr_kpconv_gather_supported(int(c), int(H)):
        return 2
    return 1


DEFAULT_KPCONV_GEN = 3


def kpconv_forward_prepared(q_pts, neighb_inds, feats: PreparedFeatures, weights, kernel_points, extent: float,
                            order=None, generation=None):
    """KPConv on the tensor-core path from inputs prepared by instance_norm_lrelu_ex (no pre-pass kernels).
    order: optional int32 permutation of the queries (CellGrid.order()) = processing order."""
    L = _lib.lib()
    if generation is None:
        generation = 1
        if feats.planar:  # planar rows are read by generations 2 and 3 only
            generation = kpconv_kernel_generation(feats.c, neighb_inds.shape[1])
            if generation == 1:
                generation = 3
    if generation == 3:
        if not L.spr_kpconv_staged_supported(int(feats.c), int(neighb_inds.shape[1])):
            raise RuntimeError("kpconv_forward_prepared: the generation-3 kernel does not support this shape")
        if not feats.planar:
            raise RuntimeError("kpconv_forward_prepared: the generation-3 kernel needs planar pre-split rows "
                               "(instance_norm_lrelu_ex(..., kpconv_planar=True))")
        return _kpconv_forward_staged(q_pts, neighb_inds, feats, weights, kernel_points, extent, order)
    if generation == 2:
        if not L.spr_kpconv_gather_supported(int(feats.c), int(neighb_inds.shape[1])):
            raise RuntimeError("kpconv_forward_prepared: the generation-2 kernel does not support this shape")
        if not feats.planar:
            raise RuntimeError("kpconv_forward_prepared: the generation-2 kernel needs planar pre-split rows "
                               "(instance_norm_lrelu_ex(..., kpconv_planar=True))")
        return _kpconv_forward_gather(q_pts, neighb_inds, feats, weights, kernel_points, extent, order)
    if feats.planar:
        raise RuntimeError("kpconv_forward_prepared: the generation-1 kernel needs interleaved pre-split rows")
    q = _f32c(q_pts, "q_pts")
    kp = _f32c(kernel_points, "kernel_points")
    idx, is64, stride, H = _idx_arg(neighb_inds)
    nq, ns = q.shape[0], feats.x16.shape[0]
    if order is not None and (order.dtype != torch.int32 or order.shape[0] != nq or not order.is_cuda):
        raise RuntimeError("kpconv_forward_prepared: order must be an int32 CUDA tensor with one entry per query")
    wi = kpconv_weight_image(weights)
    if wi.c != feats.c:
        raise RuntimeError("kpconv_forward_prepared: channel mismatch between features and weights")
    out = torch.empty((nq, feats.c), dtype=torch.float32, device=q.device)
    sb = L.spr_kpconv_scratch_bytes(H, feats.c)
    scratch = _ws(sb, q.device) if sb else None
    rc = L.spr_kpconv_forward_prepared(q.data_ptr(), idx.data_ptr(), is64, stride, H, feats.pts4.data_ptr(),
                                       feats.x16.data_ptr(), feats.amax.data_ptr(), feats.c, wi.img.data_ptr(),
                                       wi.amax.data_ptr(), kp.data_ptr(), float(extent), out.data_ptr(), nq, ns,
                                       _ptr(scratch), _ptr(order), _stream())
    _lib.check(rc, "spr_kpconv_forward_prepared")
    return out


def _kpconv_forward_staged(q_pts, neighb_inds, feats: PreparedFeatures, weights, kernel_points, extent, order):
    L = _lib.lib()
    q = _f32c(q_pts, "q_pts")
    kp = _f32c(kernel_points, "kernel_points")
    idx, is64, stride, H = _idx_arg(neighb_inds)
    nq, ns = q.shape[0], feats.x16.shape[0]
    if order is not None and (order.dtype != torch.int32 or order.shape[0] != nq or not order.is_cuda):
        raise RuntimeError("kpconv_forward_prepared: order must be an int32 CUDA tensor with one entry per query")
    wi = kpconv_staged_weight_image(weights)
    if wi.c != feats.c:
        raise RuntimeError("kpconv_forward_prepared: channel mismatch between features and weights")
    out = torch.empty((nq, feats.c), dtype=torch.float32, device=q.device)
    rc = L.spr_kpconv_forward_staged(q.data_ptr(), idx.data_ptr(), is64, stride, H, feats.pts4.data_ptr(),
                                     feats.x16.data_ptr(), feats.amax.data_ptr(), feats.c, wi.img.data_ptr(),
                                     wi.amax.data_ptr(), kp.data_ptr(), float(extent), out.data_ptr(), nq, ns,
                                     _ptr(order), _stream())
    _lib.check(rc, "spr_kpconv_forward_staged")
    return out


def _kpconv_forward_gather(q_pts, neighb_inds, feats: PreparedFeatures, weights, kernel_points, extent, order):
    L = _lib.lib()
    q = _f32c(q_pts, "q_pts")
    kp = _f32c(kernel_points, "kernel_points")
    idx, is64, stride, H = _idx_arg(neighb_inds)
    nq, ns = q.shape[0], feats.x16.shape[0]
    if order is not None and (order.dtype != torch.int32 or order.shape[0] != nq or not order.is_cuda):
        raise RuntimeError("kpconv_forward_prepared: order must be an int32 CUDA tensor with one entry per query")
    wi = kpconv_gather_weight_image(weights)
    if wi.c != feats.c:
        raise RuntimeError("kpconv_forward_prepared: channel mismatch between features and weights")
    out = torch.empty((nq, feats.c), dtype=torch.float32, device=q.device)
    sb = L.spr_kpconv_gather_scratch_bytes(feats.c)
    scratch = _ws(sb, q.device) if sb else None
    rc = L.spr_kpconv_forward_gather(q.data_ptr(), idx.data_ptr(), is64, stride, H, feats.pts4.data_ptr(),
                                     feats.x16.data_ptr(), feats.amax.data_ptr(), feats.c, wi.img.data_ptr(),
                                     wi.amax.data_ptr(), kp.data_ptr(), float(extent), out.data_ptr(), nq, ns,
                                     _ptr(scratch), _ptr(order), _stream())
    _lib.check(rc, "spr_kpconv_forward_gather")
    return out


def max_pool(x, inds, order=None):
    """order (optional): i32 permutation of the pooled points = the order they are processed in (same result)."""
    L = _lib.lib()
    xx = _f32c(x, "x")
    idx, is64, stride, H = _idx_arg(inds)
    nq, (ns, c) = idx.shape[0], xx.shape
    out = torch.empty((nq, c), dtype=torch.float32, device=xx.device)
    if order is not None:
        order = _i32c(order, "order")
        if order.numel() != nq:
            raise RuntimeError("max_pool: order must have one entry per pooled point")
    rc = L.spr_max_pool(xx.data_ptr(), idx.data_ptr(), is64, stride, H, nq, ns, c, out.data_ptr(), _ptr(order), _stream())
    _lib.check(rc, "spr_max_pool")
    return out


# ------------------------------------------------------------------------------------------------
# matching and pose
# ------------------------------------------------------------------------------------------------

_small_cache = {}


def _memo(key, build):
    """Tiny LRU for device-side descriptions of a batch that depend only on its (host-known) cloud sizes -- offset
    tables, attention tile lists: a stream of same-shaped batches (and every CUDA-graph replay) rebuilds nothing."""
    hit = _small_cache.get(key)
    if hit is None:
        if len(_small_cache) >= 64:
            _small_cache.pop(next(iter(_small_cache)))
        hit = _small_cache[key] = build()
    return hit


def packed_pairs(src_lens, tgt_lens, device) -> "PackedPairs":
    src_lens, tgt_lens = tuple(int(v) for v in src_lens), tuple(int(v) for v in tgt_lens)
    return _memo(("pairs", src_lens, tgt_lens, str(device)), lambda: PackedPairs(src_lens, tgt_lens, device))


class PackedPairs:
    """Offsets describing P pairs packed back to back: src rows, tgt rows, N_p x M_p matrices, outputs."""

    def __init__(self, src_lens, tgt_lens, device):
        self.src_lens = [int(v) for v in src_lens]
        self.tgt_lens = [int(v) for v in tgt_lens]
        self.P = len(self.src_lens)
        so, to, co, oo = [0], [0], [0], [0]
        for n, m in zip(self.src_lens, self.tgt_lens):
            so.append(so[-1] + n)
            to.append(to[-1] + m)
            co.append(co[-1] + n * m)
            oo.append(oo[-1] + (m if n > m else n))
        self.total_src, self.total_tgt, self.total_corr, self.total_out = so[-1], to[-1], co[-1], oo[-1]
        self.max_n, self.max_m = max(self.src_lens), max(self.tgt_lens)
        self.h_so, self.h_to, self.h_co, self.h_oo = so, to, co, oo
        # one pinned staging buffer, one asynchronous copy: int64 so that the N x M offsets fit
        P1 = self.P + 1
        packed = to_device_async(so + to + oo + co, torch.int64, device)
        i32 = packed[:3 * P1].to(torch.int32)
        self.so, self.to, self.oo = i32[:P1], i32[P1:2 * P1], i32[2 * P1:]
        self.co = packed[3 * P1:]


def dual_softmax_match(src_feats, tgt_feats, pairs: PackedPairs, want_attn: bool = False):
    """-> corr (packed), attn (packed or None), val f32[total_out], ind i64[total_out]."""
    L = _lib.lib()
    s, t = _f32c(src_feats, "src_feats"), _f32c(tgt_feats, "tgt_feats")
    D = s.shape[1]
    dev = s.device
    corr = torch.empty(pairs.total_corr, dtype=torch.float32, device=dev)
    attn = torch.empty(pairs.total_corr, dtype=torch.float32, device=dev) if want_attn else None
    val = torch.empty(pairs.total_out, dtype=torch.float32, device=dev)
    ind = torch.empty(pairs.total_out, dtype=torch.int64, device=dev)
    ws = _ws(L.spr_match_workspace_bytes(pairs.total_src, pairs.total_tgt, pairs.P), dev)
    rc = L.spr_dual_softmax_match(s.data_ptr(), t.data_ptr(), pairs.so.data_ptr(), pairs.to.data_ptr(),
                                  pairs.co.data_ptr(), pairs.oo.data_ptr(), pairs.P, pairs.total_src, pairs.total_tgt, D,
                                  pairs.max_n, pairs.max_m, corr.data_ptr(), _ptr(attn), val.data_ptr(), ind.data_ptr(),
                                  ws.data_ptr(), ws.numel(), _stream())
    _lib.check(rc, "spr_dual_softmax_match")
    return corr, attn, val, ind


def sinkhorn_weighted_targets(corr, pairs: PackedPairs, tgt_xyz, softplus_alpha: float, exp_beta: float, n_iters: int,
                              slack: bool = True):
    """-> weighted_tgt f32[total_src,3], weights f32[total_src]."""
    L = _lib.lib()
    txyz = _f32c(tgt_xyz, "tgt_xyz")
    dev = txyz.device
    wt = torch.empty((pairs.total_src, 3), dtype=torch.float32, device=dev)
    w = torch.empty(pairs.total_src, dtype=torch.float32, device=dev)
    ws = _ws(L.spr_sinkhorn_workspace_bytes(pairs.total_src, pairs.total_tgt, pairs.P), dev)
    rc = L.spr_sinkhorn_weighted_targets(corr.data_ptr(), pairs.co.data_ptr(), pairs.so.data_ptr(), pairs.to.data_ptr(),
                                         pairs.P, pairs.total_src, pairs.total_tgt, pairs.max_n, pairs.max_m,
                                         txyz.data_ptr(), float(softplus_alpha), float(exp_beta), int(n_iters),
                                         1 if slack else 0, wt.data_ptr(), w.data_ptr(), ws.data_ptr(), ws.numel(),
                                         _stream())
    _lib.check(rc, "spr_sinkhorn_weighted_targets")
    return wt, w


def sinkhorn_affinity(affinity, pairs: PackedPairs, n_iters: int, slack: bool = True, tgt_xyz=None,
                      want_log_perm: bool = True):
    """Sinkhorn normalisation (se3_torch.py:166-202) of packed per-pair affinity matrices.
    -> (log_perm packed or None, weighted_tgt f32[total_src,3] or None, weights f32[total_src] or None)."""
    L = _lib.lib()
    a = _f32c(affinity, "affinity").reshape(-1)
    if a.numel() != pairs.total_corr:
        raise RuntimeError("sinkhorn_affinity: affinity does not match the pair sizes")
    dev = a.device
    logp = torch.empty_like(a) if want_log_perm else None
    txyz = wt = w = None
    if tgt_xyz is not None:
        txyz = _f32c(tgt_xyz, "tgt_xyz")
        wt = torch.empty((pairs.total_src, 3), dtype=torch.float32, device=dev)
        w = torch.empty(pairs.total_src, dtype=torch.float32, device=dev)
    ws = _ws(L.spr_sinkhorn_workspace_bytes(pairs.total_src, pairs.total_tgt, pairs.P), dev)
    rc = L.spr_sinkhorn_affinity(a.data_ptr(), pairs.co.data_ptr(), pairs.so.data_ptr(), pairs.to.data_ptr(), pairs.P,
                                 pairs.total_src, pairs.total_tgt, pairs.max_n, pairs.max_m, int(n_iters),
                                 1 if slack else 0, _ptr(logp), _ptr(txyz), _ptr(wt), _ptr(w), ws.data_ptr(), ws.numel(),
                                 _stream())
    _lib.check(rc, "spr_sinkhorn_affinity")
    return logp, wt, w


def weighted_procrustes(a, b, w, offsets: torch.Tensor) -> torch.Tensor:
    """Packed correspondences -> poses f32[P,3,4]."""
    L = _lib.lib()
    aa, bb = _f32c(a, "a"), _f32c(b, "b")
    ww = None if w is None else _f32c(w, "weights")
    offs = _i32c(offsets, "offsets")
    P = offs.shape[0] - 1
    out = torch.empty((P, 3, 4), dtype=torch.float32, device=aa.device)
    rc = L.spr_weighted_procrustes(aa.data_ptr(), bb.data_ptr(), _ptr(ww), offs.data_ptr(), P, out.data_ptr(), _stream())
    _lib.check(rc, "spr_weighted_procrustes")
    return out


def gather_rows3(src, ind, row_base) -> torch.Tensor:
    L = _lib.lib()
    s = _f32c(src, "src")
    _need_cuda(ind, "ind")
    ii = ind.long().contiguous()
    rb = _i32c(row_base, "row_base")
    out = torch.empty((ii.shape[0], 3), dtype=torch.float32, device=s.device)
    rc = L.spr_gather_rows3(s.data_ptr(), s.shape[0], ii.data_ptr(), rb.data_ptr(), ii.shape[0], out.data_ptr(), _stream())
    _lib.check(rc, "spr_gather_rows3")
    return out


def top2_ratio(attn, pairs: PackedPairs, lowe_thres: float):
    """RegTR.ratio_test (qk_regtr_full.py:370-384) on the packed attention -> val f32[total_out], ind i64[total_out]."""
    for n, m in zip(pairs.src_lens, pairs.tgt_lens):
        if (n if n > m else m) < 2:
            raise RuntimeError("ratio test: selected index k out of range (the reduced axis needs two entries)")
    a = _f32c(attn, "attn")
    val = torch.empty(pairs.total_out, dtype=torch.float32, device=a.device)
    ind = torch.empty(pairs.total_out, dtype=torch.int64, device=a.device)
    rc = _lib.lib().spr_top2_ratio(a.data_ptr(), pairs.co.data_ptr(), pairs.so.data_ptr(), pairs.to.data_ptr(),
                                   pairs.oo.data_ptr(), pairs.P, pairs.total_out, float(lowe_thres), val.data_ptr(),
                                   ind.data_ptr(), _stream())
    _lib.check(rc, "spr_top2_ratio")
    return val, ind


def inlier_reweight(a, b, w, poses, offsets, acceptance_radius: float) -> torch.Tensor:
    """RegTR.recompute_weights (qk_regtr_full.py:386-391) for packed pairs."""
    aa, bb, ww = _f32c(a, "a"), _f32c(b, "b"), _f32c(w, "weights")
    pp, offs = _f32c(poses, "poses"), _i32c(offsets, "offsets")
    out = torch.empty_like(ww)
    rc = _lib.lib().spr_inlier_reweight(aa.data_ptr(), bb.data_ptr(), ww.data_ptr(), pp.data_ptr(), offs.data_ptr(),
                                        offs.shape[0] - 1, ww.shape[0], float(acceptance_radius), out.data_ptr(),
                                        _stream())
    _lib.check(rc, "spr_inlier_reweight")
    return out


def local_global_registration(a, b, w, poses, offsets, acceptance_radius: float, num_refinement_steps: int):
    """RegTR.local_global_registration (qk_regtr_full.py:393-398): alternate inlier re-weighting and the pose solve."""
    for _ in range(int(num_refinement_steps)):
        w = inlier_reweight(a, b, w, poses, offsets, acceptance_radius)
        poses = weighted_procrustes(a, b, w, offsets)
    return poses


def ransac(a, b, w, offsets, sample_idx: torch.Tensor):
    """RegTR.ransac (qk_regtr_full.py:400-421) for packed pairs.  sample_idx i64 [P, n_hypotheses, sample_size]:
    per pair, row numbers local to the pair (the reference draws torch.randint(0, N, (100,)) 500 times).
    -> poses f32[P,3,4], loss f32[P,n_hypotheses], best i32[P]."""
    aa, bb, ww = _f32c(a, "a"), _f32c(b, "b"), _f32c(w, "weights")
    offs = _i32c(offsets, "offsets")
    P = offs.shape[0] - 1
    _need_cuda(sample_idx, "sample_idx")
    if sample_idx.dim() != 3 or sample_idx.shape[0] != P:
        raise RuntimeError("ransac: sample_idx must be [pairs, hypotheses, sample_size]")
    H, S = int(sample_idx.shape[1]), int(sample_idx.shape[2])
    flat = sample_idx.long().reshape(-1)
    base = offs[:-1].repeat_interleave(H * S)
    sa, sb = gather_rows3(aa, flat, base), gather_rows3(bb, flat, base)
    sw = ww[(flat + base).long()]
    seg = torch.arange(0, P * H * S + 1, S, dtype=torch.int32, device=aa.device)
    hyp = weighted_procrustes(sa, sb, sw, seg)                     # [P*H, 3, 4]
    loss = torch.empty((P, H), dtype=torch.float32, device=aa.device)
    out = torch.empty((P, 3, 4), dtype=torch.float32, device=aa.device)
    best = torch.empty(P, dtype=torch.int32, device=aa.device)
    rc = _lib.lib().spr_select_hypothesis(aa.data_ptr(), bb.data_ptr(), offs.data_ptr(), P, hyp.data_ptr(), H,
                                          loss.data_ptr(), out.data_ptr(), best.data_ptr(), _stream())
    _lib.check(rc, "spr_select_hypothesis")
    return out, loss, best


# ------------------------------------------------------------------------------------------------
# cross-encoder building blocks (packed tokens)
# ------------------------------------------------------------------------------------------------

def split_f16(x: torch.Tensor, n_scaled: int = 0, scale: float = 1.0):
    """fp32 [rows, cols] -> fp16 (hi, lo) planes with x = hi + lo to ~22 bits; columns < n_scaled are pre-multiplied
    by `scale`."""
    L = _lib.lib()
    xx = _f32c(x, "x")
    rows, cols = xx.shape
    hi = torch.empty((rows, cols), dtype=torch.float16, device=xx.device)
    lo = torch.empty_like(hi)
    rc = L.spr_split_f16(xx.data_ptr(), rows, cols, cols, hi.data_ptr(), lo.data_ptr(), cols, int(n_scaled), float(scale),
                         _stream())
    _lib.check(rc, "spr_split_f16")
    return hi, lo


def attention_generation() -> int:
    """2 (default): csrc/attention_tc.cu -- tcgen05 products, scores in tensor memory, tiles of up to 128 queries;
    1 (SPR_ATTENTION_GEN=1): csrc/attention.cu -- the warp-level tensor path of round 1, tiles of up to 64 queries."""
    return 1 if os.environ.get("SPR_ATTENTION_GEN", "2") == "1" else 2


def attention_tiles(q_offsets, q_lens, kv_offsets, kv_lens, device, block_q: Optional[int] = None,
                    pad_multiple: int = 1) -> torch.Tensor:
    """Tile list of spr_attention_varlen(_tc): one row {first query row, rows, first key row, key rows} per block_q
    queries (default: the tile height of the selected kernel generation).  pad_multiple > 1 rounds the list up with
    empty entries (0 query rows: skipped by the kernels) so that launches of similar inputs share a grid size."""
    if block_q is None:
        block_q = 128 if attention_generation() == 2 else 64
    key = ("tiles", tuple(q_offsets), tuple(q_lens), tuple(kv_offsets), tuple(kv_lens), str(device), block_q, pad_multiple)
    return _memo(key, lambda: _attention_tiles(q_offsets, q_lens, kv_offsets, kv_lens, device, block_q, pad_multiple))


def _attention_tiles(q_offsets, q_lens, kv_offsets, kv_lens, device, block_q, pad_multiple=1):
    rows = []
    for qo, qn, ko, kn in zip(q_offsets, q_lens, kv_offsets, kv_lens):
        if qn <= 0:
            continue
        if kn <= 0:
            raise RuntimeError("attention over an empty key segment")
        for q0 in range(0, qn, block_q):
            rows.append((qo + q0, min(block_q, qn - q0), ko, kn))
    while len(rows) % pad_multiple:
        rows.append((0, 0, 0, 0))
    return to_device_async(rows, torch.int32, device)


def attention_varlen(hi: torch.Tensor, lo: torch.Tensor, tiles: torch.Tensor, n_heads: int, q_col: int, k_col: int,
                     v_col: int, d_model: int, out_image: Optional[torch.Tensor] = None, image_scale: float = 1.0,
                     generation: Optional[int] = None):
    """Multi-head attention over packed tokens; hi/lo are the fp16 planes of the (pre-scaled) QKV projection.
    With out_image the result is written as the A image of the output projection (gemm_tc) instead of fp32 rows.
    `tiles` must have been built for the kernel generation that runs (attention_tiles: 128 / 64 query rows)."""
    L = _lib.lib()
    _need_cuda(hi, "hi")
    if hi.dtype != torch.float16 or lo.dtype != torch.float16 or hi.shape != lo.shape or not hi.is_contiguous() \
            or not lo.is_contiguous():
        raise RuntimeError("attention_varlen: hi/lo must be contiguous fp16 tensors of the same shape")
    rows, ld = hi.shape
    head_dim = d_model // n_heads
    out = None if out_image is not None else torch.empty((rows, d_model), dtype=torch.float32, device=hi.device)
    gen = attention_generation() if generation is None else generation
    fn = L.spr_attention_varlen_tc if gen == 2 else L.spr_attention_varlen
    rc = fn(hi.data_ptr(), lo.data_ptr(), ld, q_col, k_col, v_col, n_heads, head_dim, tiles.data_ptr(), tiles.shape[0],
            _ptr(out), d_model, _ptr(out_image), float(image_scale), _stream())
    _lib.check(rc, "spr_attention_varlen_tc" if gen == 2 else "spr_attention_varlen")
    return out_image if out is None else out


# ---- tensor-core dense layers (gemm_tc.cu) ---------------------------------------------------------
# Power-of-two scale of activations inside the fp16 (hi, lo) operand images: values up to 65504 / 16 = 4094 are
# representable.  Every activation on the path is normalised (InstanceNorm / LayerNorm outputs, soft-max averages,
# ReLU of a LayerNorm-fed Linear), far below that; a value beyond it turns into NaN (never a silently wrong number)
# and raises the sticky device flag that check_numerics() reports.
A_SCALE = 16.0
OUT_F32, OUT_PLANES, OUT_AIMG = 0, 1, 2


def check_numerics(reset: bool = True) -> None:
    """Raise if a kernel since the last check had to write an fp16 operand image outside its range (one device
    synchronisation; RegTR.forward calls it when cfg.check_numerics is set)."""
    flags = _lib.numeric_flags(reset)
    if flags & _lib.FLAG_FP16_OVERFLOW:
        raise FloatingPointError(
            f"an activation exceeded the fp16 operand range of the tensor-core GEMMs (|x| > {65504.0 / A_SCALE:.0f}): "
            "the affected outputs are NaN")


def gemm_a_image(T: int, K: int, device) -> torch.Tensor:
    return torch.empty(_lib.lib().spr_gemm_a_image_bytes(int(T), int(K)), dtype=torch.uint8, device=device)


class WeightImage:
    """fp16 (hi | lo) shared-memory image of an nn.Linear weight [N, K], built once per weight version."""

    def __init__(self, weight: torch.Tensor):
        L = _lib.lib()
        w = _f32c(weight.detach(), "weight")
        self.N, self.K = w.shape
        amax = float(w.abs().max())
        import math as _m
        self.w_scale = 2.0 ** _m.floor(_m.log2(512.0 / amax)) if amax > 0 and _m.isfinite(amax) else 1.0
        self.img = torch.empty(L.spr_gemm_w_image_bytes(self.N, self.K), dtype=torch.uint8, device=w.device)
        rc = L.spr_gemm_prepare_weight(w.data_ptr(), self.N, self.K, self.w_scale, self.img.data_ptr(), _stream())
        _lib.check(rc, "spr_gemm_prepare_weight")
        self.key = (weight.data_ptr(), weight._version)


def weight_image(weight: torch.Tensor) -> WeightImage:
    """Operand image of a weight, cached ON the tensor object (so it dies with it: an id()-keyed dict would hand a
    stale image to a new parameter that happens to reuse the id, address and version of a freed one) and rebuilt when
    the storage or the version counter changes (load_state_dict, .to(), optimiser steps)."""
    key = (weight.data_ptr(), weight._version)
    wi = getattr(weight, "_spr_weight_image", None)
    if wi is None or wi.key != key:
        wi = WeightImage(weight)
        weight._spr_weight_image = wi
    return wi


def gemm_prepare_input(x: torch.Tensor, img: Optional[torch.Tensor] = None) -> torch.Tensor:
    L = _lib.lib()
    xx = _f32c(x, "x")
    T, K = xx.shape
    if img is None:
        img = gemm_a_image(T, K, xx.device)
    rc = L.spr_gemm_prepare_input(xx.data_ptr(), T, K, K, A_SCALE, img.data_ptr(), _stream())
    _lib.check(rc, "spr_gemm_prepare_input")
    return img


def layernorm256_prepare(x, gamma, beta, pos, eps: float, img: Optional[torch.Tensor], out_f32: bool = False):
    """LayerNorm over 256 channels (+ pos) -> A image and/or fp32 rows."""
    L = _lib.lib()
    xx = _f32c(x, "x")
    T, d = xx.shape
    if d != 256:
        raise RuntimeError("layernorm256_prepare: d_model must be 256")
    out = torch.empty_like(xx) if out_f32 else None
    rc = L.spr_layernorm256_prepare(xx.data_ptr(), _ptr(gamma), _ptr(beta), _ptr(pos), T, float(eps), A_SCALE,
                                    _ptr(img), _ptr(out), _stream())
    _lib.check(rc, "spr_layernorm256_prepare")
    return out


def gemm_tc(a_img: torch.Tensor, wi: WeightImage, bias, T: int, mode: int = OUT_F32, residual=None, relu: bool = False,
            out=None, out_lo=None, n_scaled: int = 0, col_scale: float = 1.0, stats16=None):
    """Y = act(X W^T + b) (+ residual) on the tcgen05 tensor cores from operand images; see include/spr_b200.h."""
    L = _lib.lib()
    N, K = wi.N, wi.K
    dev = a_img.device
    ld_out = N
    if mode == OUT_F32:
        if out is None:
            out = torch.empty((T, N), dtype=torch.float32, device=dev)
    elif mode == OUT_PLANES:
        if out is None:
            out = torch.empty((T, N), dtype=torch.float16, device=dev)
            out_lo = torch.empty_like(out)
    else:
        if out is None:
            out = gemm_a_image(T, N, dev)
    rc = L.spr_gemm_tc(a_img.data_ptr(), wi.img.data_ptr(), _ptr(bias), _ptr(residual),
                       residual.shape[1] if residual is not None else 0, int(T), N, K, 1.0 / (A_SCALE * wi.w_scale),
                       1 if relu else 0, int(mode), out.data_ptr(), _ptr(out_lo), ld_out, int(n_scaled),
                       float(col_scale), A_SCALE, _ptr(stats16), _stream())
    _lib.check(rc, "spr_gemm_tc")
    return (out, out_lo) if mode == OUT_PLANES else out


class CrossEncoderTable:
    """Host pointer / scalar tables of spr_cross_encoder_forward for a stack of encoder layers, rebuilt when any of the
    parameters changes (storage or version counter)."""

    def __init__(self, layers):
        import ctypes
        self.key = self.signature(layers)
        ptrs, scal, self.keep = [], [], []
        for layer in layers:
            sa, ca = layer.self_attn, layer.multihead_attn
            wis = [weight_image(sa.in_proj_weight), weight_image(sa.out_proj.weight), weight_image(ca.in_proj_weight),
                   weight_image(ca.out_proj.weight), weight_image(layer.linear1.weight), weight_image(layer.linear2.weight)]
            self.keep += wis
            for n in (layer.norm1, layer.norm2, layer.norm3):
                ptrs += [_ptr(n.weight), _ptr(n.bias)]
            ptrs += [wis[0].img.data_ptr(), _ptr(sa.in_proj_bias), wis[1].img.data_ptr(), _ptr(sa.out_proj.bias),
                     wis[2].img.data_ptr(), _ptr(ca.in_proj_bias), wis[3].img.data_ptr(), _ptr(ca.out_proj.bias),
                     wis[4].img.data_ptr(), _ptr(layer.linear1.bias), wis[5].img.data_ptr(), _ptr(layer.linear2.bias)]
            scal += [layer.norm1.eps, layer.norm2.eps, layer.norm3.eps] + [w.w_scale for w in wis]
        self.ptrs = (ctypes.c_void_p * len(ptrs))(*ptrs)
        self.scal = (ctypes.c_float * len(scal))(*scal)
        self.n_layers = len(layers)
        self.d_ff = layers[0].linear1.out_features

    @staticmethod
    def signature(layers):
        sig = []
        for layer in layers:
            for t in (layer.norm1.weight, layer.norm1.bias, layer.norm2.weight, layer.norm2.bias, layer.norm3.weight,
                      layer.norm3.bias, layer.self_attn.in_proj_weight, layer.self_attn.in_proj_bias,
                      layer.self_attn.out_proj.weight, layer.self_attn.out_proj.bias, layer.multihead_attn.in_proj_weight,
                      layer.multihead_attn.in_proj_bias, layer.multihead_attn.out_proj.weight,
                      layer.multihead_attn.out_proj.bias, layer.linear1.weight, layer.linear1.bias, layer.linear2.weight,
                      layer.linear2.bias):
                sig.append((0, 0) if t is None else (t.data_ptr(), t._version))
        return tuple(sig)


def cross_encoder_forward(x: torch.Tensor, pos: Optional[torch.Tensor], table: CrossEncoderTable, n_heads: int,
                          sa_tiles: torch.Tensor, ca_tiles: torch.Tensor, final_norm=None) -> torch.Tensor:
    """All layers of the packed cross-encoder (+ the final LayerNorm) from ONE library call (csrc/encoder_seq.cu): the
    launches of TransformerCrossEncoderLayer.forward_fused in the same order, without the interpreter between them.
    x [T, 256] is not modified (a copy is updated in place)."""
    import ctypes
    L = _lib.lib()
    xx = _f32c(x, "x").clone()
    T, d = xx.shape
    pp = None if pos is None else _f32c(pos, "pos")
    dev = xx.device
    img = gemm_a_image(T, d, dev)
    img_ffn = gemm_a_image(T, table.d_ff, dev)
    hi = torch.empty((T, 3 * d), dtype=torch.float16, device=dev)
    lo = torch.empty_like(hi)
    out = torch.empty_like(xx) if final_norm is not None else None
    rc = L.spr_cross_encoder_forward(
        xx.data_ptr(), _ptr(pp), T, d, int(n_heads), table.d_ff, table.n_layers,
        ctypes.cast(table.ptrs, ctypes.c_void_p), ctypes.cast(table.scal, ctypes.c_void_p), sa_tiles.data_ptr(),
        sa_tiles.shape[0], ca_tiles.data_ptr(), ca_tiles.shape[0], img.data_ptr(), img_ffn.data_ptr(), hi.data_ptr(),
        lo.data_ptr(), A_SCALE, attention_generation(),
        None if final_norm is None else final_norm.weight.data_ptr(),
        None if final_norm is None else final_norm.bias.data_ptr(),
        0.0 if final_norm is None else float(final_norm.eps), _ptr(out), _stream())
    _lib.check(rc, "spr_cross_encoder_forward")
    return xx if out is None else out


class _EncoderGraph:
    """One captured replay of spr_cross_encoder_forward for a shape bucket (rows padded to a multiple of 64, tile lists
    to a multiple of TILE_PAD): static input buffers, the launches of all layers as ONE graph launch."""

    def __init__(self, table, n_heads, t_pad, n_sa, n_ca, has_pos, final_norm, device):
        self.x = torch.zeros((t_pad, 256), dtype=torch.float32, device=device)
        self.pos = torch.zeros((t_pad, 256), dtype=torch.float32, device=device) if has_pos else None
        self.sa = torch.zeros((n_sa, 4), dtype=torch.int32, device=device)   # all-padding lists: every CTA returns at once
        self.ca = torch.zeros((n_ca, 4), dtype=torch.int32, device=device)
        side = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):   # eager once: per-kernel attributes are set outside the capture
            cross_encoder_forward(self.x, self.pos, table, n_heads, self.sa, self.ca, final_norm)
        torch.cuda.current_stream(device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        # thread-local error mode: other threads of the process (NCCL's watchdog, a data loader) may call into CUDA
        # while this thread captures
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            self.out = cross_encoder_forward(self.x, self.pos, table, n_heads, self.sa, self.ca, final_norm)

    def run(self, x, pos, sa_tiles, ca_tiles):
        t = x.shape[0]
        self.x[:t].copy_(x)            # rows past t stay zero: LayerNorm of a zero row is finite, nobody reads the result
        if self.pos is not None:
            self.pos[:t].copy_(pos)
        self.sa.copy_(sa_tiles)
        self.ca.copy_(ca_tiles)
        self.graph.replay()
        return self.out[:t].clone()    # the static output is overwritten by the next replay


ENCODER_GRAPH_MAX_ROWS = 16384   # above this the forward is bound by the device, not by launch overhead
ENCODER_GRAPH_TILE_PAD = 8
ENCODER_GRAPH_CACHE = 12


def cross_encoder_forward_graphed(x, pos, table: CrossEncoderTable, n_heads: int, sa_tiles, ca_tiles, final_norm=None):
    """cross_encoder_forward through a CUDA graph per shape bucket (token rows rounded up to 64, tile lists to 8 entries):
    ~70 kernel launches become one graph launch -- the forward of a single pair is bound by launch overhead
    (tools/host_profile.py).  Bit-identical to the eager call; the tile lists must be padded (attention_tiles(...,
    pad_multiple=ENCODER_GRAPH_TILE_PAD))."""
    t = x.shape[0]
    t_pad = (t + 63) // 64 * 64
    key = (t_pad, sa_tiles.shape[0], ca_tiles.shape[0], pos is not None, attention_generation(), str(x.device),
           None if final_norm is None else (final_norm.weight.data_ptr(), final_norm.bias.data_ptr(), float(final_norm.eps)))
    graphs = table.__dict__.setdefault("graphs", {})
    g = graphs.pop(key, None)
    if g is None:
        if len(graphs) >= ENCODER_GRAPH_CACHE:
            graphs.pop(next(iter(graphs)))   # least recently used
        g = _EncoderGraph(table, n_heads, t_pad, sa_tiles.shape[0], ca_tiles.shape[0], pos is not None, final_norm, x.device)
    graphs[key] = g                          # most recently used last
    return g.run(_f32c(x, "x"), None if pos is None else _f32c(pos, "pos"), sa_tiles, ca_tiles)


def linear_tc(x: torch.Tensor, weight: torch.Tensor, bias=None, relu: bool = False, residual=None) -> torch.Tensor:
    """Drop-in for F.linear on fp32 rows: builds the operand image, then one tensor-core GEMM."""
    img = gemm_prepare_input(x)
    return gemm_tc(img, weight_image(weight), bias, x.shape[0], OUT_F32, residual=residual, relu=relu)
