"""KPConv preprocessing (pyramid of subsampled clouds + neighbour / pool / upsample index matrices) and the
encoder, on B200.

Mirrors the operator surface of the reference's models/backbone_kpconv/kpconv.py:
  batch_grid_subsampling_kpconv   kpconv.py:174-214  (cpp_subsampling.subsample_batch)
  batch_neighbors_kpconv          kpconv.py:247-262  (cpp_neighbors.batch_query + truncation)
  Preprocessor                    kpconv.py:295-418  (the CPU pyramid builder north_star names as oracle)
  KPFEncoder                      kpconv.py:22-92
Same names, argument meaning, output dict keys ('points', 'neighbors', 'pools', 'upsamples',
'stack_lengths'), dtypes (int64 indices, shadow = number of support points, int32 lengths) and errors
(RuntimeError from the native layer).  Everything runs on the device the inputs live on; there is no CPU
round trip (the reference does `.cpu()` in, `.to(device)` out, kpconv.py:313-314,410-416) and no CPU fallback.

Inside, the pyramid is int32 and lazy where nobody looks: the searcher writes 4-byte indices (half the bytes the
search writes and the KPConv / max-pool kernels read back), the int64 tensors of the reference's dict are made
when a caller reads an entry, and 'upsamples' -- which the reference computes (kpconv.py:384) but the encoder-only
model never reads -- is searched on first access.  The contract is the key and its value, not eagerness.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .kpconv_blocks import block_decider


# ------------------------------------------------------------------------------------------------
# thin operator wrappers with the reference's names
# ------------------------------------------------------------------------------------------------

def batch_grid_subsampling_kpconv(points, batches_len, features=None, labels=None, sampleDl=0.1, max_p=0, verbose=0,
                                  random_grid_orient=True):
    """Grid subsampling (barycentres) of stacked clouds: -> (s_points f32[M,3], s_len i32[B]).

    Points come out per cloud in first-occurrence voxel order (the reference's order is its hash table's,
    grid_subsampling.cpp:85); coordinates are bit-identical.
    """
    if features is not None or labels is not None:
        raise NotImplementedError("feature/label subsampling is not on the registration path (kpconv.py:370)")
    if max_p != 0:
        raise NotImplementedError("max_p > 0 truncates in emission order, which differs from the reference's hash order")
    return ops.grid_subsample_batch(points, batches_len, float(sampleDl))


def batch_neighbors_kpconv(queries, supports, q_batches, s_batches, radius, max_neighbors):
    """Radius neighbours of stacked clouds: -> i64[Nq, min(max_count, max_neighbors)], shadow = Ns.

    max_neighbors <= 0 (keep everything, kpconv.py:261-262) is served with the library's maximum row width.
    """
    limit = int(max_neighbors) if max_neighbors > 0 else 128
    idx, mc = ops.radius_neighbors_batch(queries, supports, q_batches, s_batches, float(radius), limit)
    width = min(int(mc.item()), limit)
    if max_neighbors <= 0 and int(mc.item()) > limit:
        raise RuntimeError(f"neighbourhood of {int(mc.item())} points exceeds the supported row width {limit}")
    return idx[:, :max(width, 0)] if width > 0 else idx[:, :0]


# ------------------------------------------------------------------------------------------------
# pyramid
# ------------------------------------------------------------------------------------------------

def _split_levels(architecture: Sequence[str]):
    """Group the block list into pyramid levels: (has_conv_blocks, ends_with_stride) per level.

    Same grouping rule as kpconv.py:335-346: a level collects blocks until a pooling/strided block (which
    closes it with a subsampling) or the end of the encoder; 'global'/'upsample' blocks end the walk.
    """
    levels = []
    pending = 0
    blocks = []
    for name in architecture:
        if "global" in name or "upsample" in name:
            break
        blocks.append(name)
    for i, name in enumerate(blocks):
        if any(t in name for t in ("deformable",)):
            raise NotImplementedError("deformable KPConv is not used by any shipped configuration")
        strided = "pool" in name or "strided" in name
        if strided:
            levels.append((pending > 0, True))
            pending = 0
        else:
            pending += 1
            if i == len(blocks) - 1:
                levels.append((True, False))
    return levels


class _IndexList:
    """One list entry of the pyramid dict ('neighbors' / 'pools' / 'upsamples'): a sequence of index matrices that
    hands out the reference's dtype (int64) on access while the kernels keep reading the int32 matrices the searcher
    wrote (`raw(l)`).  Entries may be thunks: they are searched when first read."""

    def __init__(self, n_levels: int, dtype: torch.dtype):
        self._raw = [None] * n_levels       # int32 (or the requested index dtype) tensors, or callables producing them
        self._out = [None] * n_levels
        self._dtype = dtype

    def set(self, level: int, value) -> None:
        self._raw[level] = value
        self._out[level] = None

    def raw(self, level: int) -> torch.Tensor:
        v = self._raw[level]
        if callable(v):
            v = v()
            self._raw[level] = v
        return v

    def __len__(self):
        return len(self._raw)

    def __getitem__(self, level):
        if isinstance(level, slice):
            return [self[i] for i in range(*level.indices(len(self)))]
        if level < 0:
            level += len(self)
        if self._out[level] is None:
            r = self.raw(level)
            self._out[level] = r if r.dtype == self._dtype else r.to(self._dtype)
        return self._out[level]

    def __iter__(self):
        return (self[i] for i in range(len(self)))

    def __repr__(self):
        return f"_IndexList({[None if callable(r) or r is None else tuple(r.shape) for r in self._raw]})"


class Preprocessor(nn.Module):
    """Computes the metadata used by the KPConv encoder (drop-in for kpconv.py:295-418)."""

    def __init__(self, cfg, exact_width: bool = True, index_dtype: torch.dtype = torch.int64,
                 lazy_upsamples: bool = True, mode: str = "reference"):
        super().__init__()
        self.cfg = cfg
        # mode: "reference" = the CPU Preprocessor (kpconv.py:295-418), the oracle north_star names: the `limit` NEAREST
        # neighbours, barycentres on per-cloud lattices, matrices trimmed to the batch-wide max count.
        # "gpu_compat" = the PreprocessorGPU the reference's model instantiates (kpconv.py:421-549,
        # qk_regtr_full.py:40): the FIRST `limit` neighbours in index order (pytorch3d ball_query), voxel means on the
        # global lattice floor(p / dl) (MinkowskiEngine quantisation), `limit` columns always, int64 stack_lengths.
        # PARITY UNPINNED for gpu_compat: pytorch3d / MinkowskiEngine are neither vendored nor installed and the
        # reference calls its own result "not deterministic" (kpconv.py:219-220,422-424); tests compare against a
        # NumPy restatement of the documented behaviour (oracle/numpy_ops.py).
        if mode not in ("reference", "gpu_compat"):
            raise ValueError(f"Preprocessor mode '{mode}': expected 'reference' or 'gpu_compat'")
        self.mode = mode
        # exact_width: trim each index matrix to min(batch max_count, limit) columns like the reference
        # (one extra host sync at the end).  False keeps `limit` columns; the extra columns are all-shadow.
        self.exact_width = exact_width
        # index_dtype: what the dict entries hand out (the reference: int64, kpconv.py:396-398).  The searcher always
        # writes int32; the wide copy of an entry is made when (and only if) a caller reads it.
        self.index_dtype = index_dtype
        # lazy_upsamples: 'upsamples' (3 of the 10 searches of a 4-stage pyramid, never read by the encoder-only
        # model) are searched on first access instead of up front
        self.lazy_upsamples = lazy_upsamples

    @torch.no_grad()
    def forward(self, pts: List[torch.Tensor]) -> Dict[str, List[torch.Tensor]]:
        cfg = self.cfg
        if len(pts) == 0:
            raise RuntimeError("Preprocessor: empty point cloud list")
        device = pts[0].device
        if device.type != "cuda":
            raise RuntimeError("Preprocessor: inputs must be CUDA tensors (no CPU fallback on the B200 path)")
        limits = cfg.neighborhood_limits
        levels = _split_levels(cfg.architecture)
        n_levels = len(levels)

        points = torch.cat([p.to(torch.float32) for p in pts], dim=0).contiguous()
        host_lengths = [int(p.shape[0]) for p in pts]
        lengths = ops.to_device_async(host_lengths, torch.int32, device)
        r = float(cfg.first_subsampling_dl) * float(cfg.conv_radius)

        out_points, out_lens = [], []
        out_neighbors, out_pools, out_ups = (_IndexList(n_levels, self.index_dtype) for _ in range(3))
        out_order = []  # per level: the points in cell order (private: processing order of the KPConv kernels)
        out_host_lens = []  # per level: stack_lengths as a host list (private: saves consumers a device read)
        widths = []  # (list, level, max_count tensor, limit)
        grid = ops.CellGrid(points, lengths, r)
        compat = self.mode == "gpu_compat"
        exact = self.exact_width and not compat          # ball_query matrices are always `limit` wide
        sub_mode = "mean" if compat else "reference"
        empty_idx = lambda: torch.zeros((0, 1), dtype=torch.int32, device=device)

        def upsample_search(next_grid, q_pts, q_lens, limit):
            up_i, mc = next_grid.query(q_pts, q_lens, limit, index_dtype=torch.int32, by_index=compat)
            if exact:
                up_i = up_i[:, :min(int(mc.item()), limit)]
            return up_i

        for li, (has_conv, strided) in enumerate(levels):
            limit = int(limits[li])
            pending = None
            if strided:
                # the subsampling goes first and its sizes travel to the host while the convolution neighbours
                # (independent of them) are searched: the device is never idle waiting for the host to learn M
                dl = 2.0 * r / float(cfg.conv_radius)
                pending = ops.grid_subsample_batch_async(points, lengths, dl, sub_mode)
            if has_conv:
                conv_i, mc = grid.query(points, lengths, limit, index_dtype=torch.int32, by_index=compat)
                widths.append((out_neighbors, li, mc, limit))
            else:
                conv_i = empty_idx()
            if strided:
                pool_p, pool_b, pool_host = pending.finish()
                pool_i, mc = grid.query(pool_p, pool_b, limit, index_dtype=torch.int32, by_index=compat)
                widths.append((out_pools, li, mc, limit))
                next_grid = ops.CellGrid(pool_p, pool_b, 2.0 * r)
                if self.lazy_upsamples:
                    up_i = (lambda g=next_grid, q=points, ql=lengths, lim=limit: upsample_search(g, q, ql, lim))
                else:
                    up_i, mc = next_grid.query(points, lengths, limit, index_dtype=torch.int32, by_index=compat)
                    widths.append((out_ups, li, mc, limit))
            else:
                pool_i, up_i = empty_idx(), empty_idx()
                pool_p = torch.zeros((0, 3), dtype=torch.float32, device=device)
                pool_b = torch.zeros((0,), dtype=torch.int64, device=device)
                next_grid, pool_host = None, []
            out_points.append(points)
            out_order.append(grid.order() if grid is not None else None)
            out_neighbors.set(li, conv_i)
            out_pools.set(li, pool_i)
            out_ups.set(li, up_i)
            out_lens.append(lengths)
            out_host_lens.append(host_lengths)
            points, lengths, grid, host_lengths = pool_p, pool_b, next_grid, pool_host
            r *= 2.0

        if exact and widths:
            counts = torch.cat([w[2] for w in widths]).tolist()  # the single sync for all widths
            for (lst, li, _, limit), mc in zip(widths, counts):
                w = min(int(mc), limit)
                lst.set(li, lst.raw(li)[:, :w])
        lens32 = out_lens
        if compat:   # batched_lengths are int64 in PreprocessorGPU (kpconv.py:453,235-243); the kernels keep the int32 ones
            out_lens = [l.to(torch.int64) for l in out_lens]
        meta = Pyramid({"points": out_points, "neighbors": out_neighbors, "pools": out_pools, "upsamples": out_ups,
                        "stack_lengths": out_lens})
        meta.order, meta.host_lengths, meta.lengths32 = out_order, out_host_lens, lens32
        return meta


class PreprocessorGPU(Preprocessor):
    """Drop-in for the reference's PreprocessorGPU (kpconv.py:421-549): Preprocessor(cfg, mode="gpu_compat")."""

    def __init__(self, cfg, **kw):
        kw.setdefault("mode", "gpu_compat")
        super().__init__(cfg, **kw)


class Pyramid(dict):
    """The reference's collate dict (exactly its five keys) plus two by-products of building it, kept as attributes
    so that code iterating over the dict sees nothing new: `order[l]` = the level's points in cell order (the
    KPConv kernels walk their queries in it), `host_lengths[l]` = stack_lengths[l] as a host list.
    `index(key, l)` is what the kernels read: the int32 matrix behind entry l of 'neighbors' / 'pools' / 'upsamples'
    (`meta[key][l]` is the same matrix in the reference's int64)."""
    order = None
    host_lengths = None
    lengths32 = None

    def index(self, key: str, level: int) -> torch.Tensor:
        lst = self[key]
        return lst.raw(level) if isinstance(lst, _IndexList) else lst[level]


# ------------------------------------------------------------------------------------------------
# encoder
# ------------------------------------------------------------------------------------------------

class KPFEncoder(nn.Module):
    """KPConv encoder (drop-in for kpconv.py:22-92): same block plan, same state_dict keys
    (`encoder_blocks.{i}.KPConv.weights` ...), every block running on the fused sm_100a kernels."""

    def __init__(self, config, d_bottle=None, increase_channel_when_downsample: bool = True):
        super().__init__()
        radius = config.first_subsampling_dl * config.conv_radius
        in_dim, out_dim = config.in_feats_dim, config.first_feats_dim
        level = 0
        self.encoder_blocks = nn.ModuleList()
        self.encoder_skips: List[int] = []
        self.encoder_skip_dims: List[int] = []
        last_name, last_index = None, -1
        for index, name in enumerate(config.architecture):
            last_name, last_index = name, index
            if "equivariant" in name and out_dim % 3 != 0:
                raise ValueError("Equivariant block but features dimension is not a factor of 3")
            changes_level = any(t in name for t in ("pool", "strided", "upsample", "global"))
            if changes_level:
                self.encoder_skips.append(index)
                self.encoder_skip_dims.append(in_dim)
            if "upsample" in name:
                break
            self.encoder_blocks.append(block_decider(name, radius, in_dim, out_dim, level, config))
            in_dim = out_dim // 2 if "simple" in name else out_dim
            if "pool" in name or "strided" in name:
                level += 1
                radius *= 2
                if increase_channel_when_downsample:
                    out_dim *= 2
        if last_name is not None and "upsample" not in last_name:
            # encoder-only network: record the final feature width (kpconv.py:74-79)
            self.encoder_skips.append(last_index)
            self.encoder_skip_dims.append(in_dim)

    def forward(self, x, batch):
        skips = []
        for index, block in enumerate(self.encoder_blocks):
            if index in self.encoder_skips:
                skips.append(x)
            x = block(x, batch)
        return x, skips
