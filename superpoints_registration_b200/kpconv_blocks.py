"""KPConv operator and the encoder blocks built on it, on B200.

Mirrors the operator surface of the reference's models/backbone_kpconv/kpconv_blocks.py:
  KPConv                 :175-420   -> one fused kernel (spr_kpconv_forward)
  BatchNormBlock         :474-530   -> per-cloud instance norm, fused with the following LeakyReLU / residual
  UnaryBlock             :533-567
  SimpleBlock            :590-646
  ResnetBottleneckBlock  :649-741
  max_pool               :127-143   -> spr_max_pool
  block_decider          :429-471
Module / parameter names are the reference's, so a reference state_dict loads key for key
(`KPConv.weights`, `KPConv.kernel_points`, `unary1.mlp.weight`, `unary2.mlp.weight`,
`unary_shortcut.mlp.weight`).  The Linear layers of the unary blocks run on the tcgen05 split-precision GEMM
(ops.linear_tc / ops.gemm_tc: fp16 hi/lo operand pairs, fp32 accumulation, fp32-level accuracy); nothing on the
path goes through cuBLAS.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

from . import ops
from .kernel_points import load_kernels

LRELU_SLOPE = 0.1
IN_EPS = 1e-5  # nn.InstanceNorm1d default eps


def max_pool(x, inds, order=None):
    """Pools features with the maximum over each pooling neighbourhood (shadow index -> zero row)."""
    return ops.max_pool(x, inds, order)


class KPConv(nn.Module):
    """Rigid kernel-point convolution (linear influence, sum aggregation).

    Same constructor signature as the reference (kpconv_blocks.py:177-179).  Modes no shipped configuration
    uses raise NotImplementedError.
    """

    def __init__(self, kernel_size, p_dim, in_channels, out_channels, KP_extent, radius,
                 fixed_kernel_points='center', KP_influence='linear', aggregation_mode='sum',
                 deformable=False, modulated=False):
        super().__init__()
        if deformable or modulated:
            raise NotImplementedError("deformable / modulated KPConv is not used by any shipped configuration")
        if KP_influence != 'linear':
            raise NotImplementedError(f"KP_influence='{KP_influence}': only 'linear' is on the registration path")
        if aggregation_mode != 'sum':
            raise NotImplementedError(f"aggregation_mode='{aggregation_mode}': only 'sum' is on the registration path")
        if p_dim != 3:
            raise NotImplementedError("3-D points only")
        self.K = kernel_size
        self.p_dim = p_dim
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.radius = radius
        self.KP_extent = KP_extent
        self.fixed_kernel_points = fixed_kernel_points
        self.KP_influence = KP_influence
        self.aggregation_mode = aggregation_mode
        self.deformable = deformable
        self.modulated = modulated
        self.mode = None  # None: ops.default_kpconv_mode (tcgen05 where available); 0 forces the fp32 CUDA-core contraction
        self.weights = Parameter(torch.zeros((self.K, in_channels, out_channels), dtype=torch.float32))
        nn.init.kaiming_uniform_(self.weights, a=math.sqrt(5))
        kp = load_kernels(self.radius, self.K, dimension=self.p_dim, fixed=self.fixed_kernel_points)
        self.kernel_points = Parameter(torch.tensor(kp, dtype=torch.float32), requires_grad=False)

    def forward(self, q_pts, s_pts, neighb_inds, x, order=None):
        return ops.kpconv_forward(q_pts, s_pts, neighb_inds, x, self.weights, self.kernel_points,
                                  float(self.KP_extent), mode=self.mode, order=order)

    def __repr__(self):
        return (f'KPConv(radius: {self.radius:.2f}, extent: {self.KP_extent:.2f}, in_feat: {self.in_channels:d}, '
                f'out_feat: {self.out_channels:d})')


class BatchNormBlock(nn.Module):
    """Per-cloud InstanceNorm (no affine, no running statistics) or, when `use_bn` is False, a bias."""

    def __init__(self, in_dim, use_bn, bn_momentum):
        super().__init__()
        self.bn_momentum = bn_momentum
        self.use_bn = use_bn
        self.in_dim = in_dim
        if not self.use_bn:
            self.bias = Parameter(torch.zeros(in_dim, dtype=torch.float32))

    def forward(self, x, stack_lengths, slope: float = 1.0, residual=None):
        """`slope` / `residual` let callers fuse the LeakyReLU and the residual sum that always follow."""
        # (the reference asserts x.shape[0] == stack_lengths.sum() here, :499 -- a device sync we do not pay)
        if self.use_bn:
            return ops.instance_norm_lrelu(x, stack_lengths, IN_EPS, slope, residual)
        y = x + self.bias
        if residual is not None:
            y = y + residual
        return y if slope == 1.0 else torch.nn.functional.leaky_relu(y, slope)

    def forward_ex(self, x, stack_lengths, slope: float = 1.0, residual=None, want_f32=True, want_image=False,
                   kpconv_points=None, stats16=None, kpconv_planar=False, residual_stats16=None):
        """Format-aware variant (ops.instance_norm_lrelu_ex): the consumer's operand formats come out of the
        normalisation kernel itself.  Only with instance norm (every shipped config)."""
        if not self.use_bn:
            raise NotImplementedError("format-aware outputs need use_batch_norm=True")
        return ops.instance_norm_lrelu_ex(x, stack_lengths, IN_EPS, slope, residual, want_f32=want_f32,
                                          want_image=want_image, kpconv_points=kpconv_points, stats16=stats16,
                                          kpconv_planar=kpconv_planar, residual_stats16=residual_stats16)

    def __repr__(self):
        return f'BatchNormBlock(in_feat: {self.in_dim:d}, momentum: {self.bn_momentum:.3f}, only_bias: {not self.use_bn})'


class UnaryBlock(nn.Module):
    """Linear (no bias) -> per-cloud InstanceNorm -> optional LeakyReLU(0.1)."""

    def __init__(self, in_dim, out_dim, use_bn, bn_momentum, no_relu=False):
        super().__init__()
        self.bn_momentum = bn_momentum
        self.use_bn = use_bn
        self.no_relu = no_relu
        self.in_dim = in_dim
        self.out_dim = out_dim
        self.mlp = nn.Linear(in_dim, out_dim, bias=False)
        self.batch_norm = BatchNormBlock(out_dim, self.use_bn, self.bn_momentum)

    def forward(self, x, stack_lengths=None, residual=None, slope=None):
        if slope is None:
            slope = 1.0 if self.no_relu else LRELU_SLOPE
        # the Linear runs on the tcgen05 split-precision GEMM (fp32-level accuracy); CPU tensors raise, as everywhere
        y = ops.linear_tc(x, self.mlp.weight, self.mlp.bias)
        return self.batch_norm(y, stack_lengths, slope=slope, residual=residual)

    def forward_ex(self, x_image, n_rows, stack_lengths, residual=None, slope=None, **wants):
        """Same operator from an operand image of x; `wants` selects the output formats (BatchNormBlock.forward_ex)."""
        if slope is None:
            slope = 1.0 if self.no_relu else LRELU_SLOPE
        # the GEMM epilogue also sums its rows in 16-row blocks, so the normalisation reads them once, not twice
        stats = ops.block_stats(n_rows, self.out_dim, x_image.device) if self.use_bn else None
        y = ops.gemm_tc(x_image, ops.weight_image(self.mlp.weight), self.mlp.bias, n_rows, ops.OUT_F32, stats16=stats)
        return self.batch_norm.forward_ex(y, stack_lengths, slope=slope, residual=residual, stats16=stats, **wants)

    def linear_raw(self, x_image, n_rows):
        """The Linear alone, with the 16-row block sums of its output: (y, stats16).  For a consumer that applies this
        block's normalisation itself (the projected shortcut, normalised inside unary2's apply kernel)."""
        stats = ops.block_stats(n_rows, self.out_dim, x_image.device)
        y = ops.gemm_tc(x_image, ops.weight_image(self.mlp.weight), self.mlp.bias, n_rows, ops.OUT_F32, stats16=stats)
        return y, stats

    def __repr__(self):
        return (f'UnaryBlock(in_feat: {self.in_dim:d}, out_feat: {self.out_dim:d}, BN: {self.use_bn}, '
                f'ReLU: {not self.no_relu})')


def _level_io(batch, layer_ind: int, strided: bool):
    # a Pyramid hands the kernels its int32 index matrices (batch.index); any other mapping with the reference's keys
    # (e.g. a dict of int64 tensors built by a caller) is used as is
    index = getattr(batch, 'index', None) or (lambda key, l: batch[key][l])
    lens = getattr(batch, 'lengths32', None) or batch['stack_lengths']
    if strided:
        return (batch['points'][layer_ind + 1], batch['points'][layer_ind], index('pools', layer_ind), lens[layer_ind + 1])
    return (batch['points'][layer_ind], batch['points'][layer_ind], index('neighbors', layer_ind), lens[layer_ind])


class SimpleBlock(nn.Module):
    """KPConv -> InstanceNorm -> LeakyReLU (kpconv_blocks.py:590-646)."""

    def __init__(self, block_name, in_dim, out_dim, radius, layer_ind, config):
        super().__init__()
        self.bn_momentum = config.batch_norm_momentum
        self.use_bn = config.use_batch_norm
        self.layer_ind = layer_ind
        self.block_name = block_name
        self.in_dim = in_dim
        self.out_dim = out_dim
        extent = radius * config.KP_extent / config.conv_radius
        self.KPConv = KPConv(config.num_kernel_points, config.in_points_dim, in_dim, out_dim // 2, extent, radius,
                             fixed_kernel_points=config.fixed_kernel_points, KP_influence=config.KP_influence,
                             aggregation_mode=config.aggregation_mode, deformable='deform' in block_name,
                             modulated=config.modulated)
        self.batch_norm = BatchNormBlock(out_dim // 2, self.use_bn, self.bn_momentum)

    def forward(self, x, batch):
        strided = 'strided' in self.block_name
        q_pts, s_pts, inds, lengths = _level_io(batch, self.layer_ind, strided)
        orders = getattr(batch, 'order', None)
        order = orders[self.layer_ind + 1 if strided else self.layer_ind] if orders is not None else None
        x = self.KPConv(q_pts, s_pts, inds, x, order)
        if self.use_bn and x.shape[1] % 32 == 0:
            o = self.batch_norm.forward_ex(x, lengths, slope=LRELU_SLOPE, want_f32=True, want_image=True)
            batch['_operand_image'] = (o['f32'], o['image'])   # for the next block's unary GEMMs (matched by identity)
            return o['f32']
        return self.batch_norm(x, lengths, slope=LRELU_SLOPE)


class ResnetBottleneckBlock(nn.Module):
    """unary1 -> KPConv -> norm -> lrelu -> unary2 ; (+) shortcut ; lrelu   (kpconv_blocks.py:649-741)."""

    def __init__(self, block_name, in_dim, out_dim, radius, layer_ind, config):
        super().__init__()
        self.bn_momentum = config.batch_norm_momentum
        self.use_bn = config.use_batch_norm
        self.block_name = block_name
        self.layer_ind = layer_ind
        self.in_dim = in_dim
        self.out_dim = out_dim
        mid = out_dim // 4
        extent = radius * config.KP_extent / config.conv_radius
        self.unary1 = UnaryBlock(in_dim, mid, self.use_bn, self.bn_momentum) if in_dim != mid else nn.Identity()
        self.KPConv = KPConv(config.num_kernel_points, config.in_points_dim, mid, mid, extent, radius,
                             fixed_kernel_points=config.fixed_kernel_points, KP_influence=config.KP_influence,
                             aggregation_mode=config.aggregation_mode, deformable='deform' in block_name,
                             modulated=config.modulated)
        self.batch_norm_conv = BatchNormBlock(mid, self.use_bn, self.bn_momentum)
        self.unary2 = UnaryBlock(mid, out_dim, self.use_bn, self.bn_momentum, no_relu=True)
        self.unary_shortcut = (UnaryBlock(in_dim, out_dim, self.use_bn, self.bn_momentum, no_relu=True)
                               if in_dim != out_dim else nn.Identity())

    def _fusable(self, features):
        mid = self.out_dim // 4
        return (self.use_bn and isinstance(self.unary1, UnaryBlock) and features.is_cuda and self.KPConv.mode is None
                and mid in ops._TC_CHANNELS and self.in_dim % 32 == 0 and self.out_dim % 32 == 0
                and self.out_dim <= 1024)

    def _forward_fused(self, features, batch):
        """Same arithmetic as forward(); every intermediate is written once, directly in the format its consumer
        reads: unary1's normalised output as pre-split KPConv rows, the KPConv norm as the operand image of unary2,
        the block output as rows + operand image of the next block's unary GEMMs."""
        strided = 'strided' in self.block_name
        pre_lengths = (getattr(batch, 'lengths32', None) or batch['stack_lengths'])[self.layer_ind]
        q_pts, s_pts, inds, post_lengths = _level_io(batch, self.layer_ind, strided)
        n_in = features.shape[0]
        stash = batch.get('_operand_image')
        # the stash holds the tensor itself: identity, not address, decides (a freed block output's address can be
        # handed to a later tensor)
        f_img = stash[1] if stash is not None and stash[0] is features else ops.gemm_prepare_input(features)

        planar = ops.kpconv_kernel_generation(self.out_dim // 4, inds.shape[1]) in (2, 3)
        x = self.unary1.forward_ex(f_img, n_in, pre_lengths, want_f32=False, kpconv_points=s_pts,
                                   kpconv_planar=planar)['kpconv']
        orders = getattr(batch, 'order', None)
        order = orders[self.layer_ind + 1 if strided else self.layer_ind] if orders is not None else None
        x = ops.kpconv_forward_prepared(q_pts, inds, x, self.KPConv.weights, self.KPConv.kernel_points,
                                        self.KPConv.KP_extent, order=order)
        x_img = self.batch_norm_conv.forward_ex(x, post_lengths, slope=LRELU_SLOPE, want_f32=False,
                                                want_image=True)['image']
        n_out = q_pts.shape[0]
        if strided:
            shortcut = max_pool(features, inds, order)
            s_img = ops.gemm_prepare_input(shortcut) if isinstance(self.unary_shortcut, UnaryBlock) else None
        else:
            shortcut, s_img = features, f_img
        shortcut_stats = None
        if isinstance(self.unary_shortcut, UnaryBlock):
            # the projected shortcut stays raw: its normalisation (no activation) happens inside unary2's apply kernel,
            # which saves one write and one read of an [n, out_dim] tensor per block
            shortcut, shortcut_stats = self.unary_shortcut.linear_raw(s_img, n_out)
        o = self.unary2.forward_ex(x_img, n_out, post_lengths, residual=shortcut, slope=LRELU_SLOPE, want_f32=True,
                                   want_image=True, residual_stats16=shortcut_stats)
        batch['_operand_image'] = (o['f32'], o['image'])
        return o['f32']

    def forward(self, features, batch):
        if self._fusable(features):
            return self._forward_fused(features, batch)
        batch['_operand_image'] = None  # this route writes plain rows only
        strided = 'strided' in self.block_name
        pre_lengths = (getattr(batch, 'lengths32', None) or batch['stack_lengths'])[self.layer_ind]
        q_pts, s_pts, inds, post_lengths = _level_io(batch, self.layer_ind, strided)

        x = self.unary1(features, pre_lengths) if isinstance(self.unary1, UnaryBlock) else features
        x = self.KPConv(q_pts, s_pts, inds, x)
        x = self.batch_norm_conv(x, post_lengths, slope=LRELU_SLOPE)

        shortcut = max_pool(features, inds) if strided else features
        if isinstance(self.unary_shortcut, UnaryBlock):
            shortcut = self.unary_shortcut(shortcut, post_lengths)
        # unary2 has no activation of its own: its norm, the residual sum and the block's final LeakyReLU
        # (kpconv_blocks.py:730-741) run as one kernel.
        return self.unary2(x, post_lengths, residual=shortcut, slope=LRELU_SLOPE)


def block_decider(block_name, radius, in_dim, out_dim, layer_ind, config):
    if block_name == 'unary':
        return UnaryBlock(in_dim, out_dim, config.use_batch_norm, config.batch_norm_momentum)
    if block_name in ('simple', 'simple_strided'):
        return SimpleBlock(block_name, in_dim, out_dim, radius, layer_ind, config)
    if block_name in ('resnetb', 'resnetb_strided'):
        return ResnetBottleneckBlock(block_name, in_dim, out_dim, radius, layer_ind, config)
    if any(t in block_name for t in ('deformable', 'invariant', 'equivariant', 'max_pool', 'global_average',
                                     'nearest_upsample', 'unary2')):
        raise NotImplementedError(f"block '{block_name}' is not used by the registration encoder")
    raise ValueError('Unknown block name in the architecture definition : ' + block_name)
