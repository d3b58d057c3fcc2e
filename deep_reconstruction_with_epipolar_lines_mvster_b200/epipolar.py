"""``EpipolarAggregate``: the fused cost-volume op as a ``torch.autograd.Function``.

Replaces ``stagenet.forward`` steps 1-2 of the reference (models/mvs4net_utils.py:1030-1102, including
``homo_warping`` :21-67) with one CUDA kernel forward and one backward (libmvster_b200.so).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from . import ops


class EpipolarAggregate(torch.autograd.Function):
    """``forward(ref, hypo, proj, groups, attn_temp, feature_dtype, *srcs) -> cor_feats [B,G,D,H,W]`` (fp32).

    ``ref`` / ``srcs`` are ``[B,C,H,W]``-shaped feature tensors (NCHW or channels_last strides, fp32 or bf16),
    ``hypo`` ``[B,D,H,W]``, ``proj`` ``[B,N,2,4,4]``.  Gradients flow to ``ref`` and ``srcs`` only, exactly as in
    the reference where the sampling grid is built under ``no_grad`` (mvs4net_utils.py:31).
    Only the inputs, the output and the [B,D,H,W] weight sums are saved for backward.
    """

    @staticmethod
    def forward(ctx, ref, hypo, proj, groups, attn_temp, feature_dtype, *srcs):
        if len(srcs) < 1:
            raise RuntimeError("EpipolarAggregate needs at least one source view")
        if len(srcs) > ops.MAX_SRC_VIEWS:
            raise RuntimeError("at most %d source views are supported per call, got %d" % (ops.MAX_SRC_VIEWS, len(srcs)))
        c = ref.shape[1]
        if c % groups != 0:
            raise RuntimeError("C=%d is not divisible by group_cor_dim=%d" % (c, groups))
        ref_n = ops.to_nhwc(ref, feature_dtype)
        srcs_n = [ops.to_nhwc(s, feature_dtype) for s in srcs]
        rt = ops.compose_homographies(proj)
        needs_grad = any(ctx.needs_input_grad[i] for i in [0] + list(range(6, 6 + len(srcs))))
        out, wsum, _ = ops.epi_fwd(ref_n, srcs_n, rt, hypo, groups, attn_temp, want_wsum=needs_grad)
        if needs_grad:
            ctx.save_for_backward(ref_n, rt, hypo.detach().float().contiguous(), out, wsum, *srcs_n)
            ctx.groups = groups
            ctx.attn_temp = attn_temp
            ctx.in_dtypes = [ref.dtype] + [s.dtype for s in srcs]
        return out

    @staticmethod
    def backward(ctx, gout):
        ref_n, rt, hypo, out, wsum, *srcs_n = ctx.saved_tensors
        grad_ref, grad_srcs = ops.epi_bwd(ref_n, srcs_n, rt, hypo, out, wsum, gout, ctx.groups, ctx.attn_temp)
        # gradients are returned NCHW-shaped with channels_last strides (zero-copy views of the NHWC buffers)
        gr = grad_ref.permute(0, 3, 1, 2).to(ctx.in_dtypes[0]) if ctx.needs_input_grad[0] else None
        gs = [g.permute(0, 3, 1, 2).to(dt) if ctx.needs_input_grad[6 + i] else None
              for i, (g, dt) in enumerate(zip(grad_srcs, ctx.in_dtypes[1:]))]
        return (gr, None, None, None, None, None, *gs)


def epipolar_aggregate(features: Sequence[torch.Tensor], proj_matrices: torch.Tensor, depth_hypo: torch.Tensor,
                       group_cor_dim: int, attn_temp: float, feature_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Functional form: ``features`` = [ref, src_1, ...] as the reference passes them to ``stagenet.forward``."""
    return EpipolarAggregate.apply(features[0], depth_hypo, proj_matrices, int(group_cor_dim), float(attn_temp),
                                   feature_dtype, *features[1:])


class EpipolarAggregateVariant(torch.autograd.Function):
    """``EpipolarAggregate`` for the reference's ``group_cor=False`` (variance cost, mvs4net_utils.py:1071) and / or
    ``attn_fuse_d=False`` (per-pixel weight, :1078-1081,1098) options: ``mvster_epi_fwd_mode_ex`` forward,
    ``mvster_epi_bwd_mode`` backward, fp32 features.  Saves the inputs, the volume and the weight sums only."""

    @staticmethod
    def forward(ctx, ref, hypo, proj, groups, attn_temp, group_cor, attn_fuse_d, *srcs):
        ref_n = ops.to_nhwc(ref, torch.float32)
        srcs_n = [ops.to_nhwc(s, torch.float32) for s in srcs]
        rt = ops.compose_homographies(proj)
        needs_grad = any(ctx.needs_input_grad[i] for i in [0] + list(range(7, 7 + len(srcs))))
        res = ops.epi_fwd_mode(ref_n, srcs_n, rt, hypo, groups, attn_temp, group_cor, attn_fuse_d, want_wsum=needs_grad)
        if not needs_grad:
            return res
        out, wsum = res
        ctx.save_for_backward(ref_n, rt, hypo.detach().float().contiguous(), out, wsum, *srcs_n)
        ctx.cfg = (groups, attn_temp, group_cor, attn_fuse_d)
        ctx.in_dtypes = [ref.dtype] + [s.dtype for s in srcs]
        return out

    @staticmethod
    def backward(ctx, gout):
        ref_n, rt, hypo, out, wsum, *srcs_n = ctx.saved_tensors
        groups, attn_temp, group_cor, attn_fuse_d = ctx.cfg
        grad_ref, grad_srcs = ops.epi_bwd_mode(ref_n, srcs_n, rt, hypo, out, wsum, gout, groups, attn_temp, group_cor,
                                               attn_fuse_d)
        gr = grad_ref.permute(0, 3, 1, 2).to(ctx.in_dtypes[0]) if ctx.needs_input_grad[0] else None
        gs = [g.permute(0, 3, 1, 2).to(dt) if ctx.needs_input_grad[7 + i] else None
              for i, (g, dt) in enumerate(zip(grad_srcs, ctx.in_dtypes[1:]))]
        return (gr, None, None, None, None, None, None, *gs)


def epipolar_aggregate_variant(features: Sequence[torch.Tensor], proj_matrices: torch.Tensor, depth_hypo: torch.Tensor,
                               group_cor: bool, group_cor_dim: int, attn_fuse_d: bool, attn_temp: float) -> torch.Tensor:
    """The reference's ``group_cor=False`` (variance cost, mvs4net_utils.py:1071) and/or ``attn_fuse_d=False``
    (per-pixel weight, :1078-1081) options, differentiable with respect to the features like the default path."""
    if len(features) < 2:
        raise RuntimeError("epipolar_aggregate_variant needs at least one source view")
    return EpipolarAggregateVariant.apply(features[0], depth_hypo, proj_matrices, int(group_cor_dim), float(attn_temp),
                                          bool(group_cor), bool(attn_fuse_d), *features[1:])


def epipolar_weights(features: Sequence[torch.Tensor], proj_matrices: torch.Tensor, depth_hypo: torch.Tensor,
                     group_cor_dim: int, attn_temp: float, feature_dtype: Optional[torch.dtype] = None):
    """Volume plus the per-view attention weights ``[B,N-1,D,H,W]`` (the reference's ``cor_weight``,
    mvs4net_utils.py:1083; what its debug bit 6 displays).  No autograd."""
    with torch.no_grad():
        ref_n = ops.to_nhwc(features[0], feature_dtype)
        srcs_n = [ops.to_nhwc(s, feature_dtype) for s in features[1:]]
        rt = ops.compose_homographies(proj_matrices)
        out, _, weights = ops.epi_fwd(ref_n, srcs_n, rt, depth_hypo, int(group_cor_dim), float(attn_temp),
                                      want_weights=True)
    return out, weights
