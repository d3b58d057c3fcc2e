"""Training losses on the fused Sinkhorn kernel (K3) - drop-ins for the reference's loss functions.

``sinkhorn`` and ``MVS4net_loss`` keep the call signatures, return values and keyword names of the reference
(models/mvs4net_utils.py:1164-1210, models/MVS4Net.py:195-240), including its quirks: the transport cost enters the
exponent with a POSITIVE sign (``D_map/eps``), and the mean over an empty mask is NaN.  The optimal-transport loss of a
stage (2*iters logsumexp passes over [B,HW,D,D] tensors in the reference, plus autograd through all of them) is ONE
kernel here that returns the loss and its gradient w.r.t. ``attn_weight``.  No CPU fallback.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch

from . import ops


class SinkhornLoss(torch.autograd.Function):
    """stats, T_map = SinkhornLoss.apply(gt_depth, hypo_depth, attn_weight, mask, iters, eps, continuous, inverse, want_tmap)

    ``stats`` = (mean OT loss over masked pixels, masked-pixel count, range_err_ratio); only ``stats[0]`` is
    differentiable, and only w.r.t. ``attn_weight`` (``hypo_depth`` is detached upstream, models/MVS4Net.py:116).
    ``T_map`` is the transport map when ``want_tmap`` (else an empty tensor), never differentiated."""

    @staticmethod
    def forward(ctx, gt_depth, hypo_depth, attn_weight, mask, iters, eps, continuous, inverse_depth, want_tmap=False):
        # the in-kernel reverse sweep (shared-memory history, a [B,D,H,W] gradient buffer) only when autograd will
        # actually ask for it: needs_input_grad is all-False under torch.no_grad() (validation passes on outputs that
        # still carry requires_grad=True), unlike attn_weight.requires_grad
        need = bool(ctx.needs_input_grad[2])
        ctx.has_grad = need
        stats, grad_px, tmap = ops.sinkhorn_fwd(gt_depth, hypo_depth, attn_weight, mask, iters, eps, continuous,
                                                inverse_depth, want_grad=need, want_tmap=want_tmap)
        if need:
            ctx.save_for_backward(grad_px, stats)
        if tmap is None:
            tmap = stats.new_empty(0)
        ctx.mark_non_differentiable(tmap)
        return stats, tmap

    @staticmethod
    def backward(ctx, grad_stats, _grad_tmap):
        if not ctx.has_grad:
            return None, None, None, None, None, None, None, None, None
        grad_px, stats = ctx.saved_tensors
        grad_attn = ops.sinkhorn_bwd(grad_px, stats, grad_stats.contiguous()[0:1])
        return None, None, grad_attn, None, None, None, None, None, None


def sinkhorn(gt_depth, hypo_depth, attn_weight, mask, iters, eps=1, continuous=False):
    """Reference signature (models/mvs4net_utils.py:1164): returns ``(T_map [B,HW,D,D(+1)], loss)`` from ONE launch.

    ``T_map`` is returned detached (the reference never differentiates through it: every caller takes ``[1]``,
    models/MVS4Net.py:234,281); ``loss`` carries the gradient to ``attn_weight``."""
    stats, tmap = SinkhornLoss.apply(gt_depth, hypo_depth, attn_weight, mask, int(iters), float(eps), bool(continuous),
                                     False, True)
    return tmap, stats[0]


def MVS4net_loss(inputs: Dict[str, dict], depth_gt_ms: Dict[str, torch.Tensor], mask_ms: Dict[str, torch.Tensor],
                 **kwargs) -> Tuple[torch.Tensor, List[torch.Tensor], List[torch.Tensor], List[torch.Tensor]]:
    """Drop-in for the reference ``MVS4net_loss`` (models/MVS4Net.py:195-240).

    Returns ``(total_loss, stage_l1_loss, stage_ot_loss, range_err_ratio)``.  Per stage, the OT loss and the
    out-of-range statistic come from one launch of the fused kernel; the optional mono L1 term is a masked mean."""
    stage_lw = kwargs.get("stage_lw", [1, 1, 1, 1])
    l1ot_lw = kwargs.get("l1ot_lw", [0, 1])
    inverse = kwargs.get("inverse_depth", False)
    ot_iter = kwargs.get("ot_iter", 3)
    ot_eps = kwargs.get("ot_eps", 1)
    ot_continous = kwargs.get("ot_continous", False)
    mono = kwargs.get("mono", False)
    dev = mask_ms["stage1"].device
    total_loss = torch.tensor(0.0, dtype=torch.float32, device=dev)
    stage_ot_loss, stage_l1_loss, range_err_ratio = [], [], []
    for stage_idx, stage_key in enumerate([k for k in inputs.keys() if "stage" in k]):
        stage_inputs = inputs[stage_key]
        mask = mask_ms[stage_key] > 0.5
        depth_gt = depth_gt_ms[stage_key]
        if mono and stage_idx != 0:
            this_l1 = torch.nn.functional.l1_loss(stage_inputs["mono_depth"][mask], depth_gt[mask], reduction="mean")
        else:
            this_l1 = torch.tensor(0.0, dtype=torch.float32, device=dev)
        stats, _ = SinkhornLoss.apply(depth_gt, stage_inputs["hypo_depth"], stage_inputs["attn_weight"], mask,
                                      int(ot_iter), float(ot_eps), bool(ot_continous), bool(inverse))
        this_ot = stats[0]
        range_err_ratio.append(stats[2].detach())
        stage_l1_loss.append(this_l1)
        stage_ot_loss.append(this_ot)
        total_loss = total_loss + stage_lw[stage_idx] * (l1ot_lw[0] * this_l1 + l1ot_lw[1] * this_ot)
    return total_loss, stage_l1_loss, stage_ot_loss, range_err_ratio
