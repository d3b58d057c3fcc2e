"""Drop-in replacements for the reference's hot-path Python seams (models/mvs4net_utils.py).

  * ``stagenet``                  - same constructor and ``forward`` signature / return dict as mvs4net_utils.py:1017-1162;
                                    steps 1-2 run as the fused CUDA kernel, steps 3-4 as the fused tail kernel.
                                    It holds no parameters, so ``model.stagenet = stagenet(...)`` on a reference
                                    ``MVS4net`` keeps checkpoints loadable.
  * ``homo_warping``              - mvs4net_utils.py:21 (compatibility only; materialises the warped volume)
  * ``init_inverse_range``        - mvs4net_utils.py:79
  * ``schedule_inverse_range``    - mvs4net_utils.py:87
  * ``depth_regression``          - models/module.py:935
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import ops
from .epipolar import EpipolarAggregate, epipolar_aggregate_variant


class _Tail(torch.autograd.Function):
    """softmax over D + arg-max gather (or regression) + confidence + inverse range, one kernel; differentiable
    w.r.t. the logits through ``attn_weight`` (and ``depth`` in regression mode)."""

    @staticmethod
    def forward(ctx, logits, hypo, split_itv, want_conf, inverse_depth, depth_mode):
        attn, depth, conf, inv_min, inv_max = ops.tail(logits, hypo, split_itv, want_conf, inverse_depth, depth_mode)
        if ctx.needs_input_grad[0]:
            ctx.save_for_backward(attn, hypo.detach().float().contiguous(), depth)
            ctx.depth_mode = depth_mode
        ctx.mark_non_differentiable(*[t for t in (conf, inv_min, inv_max) if t is not None])
        if depth_mode == ops.DEPTH_ARGMAX:
            ctx.mark_non_differentiable(depth)
        return attn, depth, conf, inv_min, inv_max

    @staticmethod
    def backward(ctx, g_attn, g_depth, g_conf, g_min, g_max):
        attn, hypo, depth = ctx.saved_tensors
        g = ops.tail_bwd(attn, hypo, depth, g_attn, g_depth if ctx.depth_mode == ops.DEPTH_REGRESS else None,
                         ctx.depth_mode)
        return g, None, None, None, None, None


class stagenet(nn.Module):
    """B200-native ``stagenet`` (reference models/mvs4net_utils.py:1017).

    Extra keyword-only options (defaults reproduce the reference): ``feature_dtype`` (``torch.bfloat16`` stores the
    features the kernel gathers in bf16, fp32 accumulate) and ``depth_mode`` ("argmax" | "regress").
    """

    def __init__(self, inverse_depth=False, mono=False, attn_fuse_d=True, vis_ETA=False, attn_temp=1, debug=0, *,
                 feature_dtype: Optional[torch.dtype] = None, depth_mode: str = "argmax"):
        super(stagenet, self).__init__()
        self.inverse_depth = inverse_depth
        self.mono = mono
        self.attn_fuse_d = attn_fuse_d
        self.vis_ETA = vis_ETA
        self.attn_temp = attn_temp
        self.debug = debug
        self.feature_dtype = feature_dtype
        if depth_mode not in ("argmax", "regress"):
            raise ValueError("depth_mode must be 'argmax' or 'regress'")
        self.depth_mode = ops.DEPTH_ARGMAX if depth_mode == "argmax" else ops.DEPTH_REGRESS

    def forward(self, features, proj_matrices, depth_hypo, regnet, stage_idx, group_cor=False, group_cor_dim=8,
                split_itv=1, fn=None):
        if self.vis_ETA:
            raise NotImplementedError("stagenet(B200): vis_ETA .npy dumps are a debugging aid of the reference and "
                                      "are not produced; use epipolar.epipolar_weights() for the attention weights")
        ref_feature = features[0]
        # steps 1-2: fused homography warp + group correlation + epipolar attention + view aggregation
        if group_cor and self.attn_fuse_d:
            cor_feats = EpipolarAggregate.apply(ref_feature, depth_hypo, proj_matrices, int(group_cor_dim),
                                                float(self.attn_temp), self.feature_dtype, *features[1:])
        else:  # variance cost (:1071) and / or per-pixel weight (:1078-1081): fused forward + backward kernels, fp32
            cor_feats = epipolar_aggregate_variant(features, proj_matrices, depth_hypo, bool(group_cor), int(group_cor_dim),
                                                   bool(self.attn_fuse_d), float(self.attn_temp))
        # step 3: regularisation (unchanged, cuDNN)
        attn_weight = regnet(cor_feats)  # B D H W
        del cor_feats
        # step 4 + confidence + next-stage inverse range: fused tail
        attn_weight, depth, conf, inv_min, inv_max = _Tail.apply(
            attn_weight, depth_hypo, float(split_itv), not self.training, bool(self.inverse_depth), self.depth_mode)
        if self.training:
            conf = torch.tensor(0.0, dtype=torch.float32, device=ref_feature.device, requires_grad=False)
        ret_dict = {"depth": depth, "photometric_confidence": conf, "hypo_depth": depth_hypo,
                    "attn_weight": attn_weight}
        if self.inverse_depth:
            ret_dict["inverse_min_depth"] = inv_min
            ret_dict["inverse_max_depth"] = inv_max
        if self.mono:
            ret_dict["mono_feat"] = ref_feature
        return ret_dict

    def forward_fused_regnet(self, features, proj_matrices, depth_hypo, regnet, stage_idx, group_cor=False,
                             group_cor_dim=8, split_itv=1):
        """Inference-only variant of ``forward`` for a ``network.reg2d`` regulariser: the last layers of ``regnet``
        (transposed conv + BN + ReLU, skip add, ``prob``) and steps 3-4 run as ONE kernel (``ops.regtail``), so the
        logits are never materialised.  Same return dictionary as ``forward``."""
        if self.training or torch.is_grad_enabled():
            raise RuntimeError("stagenet.forward_fused_regnet is inference-only (eval mode under torch.no_grad())")
        ref_feature = features[0]
        if group_cor and self.attn_fuse_d:
            cor_feats = EpipolarAggregate.apply(ref_feature, depth_hypo, proj_matrices, int(group_cor_dim),
                                                float(self.attn_temp), self.feature_dtype, *features[1:])
        else:
            cor_feats = epipolar_aggregate_variant(features, proj_matrices, depth_hypo, bool(group_cor), int(group_cor_dim),
                                                   bool(self.attn_fuse_d), float(self.attn_temp))
        attn_weight, depth, conf, inv_min, inv_max = regnet.forward_fused_tail(
            cor_feats, depth_hypo, float(split_itv), bool(self.inverse_depth), self.depth_mode)
        ret_dict = {"depth": depth, "photometric_confidence": conf, "hypo_depth": depth_hypo,
                    "attn_weight": attn_weight}
        if self.inverse_depth:
            ret_dict["inverse_min_depth"] = inv_min
            ret_dict["inverse_max_depth"] = inv_max
        if self.mono:
            ret_dict["mono_feat"] = ref_feature
        return ret_dict


FusedStageNet = stagenet


def homo_warping(src_fea, src_proj, ref_proj, depth_values, vis_ETA=False, fn=None):
    """Reference signature (mvs4net_utils.py:21): ``src_fea`` [B,C,Hs,Ws], ``src_proj``/``ref_proj`` [B,4,4],
    ``depth_values`` [B,D,H,W] -> warped [B,C,D,H,W].  Forward only (the fused op carries the autograd path)."""
    if vis_ETA:
        raise NotImplementedError("homo_warping(B200): vis_ETA dumps are not produced")
    if src_fea.dim() != 4:
        raise NotImplementedError("homo_warping(B200): only 4-D src_fea is supported (the 5-D branch, "
                                  "mvs4net_utils.py:61-65, is unused by stagenet)")
    if depth_values.dim() != 4:
        raise RuntimeError("depth_values must be [B,D,H,W]")
    with torch.no_grad():
        rt = ops.compose_homography_pair(src_proj, ref_proj)
        return ops.homo_warp(ops.to_nhwc(src_fea), rt, depth_values)


def init_inverse_range(cur_depth, ndepths, device=None, dtype=None, H=None, W=None):
    """mvs4net_utils.py:79-85; ``device``/``dtype`` are accepted for signature compatibility (output is fp32 on
    ``cur_depth``'s device)."""
    return ops.init_inverse_range(cur_depth, int(ndepths), int(H), int(W))


def schedule_inverse_range(inverse_min_depth, inverse_max_depth, ndepths, H, W):
    """mvs4net_utils.py:87-94."""
    return ops.schedule_inverse_range(inverse_min_depth, inverse_max_depth, int(ndepths), int(H), int(W))


def depth_regression(p, depth_values):
    """models/module.py:935-941: ``sum_d p[d] * depth_values[d]``; ``p`` are probabilities, so the tail kernel is fed
    ``log p`` (softmax(log p) = p for a normalised p)."""
    if depth_values.dim() <= 2:
        depth_values = depth_values.view(*depth_values.shape, 1, 1).expand(-1, -1, *p.shape[2:])
    _, depth, _, _, _ = ops.tail(torch.log(p), depth_values, 0.0, False, False, ops.DEPTH_REGRESS)
    return depth
