"""Callers of the hot path, widened per SURVEY.md §8f: a checkpoint-compatible ``MVS4net`` whose stage loop runs on the
fused kernels of this library.

The reference network (models/MVS4Net.py:16-199, models/mvs4net_utils.py:426-509 ``FPN4``, :884-926 ``reg2d``) is
re-stated here with the *same module / parameter names* so that ``load_state_dict`` of a reference checkpoint works
unchanged.  What differs is how it executes on a B200:

  * feature extraction runs ONCE on all N views stacked along the batch (the reference loops over views,
    MVS4Net.py:78-80) and in ``channels_last``, so the NHWC feature maps the fused K1 kernel gathers from are produced
    directly by the convolutions - no layout conversion at the K1 boundary;
  * the per-stage hot path (hypothesis schedule, homography composition, warp + correlation + epipolar attention,
    tail) is the CUDA library behind ``stagenet``;
  * in eval mode ``reg2d`` hands its last three layers (transposed conv ``conv11`` + BN + ReLU, skip add, ``prob``) to
    one fused CUDA kernel together with the stagenet tail (``mvster_regtail``), so the full
    resolution ``[B,8,D,H,W]`` activation and the logits never travel through HBM.

FPN4 and the inner layers of reg2d stay dense cuDNN convolutions (library code, as in the reference).
Only the shipped configuration family is supported: ``arch_mode='fpn'``, ``reg_net='reg2d'``, no DCN / ASFF / mono
decoder / positional encoding; anything else raises ``NotImplementedError`` instead of silently diverging.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .stagenet import init_inverse_range, schedule_inverse_range, stagenet


# ----------------------------------------------------------------------------------------------------------------------
# building blocks (parameter names follow the reference so that checkpoints load)
# ----------------------------------------------------------------------------------------------------------------------
class _BatchNormReLUTrain(torch.autograd.Function):
    """``relu?(batch_norm(x))`` in training mode on ``ops.bn_train_fwd`` / ``ops.bn_train_bwd``: batch statistics, running
    statistics updated in place, gradients to ``x``, ``weight`` and ``bias``.  Saves ``x``, the output and two [C] vectors
    (autograd through ``F.relu(bn(x))`` keeps the same two activations)."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, momentum, eps, relu):
        y, mean, invstd = ops.bn_train_fwd(x, weight, bias, running_mean, running_var, momentum, eps, relu)
        ctx.save_for_backward(x, y, weight, mean, invstd)
        ctx.relu = relu
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, weight, mean, invstd = ctx.saved_tensors
        dx, dw, db = ops.bn_train_bwd(x, y, dy, weight, mean, invstd, ctx.relu)
        return (dx if ctx.needs_input_grad[0] else None, dw if ctx.needs_input_grad[1] else None,
                db if ctx.needs_input_grad[2] else None, None, None, None, None, None)


def bn_act(bn: nn.modules.batchnorm._BatchNorm, x: torch.Tensor, relu: bool) -> torch.Tensor:
    """``relu(bn(x))`` of the reference's conv blocks.  Training mode on planar fp32 CUDA activations (the regulariser's
    [B,C,D,H,W] tensors) runs the fused B200 kernels; everything else (eval, channels_last, CPU, other dtypes,
    ``track_running_stats=False`` / ``momentum=None`` modules) is the stock PyTorch computation."""
    if (FUSED_TRAIN_BATCHNORM and bn.training and torch.is_grad_enabled() and bn.track_running_stats and bn.affine
            and bn.momentum is not None and ops.bn_train_supported(x) and x[0, 0].numel() * x.shape[0] > 1):
        if bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
        return _BatchNormReLUTrain.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, float(bn.momentum),
                                         float(bn.eps), bool(relu))
    y = bn(x)
    return F.relu(y, inplace=True) if relu else y


class _Conv3dHandWgrad(torch.autograd.Function):
    """``conv(x)`` of a bias-free ``nn.Conv3d`` / ``nn.ConvTranspose3d`` of the regulariser: cuDNN forward and data
    gradient, the weight gradient on ``ops.conv3d_wgrad`` (cuDNN's few-channel 3-D wgrad kernels were the largest
    entry of a training step)."""

    @staticmethod
    def forward(ctx, x, weight, conv):
        ctx.conv = conv
        ctx.save_for_backward(x, weight)
        return conv._conv_forward(x, weight, None) if isinstance(conv, nn.Conv3d) else \
            F.conv_transpose3d(x, weight, None, conv.stride, conv.padding, conv.output_padding, 1, conv.dilation)

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        conv = ctx.conv
        transposed = isinstance(conv, nn.ConvTranspose3d)
        gy = gy.contiguous()
        gx = None
        if ctx.needs_input_grad[0]:
            gx = torch.ops.aten.convolution_backward(gy, x, weight, None, list(conv.stride), list(conv.padding),
                                                     list(conv.dilation), transposed, list(conv.output_padding), 1,
                                                     [True, False, False])[0]
        gw = None
        if ctx.needs_input_grad[1]:
            a, b = (x, gy) if transposed else (gy, x)
            gw = ops.conv3d_wgrad(a, b, conv.kernel_size[0], conv.stride[1])
        return gx, gw, None


class _ProbHandWgrad(torch.autograd.Function):
    """``prob(x)`` (``nn.Conv3d(8, 1, 1)`` with bias): stock forward, ``grad_x = grad_out * w`` as a broadcast product,
    weight and bias gradient on ``ops.conv1x1_wgrad`` (one streaming pass)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        ctx.save_for_backward(x, weight)
        return F.conv3d(x, weight, bias)

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        gy = gy.contiguous()
        gx = gy * weight.view(1, -1, 1, 1, 1) if ctx.needs_input_grad[0] else None
        gw = gb = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            dw, db = ops.conv1x1_wgrad(x, gy)
            gw, gb = dw.view_as(weight), db
        return gx, gw, gb


def prob_train(conv: nn.Conv3d, x):
    if (HAND_WGRAD3D and conv.training and torch.is_grad_enabled() and conv.weight.requires_grad and conv.bias is not None
            and x.is_cuda and x.dtype == torch.float32 and x.dim() == 5 and x.is_contiguous() and x.shape[1] == 8
            and conv.out_channels == 1 and tuple(conv.kernel_size) == (1, 1, 1) and tuple(conv.stride) == (1, 1, 1)
            and tuple(conv.padding) == (0, 0, 0) and x[0, 0].numel() % 4 == 0 and x.data_ptr() % 16 == 0):
        return _ProbHandWgrad.apply(x, conv.weight, conv.bias)
    return conv(x)


class _Conv2dHandWgrad(torch.autograd.Function):
    """``conv(x)`` of a bias-free 3x3 stride-1 ``nn.Conv2d`` of FPN4 in training: cuDNN forward and data gradient in
    whatever memory format the activations have (channels_last in ``MVS4net.train()``), the weight gradient on
    ``ops.conv3d_wgrad`` over planar copies of ``x`` and ``grad_output`` (D = 1)."""

    @staticmethod
    def forward(ctx, x, weight, conv):
        ctx.conv = conv
        ctx.save_for_backward(x, weight)
        return conv._conv_forward(x, weight, None)

    @staticmethod
    def backward(ctx, gy):
        x, weight = ctx.saved_tensors
        conv = ctx.conv
        gx = None
        if ctx.needs_input_grad[0]:
            gx = torch.ops.aten.convolution_backward(gy, x, weight, None, list(conv.stride), list(conv.padding),
                                                     list(conv.dilation), False, [0, 0], 1, [True, False, False])[0]
        gw = None
        if ctx.needs_input_grad[1]:
            gw = ops.conv3d_wgrad(gy.contiguous().unsqueeze(2), x.contiguous().unsqueeze(2), 1, 1).squeeze(2)
            if weight.dim() == 4 and not weight.is_contiguous() and weight.is_contiguous(memory_format=torch.channels_last):
                gw = gw.contiguous(memory_format=torch.channels_last)
        return gx, gw, None


def conv2d_train(conv: nn.Conv2d, x):
    """``conv(x)`` with the hand-written weight gradient for FPN4's 3x3 stride-1 layers in training mode - only while
    cuDNN is held to fp32 (``torch.backends.cudnn.allow_tf32 = False``): measured per training step, the fp32 hand-written
    kernel beats cuDNN's fp32 NHWC weight gradient (10.8 -> 5.4 + 2.6 ms) and loses to its TF32 one (44.6 vs 39.0 ms)."""
    if (HAND_WGRAD2D and not torch.backends.cudnn.allow_tf32 and conv.training and torch.is_grad_enabled() and conv.weight.requires_grad and conv.bias is None
            and x.is_cuda and x.dtype == torch.float32 and x.dim() == 4 and conv.weight.dtype == torch.float32
            and tuple(conv.kernel_size) == (3, 3) and tuple(conv.stride) == (1, 1) and tuple(conv.padding) == (1, 1)
            and tuple(conv.dilation) == (1, 1) and conv.groups == 1 and conv.padding_mode == "zeros"
            and conv.out_channels % 8 == 0):
        return _Conv2dHandWgrad.apply(x, conv.weight, conv)
    return conv(x)


def _hand_wgrad_applies(conv, x) -> bool:
    if not (HAND_WGRAD3D and torch.is_grad_enabled() and conv.weight.requires_grad and conv.bias is None
            and x.is_cuda and x.dtype == torch.float32 and x.dim() == 5 and x.is_contiguous()
            and conv.weight.dtype == torch.float32 and conv.groups == 1 and tuple(conv.dilation) == (1, 1, 1)):
        return False
    kd, kh, kw = conv.kernel_size
    st = tuple(conv.stride)
    if (kh, kw) != (3, 3) or kd not in (1, 3) or tuple(conv.padding) != (kd // 2, 1, 1) or st[0] != 1 or st[1] != st[2]:
        return False
    if isinstance(conv, nn.ConvTranspose3d):
        return (st[1] == 2 and kd == 1 and tuple(conv.output_padding) == (0, 1, 1) and conv.in_channels % 8 == 0)
    return isinstance(conv, nn.Conv3d) and st[1] in (1, 2) and conv.out_channels % 8 == 0 and conv.padding_mode == "zeros"


def conv3d_train(conv, x):
    """``conv(x)`` with the hand-written weight gradient where it applies (training, planar fp32 CUDA, the
    regulariser's kernel shapes); the stock module call otherwise."""
    if conv.training and _hand_wgrad_applies(conv, x):
        return _Conv3dHandWgrad.apply(x, conv.weight, conv)
    return conv(x)


# module switches (tests / A-B timing; the environment variables set the initial value)
HAND_WGRAD3D = os.environ.get("MVSTER_TRAIN_CUDNN_WGRAD") is None  # False = cuDNN's weight gradient for the 3-D convolutions
HAND_WGRAD2D = os.environ.get("MVSTER_TRAIN_CUDNN_WGRAD2D") is None and HAND_WGRAD3D  # the same for FPN4's 3x3 layers
FUSED_TRAIN_BATCHNORM = os.environ.get("MVSTER_TRAIN_CUDNN_BATCHNORM") is None  # module switch (tests / A-B timing): False = stock nn.BatchNorm + F.relu in training


class Conv2d(nn.Module):
    """conv -> BatchNorm2d -> ReLU (reference mvs4net_utils.py:231-258, the ``gn=False`` branch)."""

    def __init__(self, cin: int, cout: int, kernel_size: int, stride: int = 1, padding: int = 0, relu: bool = True):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, kernel_size, stride=stride, padding=padding, bias=False)
        self.bn = nn.BatchNorm2d(cout, momentum=0.1)
        self.relu = relu

    def forward(self, x):
        return bn_act(self.bn, conv2d_train(self.conv, x), self.relu)


class _FoldedWeights:
    """Mixin of the modules that cache BatchNorm-folded eval-mode weights (host copies, device copies, kernel-parameter
    slices).  The caches are keyed on the parameters' ``_version`` + ``data_ptr()``, which an in-place update through
    ``.data`` (EMA swaps: ``p.data.copy_(ema)``) does NOT change: call ``invalidate_folded()`` after such an update.
    It is called automatically by ``train()``, ``load_state_dict()`` and ``_apply()`` (``.to()``, ``.half()``...).
    A ``GraphedMVS4net`` capture bakes the folded weights in as by-value kernel parameters: re-capture after any
    weight change."""

    def invalidate_folded(self):
        for m in self.modules():
            if isinstance(m, _FoldedWeights):
                m._drop_folded()

    def _drop_folded(self):
        if hasattr(self, "_fold_cache"):
            self._fold_cache = {}
        if hasattr(self, "_folded") and not callable(getattr(self, "_folded")):
            self._folded = None

    def train(self, mode: bool = True):
        self._drop_folded()
        return super().train(mode)

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_folded()
        return out

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._drop_folded()
        return out



class FPN4(_FoldedWeights, nn.Module):
    """Four-level feature pyramid, 8/4/2/1 x ``base_channels`` output channels at 1/8 .. 1/1 resolution
    (reference mvs4net_utils.py:426-509, ``gn=False, dcn=False``)."""

    def __init__(self, base_channels: int = 8, gn: bool = False, dcn: bool = False):
        super().__init__()
        if gn or dcn:
            raise NotImplementedError("FPN4(B200): GroupNorm / deformable-conv variants are not built")
        c = base_channels
        self.base_channels = c
        self.conv0 = nn.Sequential(Conv2d(3, c, 3, 1, 1), Conv2d(c, c, 3, 1, 1))
        self.conv1 = nn.Sequential(Conv2d(c, 2 * c, 5, 2, 2), Conv2d(2 * c, 2 * c, 3, 1, 1), Conv2d(2 * c, 2 * c, 3, 1, 1))
        self.conv2 = nn.Sequential(Conv2d(2 * c, 4 * c, 5, 2, 2), Conv2d(4 * c, 4 * c, 3, 1, 1), Conv2d(4 * c, 4 * c, 3, 1, 1))
        self.conv3 = nn.Sequential(Conv2d(4 * c, 8 * c, 5, 2, 2), Conv2d(8 * c, 8 * c, 3, 1, 1), Conv2d(8 * c, 8 * c, 3, 1, 1))
        top = 8 * c
        self.inner1 = nn.Conv2d(4 * c, top, 1, bias=True)
        self.inner2 = nn.Conv2d(2 * c, top, 1, bias=True)
        self.inner3 = nn.Conv2d(c, top, 1, bias=True)
        self.out1 = nn.Conv2d(top, 8 * c, 1, bias=False)
        self.out2 = nn.Conv2d(top, 4 * c, 3, padding=1, bias=False)
        self.out3 = nn.Conv2d(top, 2 * c, 3, padding=1, bias=False)
        self.out4 = nn.Conv2d(top, c, 3, padding=1, bias=False)
        self.out_channels = [8 * c, 4 * c, 2 * c, c]
        self.direct_convs = True  # eval: hand-written kernels for the encoder and the two finest top-down levels
        self.linear_topdown = True  # the two finest top-down levels in linearised form (False: ops.fpn_topdown)
        self._fold_cache = {}

    # ---- eval-mode path on the hand-written kernels ---------------------------------------------------------------
    def _folded(self, blk: "Conv2d", tag: str):
        """BatchNorm-folded weight of a Conv2d block as per-launch slices ``[k,k,ci,cs]`` + bias slices (host)."""
        conv, bn = blk.conv, blk.bn
        key = tuple(t._version for t in (conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var)) \
            + (conv.weight.data_ptr(),)
        hit = self._fold_cache.get(tag)
        if hit is None or hit[0] != key:
            scale = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
            shift = (bn.bias.detach().double() - bn.running_mean.detach().double() * scale).float().cpu()
            w = (conv.weight.detach().double() * scale.view(-1, 1, 1, 1)).permute(2, 3, 1, 0).float().cpu()  # [ky,kx,ci,co]
            cs = ops._CONV2D_SLICE[(conv.in_channels, conv.kernel_size[0])]
            n = conv.out_channels // cs
            hit = (key, [w[..., i * cs:(i + 1) * cs].contiguous() for i in range(n)],
                   [shift[i * cs:(i + 1) * cs].contiguous() for i in range(n)])
            self._fold_cache[tag] = hit
        return hit[1], hit[2]

    def _topdown_weights(self, out_conv: nn.Conv2d, inner: nn.Conv2d, tag: str):
        key = (out_conv.weight._version, inner.weight._version, inner.bias._version, out_conv.weight.data_ptr())
        hit = self._fold_cache.get(tag)
        if hit is None or hit[0] != key:
            w = out_conv.weight.detach().permute(2, 3, 1, 0).float().cpu()                      # [ky,kx,c64,co]
            slices = [w[..., i * 8:(i + 1) * 8].contiguous() for i in range(out_conv.out_channels // 8)]
            w_in = inner.weight.detach()[:, :, 0, 0].t().contiguous().float().cpu()             # [cl, c64]
            hit = (key, slices, w_in, inner.bias.detach().float().cpu().contiguous())
            self._fold_cache[tag] = hit
        return hit[1], hit[2], hit[3]

    def _topdown_lin_weights(self, device):
        """Weights of the two finest top-down levels in their linearised form (``ops.fpn_lin_gather`` /
        ``ops.fpn_project_up``), composed in float64:
        ``wq`` [64, 144+72] (device) projects ``intra2`` onto the taps of ``out3`` and of ``out4``;
        per level ``wc`` [9,Clat,Cout] = W[tap] Wi and ``bc`` [9,Cout] = W[tap] bi (host);
        ``wl`` [16,72] = Wp4 Wi2 and ``bl`` [72] = Wp4 b2 carry level 4's projection through level 3 (host)."""
        convs = (self.out3, self.inner2, self.out4, self.inner3)
        key = tuple(t._version for m in convs for t in m.parameters()) + (self.out3.weight.data_ptr(), str(device))
        hit = self._fold_cache.get("tdlin")
        if hit is None or hit[0] != key:
            def level(out_conv, inner):
                w = out_conv.weight.detach().double().cpu()                  # [co, c64, ky, kx]
                co = w.shape[0]
                wt = w.permute(2, 3, 0, 1).reshape(9, co, 64)                # [tap, co, c64]
                wi = inner.weight.detach().double().cpu()[:, :, 0, 0]        # [c64, cl]
                bi = inner.bias.detach().double().cpu()
                wc = torch.matmul(wt, wi).permute(0, 2, 1).contiguous().float()   # [tap, cl, co]
                bc = torch.matmul(wt, bi).contiguous().float()                    # [tap, co]
                return wt.reshape(9 * co, 64), wc, bc
            wp3, wc3, bc3 = level(self.out3, self.inner2)
            wp4, wc4, bc4 = level(self.out4, self.inner3)
            wi2 = self.inner2.weight.detach().double().cpu()[:, :, 0, 0]     # [c64, 16]
            wl = torch.matmul(wp4, wi2).t().contiguous().float()             # [16, 72]
            bl = torch.matmul(wp4, self.inner2.bias.detach().double().cpu()).contiguous().float()
            wq = torch.cat([wp3, wp4], 0).t().contiguous().float().to(device)   # [64, 216]
            hit = (key, dict(wq=wq, wc3=wc3, bc3=bc3, wc4=wc4, bc4=bc4, wl=wl, bl=bl, n3=wp3.shape[0], n4=wp4.shape[0]))
            self._fold_cache["tdlin"] = hit
        return hit[1]

    def direct_supported(self, x) -> bool:
        return (self.direct_convs and not self.training and self.base_channels == 8 and x.is_cuda
                and x.shape[2] % 8 == 0 and x.shape[3] % 16 == 0)

    def _block(self, blk: "Conv2d", tag: str, x):
        k, s = blk.conv.kernel_size[0], blk.conv.stride[0]
        if k == 5 and s == 2 and ops.conv2d_mid5_supported(blk.conv.in_channels, blk.conv.out_channels, x.shape[2], x.shape[3]):
            w, b = self._folded_dev(blk, tag, x.device)
            return ops.conv2d_mid5(x, w[0], b)
        if ops.conv2d_small_supported(blk.conv.in_channels, blk.conv.out_channels, k, s, x.shape[2], x.shape[3]):
            w, b = self._folded(blk, tag)
            return ops.conv2d_small(x, w, b, k, s, True)
        if k == 3 and s == 1 and ops.conv3d_mid_supported(blk.conv.in_channels, blk.conv.out_channels, 1, x.shape[2], x.shape[3]):
            w, b = self._folded_dev(blk, tag, x.device)
            return ops.conv3d_mid(x.unsqueeze(2), w, b).squeeze(2)
        return blk(x)

    def _folded_dev(self, blk: "Conv2d", tag: str, device):
        """BatchNorm-folded weight ``[1,3,3,ci,co]`` + bias of a Conv2d block, resident on ``device`` (cached)."""
        conv, bn = blk.conv, blk.bn
        key = tuple(t._version for t in (conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var)) \
            + (conv.weight.data_ptr(), str(device))
        hit = self._fold_cache.get(tag + "@dev")
        if hit is None or hit[0] != key:
            scale = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
            shift = (bn.bias.detach().double() - bn.running_mean.detach().double() * scale).float()
            w = (conv.weight.detach().double() * scale.view(-1, 1, 1, 1)).permute(2, 3, 1, 0).float()  # [ky,kx,ci,co]
            hit = (key, w.unsqueeze(0).contiguous().to(device), shift.contiguous().to(device))
            self._fold_cache[tag + "@dev"] = hit
        return hit[1], hit[2]

    def forward_direct(self, x, feature_dtype: Optional[torch.dtype] = None) -> Dict[str, torch.Tensor]:
        """Eval-mode FPN4 on the B200 kernels: encoder levels 0-2 as direct convolutions (BatchNorm folded), level 3
        and the coarse top-down steps on cuDNN, the two finest top-down steps fused with their output convolutions
        (``ops.fpn_topdown``).  Every output is NHWC in memory (``channels_last`` strides), ready for K1.
        ``feature_dtype=torch.bfloat16``: the four feature maps are EMITTED in bf16 (stages 3-4 by the top-down kernel's
        epilogue, stages 1-2 by the planar -> NHWC transpose that the fp32 path runs as well) - K1 consumes them
        zero-copy, there is no separate cast pass over the feature maps."""
        x = x.contiguous()
        fdt = feature_dtype or torch.float32
        c0 = self._block(self.conv0[1], "c01", self._block(self.conv0[0], "c00", x))
        c1 = c0
        for i, blk in enumerate(self.conv1):
            c1 = self._block(blk, "c1%d" % i, c1)
        c2 = c1
        for i, blk in enumerate(self.conv2):
            c2 = self._block(blk, "c2%d" % i, c2)
        top = c2
        for i, blk in enumerate(self.conv3):
            top = self._block(blk, "c3%d" % i, top)
        out = {"stage1": ops.to_nhwc(self.out1(top), fdt).permute(0, 3, 1, 2)}
        intra2 = self._up(top) + self.inner1(c2)
        if ops.conv3d_mid_supported(self.out2.in_channels, self.out2.out_channels, 1, intra2.shape[2], intra2.shape[3]):
            key = (self.out2.weight._version, self.out2.weight.data_ptr(), str(x.device))
            hit = self._fold_cache.get("out2@dev")
            if hit is None or hit[0] != key:   # plain 3x3 convolution, no BatchNorm / bias / ReLU (mvs4net_utils.py:468)
                w2 = self.out2.weight.detach().float().permute(2, 3, 1, 0).unsqueeze(0).contiguous().to(x.device)
                hit = (key, w2, torch.zeros(self.out2.out_channels, device=x.device))
                self._fold_cache["out2@dev"] = hit
            feat2 = ops.conv3d_mid(intra2.contiguous().unsqueeze(2), hit[1], hit[2], relu=False).squeeze(2)
        else:
            feat2 = self.out2(intra2)
        out["stage2"] = ops.to_nhwc(feat2, fdt).permute(0, 3, 1, 2)
        if self.linear_topdown:
            # the two finest levels through their linearity: neither the half- nor the full-resolution 64-channel
            # intra exists; one GEMM projects intra2 onto the taps of out3 and out4 (see fpn_lin.cu)
            lw = self._topdown_lin_weights(x.device)
            q = ops.fpn_project(intra2, lw["wq"])                                         # [B,H/4,W/4,144+72]
            feat3 = ops.fpn_lin_gather(q, 0, c1, lw["wc3"], lw["bc3"], fdt)               # [B,H/2,W/2,16]
            p4 = ops.fpn_project_up(q, lw["n3"], lw["n4"], c1, lw["wl"], lw["bl"])        # [B,H/2,W/2,72]
            feat4 = ops.fpn_lin_gather(p4, 0, c0, lw["wc4"], lw["bc4"], fdt)              # [B,H,W,8]
        else:
            w3, wi3, bi3 = self._topdown_weights(self.out3, self.inner2, "td3")
            feat3, intra3 = ops.fpn_topdown(intra2, c1, w3, wi3, bi3, want_intra=True, feature_dtype=fdt)
            w4, wi4, bi4 = self._topdown_weights(self.out4, self.inner3, "td4")
            feat4, _ = ops.fpn_topdown(intra3, c0, w4, wi4, bi4, want_intra=False, feature_dtype=fdt)
        out["stage3"] = feat3.permute(0, 3, 1, 2)
        out["stage4"] = feat4.permute(0, 3, 1, 2)
        return out

    @staticmethod
    def _up(x):
        return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)

    def forward(self, x) -> Dict[str, torch.Tensor]:
        c0 = self.conv0(x)
        c1 = self.conv1(c0)
        c2 = self.conv2(c1)
        top = self.conv3(c2)
        out = {"stage1": self.out1(top)}
        top = self._up(top) + self.inner1(c2)
        out["stage2"] = conv2d_train(self.out2, top)
        top = self._up(top) + self.inner2(c1)
        out["stage3"] = conv2d_train(self.out3, top)
        top = self._up(top) + self.inner3(c0)
        out["stage4"] = conv2d_train(self.out4, top)
        return out


class ConvBnReLU3D(nn.Module):
    """reference mvs4net_utils.py:123-130"""

    def __init__(self, cin, cout, kernel_size=3, stride=1, pad=1):
        super().__init__()
        self.conv = nn.Conv3d(cin, cout, kernel_size, stride=stride, padding=pad, bias=False)
        self.bn = nn.BatchNorm3d(cout)

    def forward(self, x):
        return bn_act(self.bn, conv3d_train(self.conv, x), True)


def _up3d(cin, cout):
    return nn.Sequential(
        nn.ConvTranspose3d(cin, cout, kernel_size=(1, 3, 3), padding=(0, 1, 1), output_padding=(0, 1, 1),
                           stride=(1, 2, 2), bias=False),
        nn.BatchNorm3d(cout), nn.ReLU(inplace=True))


class reg2d(_FoldedWeights, nn.Module):
    """Cost regulariser (reference mvs4net_utils.py:884-926): a 3-level U-Net over (H, W) with (1,3,3) strided and
    (3,3,3) plain convolutions, ``[B, G, D, H, W] -> [B, D, H, W]`` logits.

    ``forward(x)`` is the reference computation.  ``forward_fused_tail(x, hypo, split_itv, ...)`` (eval only) stops
    before ``conv11`` and lets ``ops.regtail`` do ``conv0 + relu(bn(conv11(.)))`` -> ``prob`` -> softmax / arg-max /
    confidence / inverse range in one kernel.
    """

    def __init__(self, input_channel=128, base_channel=32, conv_name="ConvBnReLU3D"):
        super().__init__()
        if conv_name != "ConvBnReLU3D":
            raise NotImplementedError("reg2d(B200): only agg_type='ConvBnReLU3D' is built (attention variants "
                                      "ConvBnReLU3D_CAM/PAM/... are not used by any shipped config)")
        c = base_channel
        k, p = (1, 3, 3), (0, 1, 1)
        self.conv0 = ConvBnReLU3D(input_channel, c, kernel_size=k, pad=p)
        self.conv1 = ConvBnReLU3D(c, 2 * c, kernel_size=k, stride=(1, 2, 2), pad=p)
        self.conv2 = ConvBnReLU3D(2 * c, 2 * c)
        self.conv3 = ConvBnReLU3D(2 * c, 4 * c, kernel_size=k, stride=(1, 2, 2), pad=p)
        self.conv4 = ConvBnReLU3D(4 * c, 4 * c)
        self.conv5 = ConvBnReLU3D(4 * c, 8 * c, kernel_size=k, stride=(1, 2, 2), pad=p)
        self.conv6 = ConvBnReLU3D(8 * c, 8 * c)
        self.conv7 = _up3d(8 * c, 4 * c)
        self.conv9 = _up3d(4 * c, 2 * c)
        self.conv11 = _up3d(2 * c, c)
        self.prob = nn.Conv3d(8, 1, 1, stride=1, padding=0)  # the reference hard-codes 8 (mvs4net_utils.py:914)
        self._folded = None
        self._fold_cache = {}
        self.direct_convs = True  # eval-mode fused path: hand-written direct convolutions where available

    @staticmethod
    def _up(seq: nn.Sequential, x):
        """``_up3d`` block (ConvTranspose3d, BatchNorm3d, ReLU) with the BatchNorm + ReLU pair through ``bn_act``."""
        return bn_act(seq[1], conv3d_train(seq[0], x), True)

    def _trunk(self, x):
        conv0 = self.conv0(x)
        conv2 = self.conv2(self.conv1(conv0))
        conv4 = self.conv4(self.conv3(conv2))
        x = self.conv6(self.conv5(conv4))
        x = conv4 + self._up(self.conv7, x)
        x = conv2 + self._up(self.conv9, x)
        return conv0, x

    def forward(self, x):
        conv0, x = self._trunk(x)
        x = conv0 + self._up(self.conv11, x)
        return prob_train(self.prob, x).squeeze(1)

    # ---- fused last layers + tail -----------------------------------------------------------------------------------
    def fused_tail_supported(self) -> bool:
        return (not self.training) and self.conv11[0].out_channels == 8 and self.conv11[0].in_channels == 16

    def _fold(self):
        """conv11's BatchNorm (eval) folded into the transposed-conv weight, as HOST tensors: ``w [ky,kx,ci,co]`` and
        ``params = [bn shift (8), prob weight (8), prob bias]`` (cached until a parameter changes)."""
        deconv, bn = self.conv11[0], self.conv11[1]
        key = tuple(t._version for t in (deconv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var,
                                         self.prob.weight, self.prob.bias)) + (deconv.weight.data_ptr(),)
        if self._folded is None or self._folded[0] != key:
            scale = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
            shift = bn.bias.detach().double() - bn.running_mean.detach().double() * scale
            w = deconv.weight.detach().double()[:, :, 0] * scale.view(1, -1, 1, 1)        # [ci, co, ky, kx]
            w = w.permute(2, 3, 0, 1).contiguous().float().cpu()                          # [ky, kx, ci, co]
            params = torch.cat([shift.float().cpu(), self.prob.weight.detach().reshape(-1).float().cpu(),
                                self.prob.bias.detach().reshape(-1).float().cpu()]).contiguous()
            self._folded = (key, w, params)
        return self._folded[1], self._folded[2]

    def _fold_block(self, name: str):
        """(w_host [kd,3,3,ci,co], bias_host [co]) of a conv/deconv + BatchNorm block in eval mode, cached."""
        blk = getattr(self, name)
        conv, bn = (blk.conv, blk.bn) if isinstance(blk, ConvBnReLU3D) else (blk[0], blk[1])
        key = tuple(t._version for t in (conv.weight, bn.weight, bn.bias, bn.running_mean, bn.running_var)) \
            + (conv.weight.data_ptr(),)
        hit = self._fold_cache.get(name)
        if hit is None or hit[0] != key:
            scale = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
            shift = bn.bias.detach().double() - bn.running_mean.detach().double() * scale
            w = conv.weight.detach().double()
            if isinstance(conv, nn.ConvTranspose3d):          # [ci, co, kd, ky, kx]
                w = (w * scale.view(1, -1, 1, 1, 1)).permute(2, 3, 4, 0, 1)
            else:                                             # [co, ci, kd, ky, kx]
                w = (w * scale.view(-1, 1, 1, 1, 1)).permute(2, 3, 4, 1, 0)
            hit = (key, w.contiguous().float().cpu(), shift.float().cpu().contiguous())
            self._fold_cache[name] = hit
        return hit[1], hit[2]

    def _direct(self, name: str, x, mode: int, skip=None):
        """One block through the direct-convolution kernel when it has this layer, else through cuDNN."""
        blk = getattr(self, name)
        conv = blk.conv if isinstance(blk, ConvBnReLU3D) else blk[0]
        cin, cout = conv.in_channels, conv.out_channels
        if self.direct_convs and ops.conv3d_small_supported(cin, cout, conv.kernel_size[0], mode, x.shape[3], x.shape[4]):
            w, bias = self._fold_block(name)
            return ops.conv3d_small(x, w, bias, mode, True, skip)
        if self.direct_convs and ops.conv3d_sliced_supported(cin, cout, conv.kernel_size[0], mode, x.shape[3], x.shape[4]):
            w, bias = self._fold_block(name)
            hit = self._fold_cache.get(name + "@slices")
            if hit is None or hit[0] is not w:
                cs = ops._CONV3D_SLICE[(cin, cout, mode)]
                hit = (w, [w[..., i * cs:(i + 1) * cs].contiguous() for i in range(cout // cs)],
                       [bias[i * cs:(i + 1) * cs].contiguous() for i in range(cout // cs)])
                self._fold_cache[name + "@slices"] = hit
            return ops.conv3d_sliced(x, hit[1], hit[2], mode, True, skip)
        if self.direct_convs and mode == ops.CONV_STRIDE1 and skip is None \
                and ops.conv3d_mid_supported(cin, cout, conv.kernel_size[0], x.shape[3], x.shape[4]):
            w, bias = self._fold_block(name)
            return ops.conv3d_mid(x, *self._on_device(name, w, bias, x.device))
        y = blk(x)
        return y if skip is None else skip + y

    def _on_device(self, name: str, w: torch.Tensor, bias: torch.Tensor, device):
        """Device copies of a folded filter bank (cached until the fold itself is refreshed)."""
        hit = self._fold_cache.get(name + "@dev")
        if hit is None or hit[0] is not w or hit[1].device != device:
            hit = (w, w.to(device), bias.to(device))
            self._fold_cache[name + "@dev"] = hit
        return hit[1], hit[2]

    def _trunk_eval(self, x):
        """``_trunk`` in eval mode on hand-written direct convolutions with BatchNorm folded and ReLU / skip add fused:
        conv0-conv3, conv9 with the filter bank in the kernel-parameter space, conv5 / conv7 as one launch per filter-bank
        slice, conv4 / conv6 (110 / 442 KB of weights) with device-resident weights; cuDNN only where a shape is odd."""
        conv0 = self._direct("conv0", x, ops.CONV_STRIDE1)
        conv2 = self._direct("conv2", self._direct("conv1", conv0, ops.CONV_STRIDE2), ops.CONV_STRIDE1)
        conv4 = self._direct("conv4", self._direct("conv3", conv2, ops.CONV_STRIDE2), ops.CONV_STRIDE1)
        y = self._direct("conv6", self._direct("conv5", conv4, ops.CONV_STRIDE2), ops.CONV_STRIDE1)
        y = self._direct("conv7", y, ops.CONV_TRANSPOSED2, skip=conv4)
        return conv0, self._direct("conv9", y, ops.CONV_TRANSPOSED2, skip=conv2)

    def forward_direct(self, x):
        """Eval-mode logits with the direct-convolution trunk (same result as ``forward`` up to fp32 summation order);
        used by the parity tests and when the caller wants the logits rather than the fused tail."""
        if self.training:
            raise RuntimeError("reg2d.forward_direct is an eval-mode path")
        conv0, low = self._trunk_eval(x)
        return self.prob(self._direct("conv11", low, ops.CONV_TRANSPOSED2, skip=conv0)).squeeze(1)

    def forward_fused_tail(self, x, depth_hypo, split_itv, inverse_depth=True, depth_mode=ops.DEPTH_ARGMAX):
        if not self.fused_tail_supported():
            raise RuntimeError("reg2d.forward_fused_tail: eval mode and base_channel == 8 required")
        conv0, low = self._trunk_eval(x)
        w, params = self._fold()
        return ops.regtail(low, conv0, w, params, depth_hypo, float(split_itv), bool(inverse_depth), depth_mode)


# ----------------------------------------------------------------------------------------------------------------------
# the network
# ----------------------------------------------------------------------------------------------------------------------
class MVS4net(_FoldedWeights, nn.Module):
    """Reference ``MVS4net`` (models/MVS4Net.py:16) on the B200 hot path: same constructor arguments, same
    ``forward(imgs, proj_matrices, depth_values, filename=None)`` and the same nested output dictionary."""

    def __init__(self, arch_mode="fpn", reg_net="reg2d", num_stage=4, fpn_base_channel=8, reg_channel=8,
                 stage_splits=(8, 8, 4, 4), depth_interals_ratio=(0.5, 0.5, 0.5, 1), group_cor=False,
                 group_cor_dim=(8, 8, 8, 8), inverse_depth=False, agg_type="ConvBnReLU3D", dcn=False, pos_enc=0,
                 mono=False, mono_stg_itrpl="nearest", asff=False, attn_temp=2, attn_fuse_d=True, vis_ETA=False,
                 vis_stg_features=False, debug=0, *, fuse_regnet_tail: bool = True,
                 feature_dtype: Optional[torch.dtype] = None):
        super().__init__()
        unsupported = {"arch_mode": arch_mode != "fpn", "reg_net": reg_net != "reg2d", "dcn": bool(dcn),
                       "asff": bool(asff), "mono": bool(mono), "vis_ETA": bool(vis_ETA), "debug": bool(debug),
                       "inverse_depth=False": not inverse_depth}
        bad = [k for k, v in unsupported.items() if v]
        if bad:
            raise NotImplementedError("MVS4net(B200): unsupported option(s) %s - only the shipped fpn / reg2d / "
                                      "inverse-depth configuration is built (schedule_range is broken upstream, "
                                      "mvs4net_utils.py:102)" % ", ".join(bad))
        self.arch_mode, self.num_stage = arch_mode, num_stage
        self.depth_interals_ratio = list(depth_interals_ratio)
        self.group_cor, self.group_cor_dim = group_cor, list(group_cor_dim)
        self.inverse_depth = inverse_depth
        self.stage_splits = list(stage_splits)
        self.pos_enc, self.mono, self.asff, self.debug = pos_enc, mono, asff, debug
        self.attn_ob = nn.ModuleList()
        self.pos_enc_func = nn.ModuleList()
        self.feature = FPN4(base_channels=fpn_base_channel)
        self.stagenet = stagenet(inverse_depth, mono, attn_fuse_d, vis_ETA, attn_temp, debug=debug,
                                 feature_dtype=feature_dtype)
        self.reg = nn.ModuleList()
        for idx in range(num_stage):
            in_dim = self.group_cor_dim[idx] if group_cor else self.feature.out_channels[idx]
            self.reg.append(reg2d(input_channel=in_dim, base_channel=reg_channel, conv_name=agg_type))
        self.fuse_regnet_tail = fuse_regnet_tail

    # ---- step 1: features -------------------------------------------------------------------------------------------
    def extract_features(self, imgs: Sequence[torch.Tensor]) -> List[Dict[str, torch.Tensor]]:
        """Per-view stage dictionaries, NHWC in memory.  Eval: one FPN pass over all views (BatchNorm uses running
        statistics, so stacking views along the batch is the same arithmetic as the reference's per-view loop);
        training: per view, as the reference, because batch statistics depend on what is in the batch."""
        n = len(imgs)
        if self.training:
            return [self.feature(img.contiguous(memory_format=torch.channels_last)) for img in imgs]
        b = imgs[0].shape[0]
        stacked = torch.cat(list(imgs), 0)
        if self.feature.direct_supported(stacked) and not torch.is_grad_enabled():
            out = self.feature.forward_direct(stacked, self.stagenet.feature_dtype)
        else:
            out = self.feature(stacked.contiguous(memory_format=torch.channels_last))
        return [{k: v[i * b:(i + 1) * b] for k, v in out.items()} for i in range(n)]

    # ---- forward ------------------------------------------------------------------------------------------------------
    def forward(self, imgs, proj_matrices, depth_values, filename=None):
        features = self.extract_features(imgs)
        outputs = {}
        stage_out = None
        for s in range(self.num_stage):
            key = "stage%d" % (s + 1)
            feats = [f[key] for f in features]
            proj = proj_matrices[key]
            h, w = feats[0].shape[2:]
            d = self.stage_splits[s]
            if s == 0:
                hypo = init_inverse_range(depth_values, d, None, None, h, w)
            else:
                hypo = schedule_inverse_range(stage_out["inverse_min_depth"].detach(),
                                              stage_out["inverse_max_depth"].detach(), d, h, w)
            regnet = self.reg[s]
            fused = (self.fuse_regnet_tail and not self.training and not torch.is_grad_enabled()
                     and regnet.fused_tail_supported() and d in (4, 8))
            if fused:
                stage_out = self.stagenet.forward_fused_regnet(feats, proj, hypo, regnet, s, group_cor=self.group_cor,
                                                               group_cor_dim=self.group_cor_dim[s],
                                                               split_itv=self.depth_interals_ratio[s])
            else:
                stage_out = self.stagenet(feats, proj, depth_hypo=hypo, regnet=regnet, stage_idx=s,
                                          group_cor=self.group_cor, group_cor_dim=self.group_cor_dim[s],
                                          split_itv=self.depth_interals_ratio[s], fn=filename)
            outputs[key] = stage_out
        return outputs


class GraphedMVS4net:
    """Inference server front end: one whole ``MVS4net.forward`` (images in, 4-stage depth out) captured as a CUDA graph.

    The forward issues ~150 small launches (cuDNN layers, this library's kernels, elementwise glue); at one scene per
    call the host-side launch cost is as large as the GPU time.  All shapes are static per (B, N, H0, W0), every kernel
    of this library takes raw pointers and by-value weights, so the whole forward replays as a single graph launch.

        g = GraphedMVS4net(model, batch=1, nviews=5, height=832, width=1152)
        out = g(imgs, proj_matrices, depth_values)      # same nested dict as model(...); tensors are graph-owned
    """

    def __init__(self, model: MVS4net, batch: int, nviews: int, height: int, width: int, device="cuda"):
        if model.training:
            raise RuntimeError("GraphedMVS4net: put the model in eval mode first")
        self.model = model
        dev = torch.device(device)
        self.imgs = [torch.zeros((batch, 3, height, width), device=dev) for _ in range(nviews)]
        self.proj = {"stage%d" % (s + 1): torch.zeros((batch, nviews, 2, 4, 4), device=dev) for s in range(model.num_stage)}
        self.depth_values = torch.ones((batch, 2), device=dev)
        self.graph = None
        self.out = None
        self._stage, self._pending = None, False

    def _load(self, imgs, proj_matrices, depth_values):
        for dst, src in zip(self.imgs, imgs):
            dst.copy_(src, non_blocking=True)
        for k, dst in self.proj.items():
            dst.copy_(proj_matrices[k], non_blocking=True)
        self.depth_values.copy_(depth_values[:, [0, -1]], non_blocking=True)

    def capture(self, imgs, proj_matrices, depth_values, warmup: int = 2):
        """Warm up on real inputs (cuDNN algorithm selection, lazy module loads, weight folding), then capture."""
        self._load(imgs, proj_matrices, depth_values)
        with torch.no_grad():
            for _ in range(warmup):
                self.model(self.imgs, self.proj, self.depth_values)
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = self.model(self.imgs, self.proj, self.depth_values)
        return self

    def __call__(self, imgs, proj_matrices, depth_values):
        if self.graph is None:
            self.capture(imgs, proj_matrices, depth_values)
        self._load(imgs, proj_matrices, depth_values)
        self.graph.replay()
        return self.out

    # ---- pipelined serving: the next request's inputs are uploaded while the current one computes ---------------------
    def prefetch(self, imgs, proj_matrices, depth_values):
        """Start uploading the NEXT request (pinned host tensors, or device tensors) into staging buffers on a side
        stream; returns immediately.  ``run_prefetched()`` consumes it.  While a forward runs (~6 ms per 832x1152 scene) the
        57 MB of images per scene cross the link concurrently instead of in front of it."""
        if self.graph is None:
            self.capture(imgs, proj_matrices, depth_values)
        dev = self.imgs[0].device
        if self._stage is None:
            self._stage = ([torch.empty_like(t) for t in self.imgs], {k: torch.empty_like(t) for k, t in self.proj.items()},
                           torch.empty_like(self.depth_values))
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._staged, self._consumed = torch.cuda.Event(), torch.cuda.Event()
            self._consumed.record(torch.cuda.current_stream(dev))
        s_imgs, s_proj, s_dv = self._stage
        self._copy_stream.wait_event(self._consumed)      # the previous request has left the staging buffers
        if any(t.is_cuda for t in list(imgs) + [depth_values]):
            # device-resident inputs may still be being written on the caller's stream: order the side stream behind it
            self._copy_stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(self._copy_stream):
            for dst, src in zip(s_imgs, imgs):
                dst.copy_(src, non_blocking=True)
            for k, dst in s_proj.items():
                dst.copy_(proj_matrices[k], non_blocking=True)
            s_dv.copy_(depth_values[:, [0, -1]], non_blocking=True)
            self._staged.record(self._copy_stream)
        self._pending = True

    def run_prefetched(self):
        """Forward of the request handed to ``prefetch()``: staging -> graph inputs (device copy), one graph replay.
        The returned tensors are graph-owned and valid until the next replay."""
        if not self._pending:
            raise RuntimeError("GraphedMVS4net.run_prefetched(): call prefetch(imgs, proj_matrices, depth_values) first")
        dev = self.imgs[0].device
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(self._staged)
        s_imgs, s_proj, s_dv = self._stage
        for dst, src in zip(self.imgs, s_imgs):
            dst.copy_(src, non_blocking=True)
        for k, dst in self.proj.items():
            dst.copy_(s_proj[k], non_blocking=True)
        self.depth_values.copy_(s_dv, non_blocking=True)
        self._consumed.record(cur)
        self._pending = False
        self.graph.replay()
        return self.out
