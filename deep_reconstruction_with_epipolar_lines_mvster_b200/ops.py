"""Thin torch-tensor front end of the C ABI: argument checking, output allocation, stream and pointer plumbing.

PyTorch is used here only for device memory and streams; all arithmetic happens in libmvster_b200.so.
Every function raises RuntimeError when the CUDA library is missing or a tensor is not on a CUDA device - there is
no CPU path.
"""
from __future__ import annotations

import ctypes
import os
import weakref
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib

F32, BF16 = 0, 1
DEPTH_ARGMAX, DEPTH_REGRESS = 0, 1
MAX_SRC_VIEWS = 15


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor: the MVSTER B200 path has no CPU fallback" % name)


def _stream(t: torch.Tensor) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _ptr_array(ts: Sequence[torch.Tensor]):
    arr = (ctypes.c_void_p * len(ts))()
    for i, t in enumerate(ts):
        arr[i] = t.data_ptr()
    return arr


def _dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise RuntimeError("features must be float32 or bfloat16, got %s" % t.dtype)


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    _require_cuda(t, name)
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


# ------------------------------------------------------------------------------------------------------------------
def to_nhwc(feat: torch.Tensor, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """``[B,C,H,W]``-shaped tensor -> contiguous ``[B,H,W,C]`` tensor of ``dtype`` (default: keep).

    Zero-copy when ``feat`` already has channels_last strides and the right dtype; NCHW-contiguous fp32 input goes
    through the library's transpose(+cast) kernel (one pass).
    """
    _require_cuda(feat, "features")
    if feat.dim() != 4:
        raise RuntimeError("features must be [B,C,H,W], got %s" % (tuple(feat.shape),))
    dtype = dtype or feat.dtype
    if dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError("feature dtype must be float32 or bfloat16")
    nhwc_view = feat.permute(0, 2, 3, 1)
    if nhwc_view.is_contiguous():
        out = nhwc_view if feat.dtype == dtype else nhwc_view.to(dtype)
    elif feat.dtype == torch.float32 and feat.is_contiguous():
        b, c, h, w = feat.shape
        out = torch.empty((b, h, w, c), device=feat.device, dtype=dtype)
        lib = _lib.load()
        _lib.check(lib.mvster_nchw_to_nhwc(_ptr(feat), _ptr(out), b, c, h, w, BF16 if dtype == torch.bfloat16 else F32,
                                           _stream(feat)))
    else:
        out = nhwc_view.contiguous().to(dtype)
    if out.data_ptr() % 32:
        out = out.clone()
    return out


def compose_homographies(proj_matrices: torch.Tensor) -> torch.Tensor:
    """``[B,N,2,4,4]`` -> ``rt [B,N-1,12]`` fp32 (reference mvs4net_utils.py:1047-1050 + :32-34)."""
    proj = _f32c(proj_matrices, "proj_matrices")
    if proj.dim() != 5 or tuple(proj.shape[2:]) != (2, 4, 4):
        raise RuntimeError("proj_matrices must be [B,N,2,4,4], got %s" % (tuple(proj.shape),))
    b, n = proj.shape[:2]
    rt = torch.empty((b, n - 1, 12), device=proj.device, dtype=torch.float32)
    _lib.check(_lib.load().mvster_compose_homographies(_ptr(proj), _ptr(rt), b, n, _stream(proj)))
    return rt


def compose_homography_pair(src_proj: torch.Tensor, ref_proj: torch.Tensor) -> torch.Tensor:
    sp, rp = _f32c(src_proj, "src_proj"), _f32c(ref_proj, "ref_proj")
    b = sp.shape[0]
    rt = torch.empty((b, 12), device=sp.device, dtype=torch.float32)
    _lib.check(_lib.load().mvster_compose_homography_pair(_ptr(sp), _ptr(rp), _ptr(rt), b, _stream(sp)))
    return rt


def _check_k1_inputs(ref_nhwc: torch.Tensor, srcs_nhwc: Sequence[torch.Tensor], rt: torch.Tensor, hypo: torch.Tensor):
    """Shape / dtype / device / contiguity contract shared by every K1 entry point (the kernels index raw pointers:
    a smaller or foreign source map would be read out of bounds, not rejected).  Returns ``(hypo fp32 contiguous, d)``."""
    _require_cuda(ref_nhwc, "ref")
    if ref_nhwc.dim() != 4 or not ref_nhwc.is_contiguous():
        raise RuntimeError("ref features must be contiguous NHWC [B,H,W,C]")
    b, h, w, c = ref_nhwc.shape
    nsrc = len(srcs_nhwc)
    if nsrc == 0:
        raise RuntimeError("at least one source view is required")
    for s in srcs_nhwc:
        if s.dim() != 4 or s.shape != srcs_nhwc[0].shape or s.shape[0] != b or s.shape[3] != c or s.dtype != ref_nhwc.dtype \
                or not s.is_contiguous() or s.device != ref_nhwc.device:
            raise RuntimeError("source features must share shape [B,Hs,Ws,C], dtype, device and be contiguous")
    hypo = _f32c(hypo, "depth_hypo")
    if hypo.dim() != 4 or hypo.device != ref_nhwc.device or tuple(hypo.shape) != (b, hypo.shape[1], h, w):
        raise RuntimeError("depth_hypo must be [B,D,H,W] = [%d,D,%d,%d] on %s, got %s" % (b, h, w, ref_nhwc.device,
                                                                                         tuple(hypo.shape)))
    if tuple(rt.shape) != (b, nsrc, 12) or rt.dtype != torch.float32 or not rt.is_contiguous() or rt.device != ref_nhwc.device:
        raise RuntimeError("rt must be contiguous fp32 [B,Nsrc,12] on the features' device")
    return hypo, hypo.shape[1]


def epi_fwd(ref_nhwc: torch.Tensor, srcs_nhwc: Sequence[torch.Tensor], rt: torch.Tensor, hypo: torch.Tensor,
            groups: int, attn_temp: float, want_wsum: bool = False, want_weights: bool = False
            ) -> Tuple[torch.Tensor, Optional[torch.Tensor], Optional[torch.Tensor]]:
    """Fused K1 forward on NHWC features.  Returns ``(volume [B,G,D,H,W], wsum or None, weights or None)``."""
    hypo, d = _check_k1_inputs(ref_nhwc, srcs_nhwc, rt, hypo)
    b, h, w, c = ref_nhwc.shape
    nsrc = len(srcs_nhwc)
    hs, ws = srcs_nhwc[0].shape[1:3]
    dev = ref_nhwc.device
    out = torch.empty((b, groups, d, h, w), device=dev, dtype=torch.float32)
    wsum = torch.empty((b, d, h, w), device=dev, dtype=torch.float32) if want_wsum else None
    weights = torch.empty((b, nsrc, d, h, w), device=dev, dtype=torch.float32) if want_weights else None
    _lib.check(_lib.load().mvster_epi_fwd(
        _ptr(ref_nhwc), _ptr_array(srcs_nhwc), _ptr(rt), _ptr(hypo), _ptr(out), _ptr(wsum), _ptr(weights),
        b, nsrc, c, groups, d, h, w, hs, ws, float(attn_temp), _dtype_code(ref_nhwc), _stream(ref_nhwc)))
    return out, wsum, weights


def epi_fwd_mode(ref_nhwc: torch.Tensor, srcs_nhwc: Sequence[torch.Tensor], rt: torch.Tensor, hypo: torch.Tensor,
                 groups: int, attn_temp: float, group_cor: bool, attn_fuse_d: bool, want_wsum: bool = False):
    """Forward of the reference's non-default options (variance cost / per-pixel weight).  Returns the volume, or
    ``(volume, wsum)`` with ``want_wsum`` (what ``epi_bwd_mode`` needs: ``[B,D,H,W]``, or ``[B,H,W]`` when
    ``attn_fuse_d`` is off)."""
    hypo, d = _check_k1_inputs(ref_nhwc, srcs_nhwc, rt, hypo)
    b, h, w, c = ref_nhwc.shape
    nsrc = len(srcs_nhwc)
    hs, ws = srcs_nhwc[0].shape[1:3]
    g = groups if group_cor else c
    out = torch.empty((b, g, d, h, w), device=ref_nhwc.device, dtype=torch.float32)
    wsum = None
    if want_wsum:
        wsum = torch.empty((b, d, h, w) if attn_fuse_d else (b, h, w), device=ref_nhwc.device, dtype=torch.float32)
    _lib.check(_lib.load().mvster_epi_fwd_mode_ex(
        _ptr(ref_nhwc), _ptr_array(srcs_nhwc), _ptr(rt), _ptr(hypo), _ptr(out), _ptr(wsum) if want_wsum else None,
        b, nsrc, c, g, d, h, w, hs, ws, float(attn_temp), _dtype_code(ref_nhwc), int(bool(group_cor)),
        int(bool(attn_fuse_d)), _stream(ref_nhwc)))
    return (out, wsum) if want_wsum else out


def epi_bwd_mode(ref_nhwc, srcs_nhwc, rt, hypo, out, wsum, gout, groups: int, attn_temp: float, group_cor: bool,
                 attn_fuse_d: bool) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """K1 backward of the non-default options.  Returns ``(grad_ref [B,H,W,C], [grad_src_v [B,Hs,Ws,C]])`` fp32."""
    hypo, d = _check_k1_inputs(ref_nhwc, srcs_nhwc, rt, hypo)
    b, h, w, c = ref_nhwc.shape
    nsrc = len(srcs_nhwc)
    hs, ws = srcs_nhwc[0].shape[1:3]
    g = groups if group_cor else c
    gout = _f32c(gout, "grad_output")
    want_ws = (b, d, h, w) if attn_fuse_d else (b, h, w)
    if tuple(gout.shape) != (b, g, d, h, w) or tuple(out.shape) != (b, g, d, h, w) or tuple(wsum.shape) != want_ws:
        raise RuntimeError("grad_output / out must be [B,G,D,H,W] and wsum %s of the forward call" % (want_ws,))
    dev = ref_nhwc.device
    grad_ref = torch.empty((b, h, w, c), device=dev, dtype=torch.float32)
    grad_all = torch.zeros((nsrc, b, hs, ws, c), device=dev, dtype=torch.float32)
    grad_srcs = [grad_all[v] for v in range(nsrc)]
    _lib.check(_lib.load().mvster_epi_bwd_mode(
        _ptr(ref_nhwc), _ptr_array(srcs_nhwc), _ptr(rt), _ptr(hypo), _ptr(out), _ptr(wsum), _ptr(gout),
        _ptr(grad_ref), _ptr_array(grad_srcs), b, nsrc, c, g, d, h, w, hs, ws, float(attn_temp),
        _dtype_code(ref_nhwc), int(bool(group_cor)), int(bool(attn_fuse_d)), _stream(ref_nhwc)))
    return grad_ref, grad_srcs


def epi_bwd(ref_nhwc, srcs_nhwc, rt, hypo, out, wsum, gout, groups: int, attn_temp: float
            ) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """K1 backward.  Returns ``(grad_ref [B,H,W,C] fp32, [grad_src_v [B,Hs,Ws,C] fp32])``."""
    hypo, d = _check_k1_inputs(ref_nhwc, srcs_nhwc, rt, hypo)
    b, h, w, c = ref_nhwc.shape
    nsrc = len(srcs_nhwc)
    hs, ws = srcs_nhwc[0].shape[1:3]
    gout = _f32c(gout, "grad_output")
    if tuple(gout.shape) != (b, groups, d, h, w) or tuple(out.shape) != (b, groups, d, h, w) or tuple(wsum.shape) != (b, d, h, w):
        raise RuntimeError("grad_output / out must be [B,G,D,H,W] and wsum [B,D,H,W] of the forward call")
    dev = ref_nhwc.device
    grad_ref = torch.empty((b, h, w, c), device=dev, dtype=torch.float32)
    grad_all = torch.zeros((nsrc, b, hs, ws, c), device=dev, dtype=torch.float32)  # one memset for all views
    grad_srcs = [grad_all[v] for v in range(nsrc)]
    _lib.check(_lib.load().mvster_epi_bwd(
        _ptr(ref_nhwc), _ptr_array(srcs_nhwc), _ptr(rt), _ptr(hypo), _ptr(out), _ptr(wsum), _ptr(gout),
        _ptr(grad_ref), _ptr_array(grad_srcs), b, nsrc, c, groups, d, h, w, hs, ws, float(attn_temp),
        _dtype_code(ref_nhwc), _stream(ref_nhwc)))
    return grad_ref, grad_srcs


def homo_warp(src_nhwc: torch.Tensor, rt: torch.Tensor, hypo: torch.Tensor) -> torch.Tensor:
    _require_cuda(src_nhwc, "src")
    if src_nhwc.dim() != 4 or not src_nhwc.is_contiguous():
        raise RuntimeError("src features must be contiguous NHWC [B,Hs,Ws,C]")
    b, hs, ws, c = src_nhwc.shape
    hypo = _f32c(hypo, "depth_values")
    if hypo.dim() != 4 or hypo.shape[0] != b or hypo.device != src_nhwc.device:
        raise RuntimeError("depth_values must be [B,D,H,W] with the features' batch size and device")
    if tuple(rt.shape) != (b, 12) or rt.dtype != torch.float32 or not rt.is_contiguous() or rt.device != src_nhwc.device:
        raise RuntimeError("rt must be contiguous fp32 [B,12] on the features' device")
    _, d, h, w = hypo.shape
    out = torch.empty((b, c, d, h, w), device=src_nhwc.device, dtype=torch.float32)
    _lib.check(_lib.load().mvster_homo_warp(_ptr(src_nhwc), _ptr(rt), _ptr(hypo), _ptr(out), b, c, d, h, w, hs, ws,
                                            _dtype_code(src_nhwc), _stream(src_nhwc)))
    return out


def init_inverse_range(depth_values: torch.Tensor, ndepths: int, h: int, w: int) -> torch.Tensor:
    dv = _f32c(depth_values, "depth_values")
    b, nv = dv.shape
    hypo = torch.empty((b, ndepths, h, w), device=dv.device, dtype=torch.float32)
    _lib.check(_lib.load().mvster_init_inverse_range(_ptr(dv), nv, _ptr(hypo), b, ndepths, h, w, _stream(dv)))
    return hypo


def schedule_inverse_range(inv_min: torch.Tensor, inv_max: torch.Tensor, ndepths: int, h: int, w: int) -> torch.Tensor:
    lo, hi = _f32c(inv_min, "inverse_min_depth"), _f32c(inv_max, "inverse_max_depth")
    b = lo.shape[0]
    if tuple(lo.shape) != (b, h // 2, w // 2) or lo.shape != hi.shape:
        raise RuntimeError("inverse_min/max_depth must be [B,H//2,W//2] = [%d,%d,%d], got %s / %s"
                           % (b, h // 2, w // 2, tuple(lo.shape), tuple(hi.shape)))
    hypo = torch.empty((b, ndepths, h, w), device=lo.device, dtype=torch.float32)
    _lib.check(_lib.load().mvster_schedule_inverse_range(_ptr(lo), _ptr(hi), _ptr(hypo), b, ndepths, h, w, _stream(lo)))
    return hypo


def tail(logits: torch.Tensor, hypo: torch.Tensor, split_itv: float, want_conf: bool, inverse_depth: bool,
         depth_mode: int = DEPTH_ARGMAX):
    """K2a.  Returns ``(attn [B,D,H,W], depth [B,H,W], conf or None, inv_min or None, inv_max or None)``."""
    lg, hy = _f32c(logits, "logits"), _f32c(hypo, "depth_hypo")
    if lg.shape != hy.shape or lg.dim() != 4:
        raise RuntimeError("logits and depth_hypo must both be [B,D,H,W], got %s / %s" % (tuple(lg.shape), tuple(hy.shape)))
    b, d, h, w = lg.shape
    dev = lg.device
    attn = torch.empty_like(lg)
    depth = torch.empty((b, h, w), device=dev, dtype=torch.float32)
    conf = torch.empty((b, h, w), device=dev, dtype=torch.float32) if want_conf else None
    inv_min = torch.empty((b, h, w), device=dev, dtype=torch.float32) if inverse_depth else None
    inv_max = torch.empty((b, h, w), device=dev, dtype=torch.float32) if inverse_depth else None
    _lib.check(_lib.load().mvster_tail(_ptr(lg), _ptr(hy), float(split_itv), int(depth_mode), _ptr(attn), _ptr(depth),
                                       _ptr(conf), _ptr(inv_min), _ptr(inv_max), b, d, h, w, _stream(lg)))
    return attn, depth, conf, inv_min, inv_max


def regtail(low: torch.Tensor, skip: torch.Tensor, w_host: torch.Tensor, params_host: torch.Tensor,
            hypo: torch.Tensor, split_itv: float, inverse_depth: bool, depth_mode: int = DEPTH_ARGMAX,
            want_conf: bool = True):
    """K2a': ``skip + relu(bn(conv11(low)))`` -> ``prob`` -> tail in one kernel (reference mvs4net_utils.py:923-926 +
    :1109-1156, eval mode).  ``low`` [B,16,D,H/2,W/2] and ``skip`` [B,8,D,H,W] are CUDA fp32 NCDHW; ``w_host``
    [3,3,16,8] (BatchNorm-folded transposed-conv weight, ``[ky][kx][ci][co]``) and ``params_host`` [17] (BN shift,
    prob weight, prob bias) are CPU fp32 tensors - they are passed to the kernel by value.
    Returns ``(attn, depth, conf, inv_min, inv_max)`` like :func:`tail`."""
    _require_cuda(low, "low")
    low, skip, hy = _f32c(low, "low"), _f32c(skip, "skip"), _f32c(hypo, "depth_hypo")
    if low.dim() != 5 or skip.dim() != 5 or low.shape[1] != 16 or skip.shape[1] != 8:
        raise RuntimeError("regtail: low must be [B,16,D,H/2,W/2] and skip [B,8,D,H,W], got %s / %s"
                           % (tuple(low.shape), tuple(skip.shape)))
    b, _, d, h, w = skip.shape
    if tuple(low.shape) != (b, 16, d, h // 2, w // 2) or tuple(hy.shape) != (b, d, h, w) or h % 2 or w % 2:
        raise RuntimeError("regtail: inconsistent shapes low %s skip %s hypo %s"
                           % (tuple(low.shape), tuple(skip.shape), tuple(hy.shape)))
    if w_host.device.type != "cpu" or params_host.device.type != "cpu" or w_host.dtype != torch.float32 \
            or params_host.dtype != torch.float32 or w_host.numel() != 1152 or params_host.numel() != 17 \
            or not w_host.is_contiguous() or not params_host.is_contiguous():
        raise RuntimeError("regtail: w_host [3,3,16,8] and params_host [17] must be contiguous CPU fp32 tensors")
    dev = skip.device
    attn = torch.empty((b, d, h, w), device=dev, dtype=torch.float32)
    depth = torch.empty((b, h, w), device=dev, dtype=torch.float32)
    conf = torch.empty((b, h, w), device=dev, dtype=torch.float32) if want_conf else None
    inv_min = torch.empty((b, h, w), device=dev, dtype=torch.float32) if inverse_depth else None
    inv_max = torch.empty((b, h, w), device=dev, dtype=torch.float32) if inverse_depth else None
    _lib.check(_lib.load().mvster_regtail(
        _ptr(low), _ptr(skip), ctypes.c_void_p(w_host.data_ptr()), ctypes.c_void_p(params_host.data_ptr()), _ptr(hy),
        float(split_itv), int(depth_mode), _ptr(attn), _ptr(depth), _ptr(conf), _ptr(inv_min), _ptr(inv_max),
        b, d, h, w, _stream(skip)))
    return attn, depth, conf, inv_min, inv_max


CONV_STRIDE1, CONV_STRIDE2, CONV_TRANSPOSED2 = 0, 1, 2
_SMALL_CONVS = {(4, 8, 1, 0), (8, 8, 1, 0), (8, 16, 1, 1), (16, 16, 3, 0), (16, 32, 1, 1), (32, 16, 1, 2), (16, 8, 1, 2)}


def conv3d_small_supported(cin: int, cout: int, kd: int, mode: int, h: int, w: int) -> bool:
    """True when ``mvster_conv3d_small`` has a kernel for this layer at input size ``h x w``."""
    if (cin, cout, kd, mode) not in _SMALL_CONVS:
        return False
    if mode == CONV_STRIDE1:
        return h % 2 == 0 and w % 2 == 0
    if mode == CONV_STRIDE2:
        return h % 2 == 0 and w % 4 == 0
    return True


def conv3d_small(x: torch.Tensor, w_host: torch.Tensor, bias_host: torch.Tensor, mode: int, relu: bool = True,
                 skip: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Direct fp32 convolution of a few-channel NCDHW volume (reference ConvBnReLU3D / ConvTranspose3d+BN+ReLU blocks
    of reg2d, mvs4net_utils.py:889-912, BatchNorm folded).  ``w_host`` [kd,3,3,Cin,Cout] and ``bias_host`` [Cout] are
    CPU fp32 tensors (they travel as kernel parameters)."""
    _require_cuda(x, "x")
    x = _f32c(x, "x")
    if x.dim() != 5 or w_host.dim() != 5 or w_host.device.type != "cpu" or bias_host.device.type != "cpu" \
            or w_host.dtype != torch.float32 or bias_host.dtype != torch.float32 \
            or not w_host.is_contiguous() or not bias_host.is_contiguous():
        raise RuntimeError("conv3d_small: x must be [B,Cin,D,H,W]; w_host [kd,3,3,Cin,Cout] / bias_host [Cout] "
                           "contiguous CPU fp32")
    b, cin, d, h, w = x.shape
    kd, _, _, wci, cout = w_host.shape
    if wci != cin or bias_host.numel() != cout:
        raise RuntimeError("conv3d_small: weight %s does not match input channels %d" % (tuple(w_host.shape), cin))
    oh, ow = (h, w) if mode == CONV_STRIDE1 else ((h // 2, w // 2) if mode == CONV_STRIDE2 else (2 * h, 2 * w))
    y = torch.empty((b, cout, d, oh, ow), device=x.device, dtype=torch.float32)
    if skip is not None:
        skip = _f32c(skip, "skip")
        if skip.shape != y.shape:
            raise RuntimeError("conv3d_small: skip %s does not match the output %s" % (tuple(skip.shape), tuple(y.shape)))
    _lib.check(_lib.load().mvster_conv3d_small(
        _ptr(x), ctypes.c_void_p(w_host.data_ptr()), ctypes.c_void_p(bias_host.data_ptr()), _ptr(skip), _ptr(y),
        b, cin, cout, d, h, w, int(kd), int(mode), int(bool(relu)), _stream(x)))
    return y


_CONV3D_SLICE = {(32, 64, CONV_STRIDE2): 16, (64, 32, CONV_TRANSPOSED2): 8}  # (Cin, Cout_total, mode) -> Cout per launch


def conv3d_sliced_supported(cin: int, cout: int, kd: int, mode: int, h: int, w: int) -> bool:
    if kd != 1 or (cin, cout, mode) not in _CONV3D_SLICE:
        return False
    return mode == CONV_TRANSPOSED2 or (h % 2 == 0 and w % 4 == 0)


def conv3d_sliced(x: torch.Tensor, w_slices, bias_slices, mode: int, relu: bool = True,
                  skip: Optional[torch.Tensor] = None) -> torch.Tensor:
    """reg2d.conv5 / conv7 as one launch per filter-bank slice (``mvster_conv3d_small_slice``).  ``w_slices`` /
    ``bias_slices``: per-slice contiguous CPU fp32 tensors ``[1,3,3,Cin,cs]`` / ``[cs]``."""
    x = _f32c(x, "x")
    b, cin, d, h, w = x.shape
    cs = w_slices[0].shape[-1]
    cout = cs * len(w_slices)
    oh, ow = (h // 2, w // 2) if mode == CONV_STRIDE2 else (2 * h, 2 * w)
    y = torch.empty((b, cout, d, oh, ow), device=x.device, dtype=torch.float32)
    if skip is not None:
        skip = _f32c(skip, "skip")
        if skip.shape != y.shape:
            raise RuntimeError("conv3d_sliced: skip %s does not match the output %s" % (tuple(skip.shape), tuple(y.shape)))
    lib = _lib.load()
    for i, (ws, bs) in enumerate(zip(w_slices, bias_slices)):
        if ws.device.type != "cpu" or not ws.is_contiguous() or ws.dtype != torch.float32 or ws.shape[3] != cin:
            raise RuntimeError("conv3d_sliced: weight slices must be contiguous CPU fp32 [1,3,3,Cin,cs]")
        _lib.check(lib.mvster_conv3d_small_slice(
            _ptr(x), ctypes.c_void_p(ws.data_ptr()), ctypes.c_void_p(bs.data_ptr()), _ptr(skip), _ptr(y), b, cin, cs, cout,
            i * cs, d, h, w, int(mode), int(bool(relu)), _stream(x)))
    return y


def conv2d_mid5_supported(cin: int, cout: int, h: int, w: int) -> bool:
    return cin in (8, 16, 32) and cout % 16 == 0 and h % 4 == 0 and w % 4 == 0


def conv2d_mid5(x: torch.Tensor, w_dev: torch.Tensor, bias_dev: torch.Tensor, relu: bool = True) -> torch.Tensor:
    """5x5 stride-2 convolution of FPN4's down-sampling layers, folded weights ``[5,5,Cin,Cout]`` resident on the device."""
    x = _f32c(x, "x")
    _require_cuda(w_dev, "w_dev")
    _require_cuda(bias_dev, "bias_dev")
    b, cin, h, w = x.shape
    if w_dev.dim() != 4 or tuple(w_dev.shape[:3]) != (5, 5, cin) or not w_dev.is_contiguous() or w_dev.dtype != torch.float32 \
            or bias_dev.numel() != w_dev.shape[3] or bias_dev.dtype != torch.float32:
        raise RuntimeError("conv2d_mid5: w_dev must be contiguous fp32 [5,5,Cin,Cout] and bias_dev [Cout]")
    cout = w_dev.shape[3]
    y = torch.empty((b, cout, h // 2, w // 2), device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().mvster_conv2d_mid5(_ptr(x), _ptr(w_dev), _ptr(bias_dev), _ptr(y), b, cin, cout, h, w,
                                              int(bool(relu)), _stream(x)))
    return y


def conv3d_mid_supported(cin: int, cout: int, kd: int, h: int, w: int) -> bool:
    """True when ``mvster_conv3d_mid`` has a kernel for this stride-1 layer at input size ``h x w``."""
    return cin in (32, 64) and cout % 16 == 0 and kd in (1, 3) and h % 2 == 0 and w % 2 == 0


_TC_SHAPES = {(1, 32, 32), (1, 64, 64), (1, 64, 32), (3, 32, 32), (3, 64, 64)}  # (kd, Cin, Cout) of mvster_conv3d_mid_tc
_TF32_SPLIT_CACHE: dict = {}   # id(weight tensor) -> (weakref, version, hi, lo)


def conv3d_mid_tc_mode() -> int:
    """Kernel family of the 32/64-channel layers (``MVSTER_MID_TC``): 2 = tcgen05 (UMMA, tensor memory), 1 = mma.sync,
    0 = FP32 SIMT.  All three are fp32-grade (the tensor-core ones by a 3xTF32 operand split)."""
    return int(os.environ.get("MVSTER_MID_TC", "2"))


def conv3d_mid_tc_enabled() -> bool:
    return conv3d_mid_tc_mode() != 0


_UMMA_PACK_CACHE: dict = {}   # id(weight tensor) -> (weakref, version, packed)


def umma_pack_weights(w_dev: torch.Tensor) -> torch.Tensor:
    """Folded weights [kd,3,3,Cin,Cout] -> the per-chunk hi / lo blocks ``mvster_conv3d_mid_umma`` copies in bulk."""
    _require_cuda(w_dev, "w_dev")
    hit = _UMMA_PACK_CACHE.get(id(w_dev))
    if hit is not None and hit[0]() is w_dev and hit[1] == w_dev._version:
        return hit[2]
    kd, _, _, cin, cout = w_dev.shape
    packed = torch.empty(2 * w_dev.numel(), device=w_dev.device, dtype=torch.float32)
    _lib.check(_lib.load().mvster_umma_pack_weights(_ptr(w_dev), _ptr(packed), int(kd), int(cin), int(cout), _stream(w_dev)))
    if len(_UMMA_PACK_CACHE) > 256:
        for k in [k for k, v in _UMMA_PACK_CACHE.items() if v[0]() is None]:
            del _UMMA_PACK_CACHE[k]
    _UMMA_PACK_CACHE[id(w_dev)] = (weakref.ref(w_dev), w_dev._version, packed)
    return packed


def tf32_split(w_dev: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """``(hi, lo)`` with ``hi = rna_tf32(w)``, ``lo = rna_tf32(w - hi)``, cached per weight tensor and version."""
    _require_cuda(w_dev, "w_dev")
    key = id(w_dev)
    hit = _TF32_SPLIT_CACHE.get(key)
    if hit is not None and hit[0]() is w_dev and hit[1] == w_dev._version:
        return hit[2], hit[3]
    hi, lo = torch.empty_like(w_dev), torch.empty_like(w_dev)
    _lib.check(_lib.load().mvster_tf32_split(_ptr(w_dev), _ptr(hi), _ptr(lo), w_dev.numel(), _stream(w_dev)))
    if len(_TF32_SPLIT_CACHE) > 256:
        for k in [k for k, v in _TF32_SPLIT_CACHE.items() if v[0]() is None]:
            del _TF32_SPLIT_CACHE[k]
    _TF32_SPLIT_CACHE[key] = (weakref.ref(w_dev), w_dev._version, hi, lo)
    return hi, lo


def conv3d_mid(x: torch.Tensor, w_dev: torch.Tensor, bias_dev: torch.Tensor, relu: bool = True,
               tensor_cores=None) -> torch.Tensor:
    """fp32 stride-1 convolution of a 32/64-channel NCDHW volume with folded weights resident on the device
    (reg2d conv4 / conv6, FPN4 conv3.1 / conv3.2 / out2 with D = 1).  ``w_dev`` [kd,3,3,Cin,Cout], ``bias_dev`` [Cout].
    Shapes compiled for ``mvster_conv3d_mid_tc`` run as 3xTF32 implicit GEMMs on the tensor cores (fp32-grade accuracy;
    ``tensor_cores=False`` or ``MVSTER_MID_TC=0``: the FP32 SIMT kernel), the others on the SIMT kernel (even H, W)."""
    x = _f32c(x, "x")
    _require_cuda(w_dev, "w_dev")
    _require_cuda(bias_dev, "bias_dev")
    if x.dim() != 5 or w_dev.dim() != 5 or not w_dev.is_contiguous() or w_dev.dtype != torch.float32 \
            or bias_dev.dtype != torch.float32 or not bias_dev.is_contiguous():
        raise RuntimeError("conv3d_mid: x must be [B,Cin,D,H,W]; w_dev [kd,3,3,Cin,Cout] / bias_dev [Cout] contiguous fp32")
    b, cin, d, h, w = x.shape
    kd, _, _, wci, cout = w_dev.shape
    if wci != cin or bias_dev.numel() != cout:
        raise RuntimeError("conv3d_mid: weight %s does not match input channels %d" % (tuple(w_dev.shape), cin))
    y = torch.empty((b, cout, d, h, w), device=x.device, dtype=torch.float32)
    mode = conv3d_mid_tc_mode() if tensor_cores is None else int(tensor_cores)   # 0 SIMT, 1 / True mma.sync, 2 tcgen05
    if mode == 2 and (int(kd), cin, cout) in _TC_SHAPES:
        packed = umma_pack_weights(w_dev)
        _lib.check(_lib.load().mvster_conv3d_mid_umma(_ptr(x), _ptr(packed), _ptr(bias_dev), _ptr(y), b, cin, cout, d, h, w,
                                                      int(kd), int(bool(relu)), _stream(x)))
        return y
    if mode == 1 and (int(kd), cin, cout) in _TC_SHAPES:
        hi, lo = tf32_split(w_dev)
        _lib.check(_lib.load().mvster_conv3d_mid_tc(_ptr(x), _ptr(hi), _ptr(lo), _ptr(bias_dev), _ptr(y), b, cin, cout, d, h,
                                                    w, int(kd), int(bool(relu)), _stream(x)))
        return y
    _lib.check(_lib.load().mvster_conv3d_mid(_ptr(x), _ptr(w_dev), _ptr(bias_dev), _ptr(y), b, cin, cout, d, h, w, int(kd),
                                             int(bool(relu)), _stream(x)))
    return y


_CONV2D_SLICE = {(3, 3): 8, (8, 3): 8, (16, 3): 16, (32, 3): 16, (8, 5): 16, (16, 5): 16}  # (Cin, k) -> Cout per launch


def conv2d_small_supported(cin: int, cout: int, ksize: int, stride: int, h: int, w: int) -> bool:
    cs = _CONV2D_SLICE.get((cin, ksize))
    if cs is None or cout % cs or (ksize, stride) not in ((3, 1), (5, 2)):
        return False
    return (h % 2 == 0 and w % 2 == 0) if stride == 1 else (h % 2 == 0 and w % 4 == 0)


def conv2d_small(x: torch.Tensor, w_slices: Sequence[torch.Tensor], b_slices: Sequence[torch.Tensor], ksize: int,
                 stride: int, relu: bool = True) -> torch.Tensor:
    """Conv2d + folded BatchNorm + ReLU of an FPN4 encoder block (mvs4net_utils.py:231-258) on NCHW planar fp32.
    ``w_slices[i]`` [k,k,Cin,cs] / ``b_slices[i]`` [cs] are CPU fp32 tensors, one per slice of ``cs`` output channels."""
    _require_cuda(x, "x")
    x = _f32c(x, "x")
    b, cin, h, w = x.shape
    cs = w_slices[0].shape[3]
    cout = cs * len(w_slices)
    oh, ow = (h, w) if stride == 1 else (h // 2, w // 2)
    y = torch.empty((b, cout, oh, ow), device=x.device, dtype=torch.float32)
    lib = _lib.load()
    for i, (ws, bs) in enumerate(zip(w_slices, b_slices)):
        if ws.device.type != "cpu" or ws.dtype != torch.float32 or not ws.is_contiguous() \
                or tuple(ws.shape) != (ksize, ksize, cin, cs) or bs.numel() != cs or not bs.is_contiguous():
            raise RuntimeError("conv2d_small: weight slices must be contiguous CPU fp32 [k,k,Cin,cs]")
        _lib.check(lib.mvster_conv2d_small(_ptr(x), ctypes.c_void_p(ws.data_ptr()), ctypes.c_void_p(bs.data_ptr()),
                                           _ptr(y), b, cin, cs, cout, i * cs, h, w, int(ksize), int(stride),
                                           int(bool(relu)), _stream(x)))
    return y


def fpn_topdown(prev: torch.Tensor, lat: torch.Tensor, w_out_slices: Sequence[torch.Tensor], w_in_host: torch.Tensor,
                b_in_host: torch.Tensor, want_intra: bool, feature_dtype: Optional[torch.dtype] = None):
    """One FPN4 top-down level (mvs4net_utils.py:488-495): ``intra = up2(prev) + inner(lat); feat = out_conv(intra)``.
    ``prev`` [B,64,H/2,W/2], ``lat`` [B,Clat,H,W] planar CUDA fp32; ``w_out_slices[i]`` [3,3,64,8] CPU.
    Returns ``(feat NHWC [B,H,W,8*len(slices)], intra [B,64,H,W] or None)``; ``intra`` is materialised only when asked
    for or when a second output-channel slice has to reload it.  ``feature_dtype=torch.bfloat16`` makes the kernel emit
    the bf16 feature map K1 gathers from (rounded once from the fp32 accumulators; no cast pass)."""
    _require_cuda(prev, "prev")
    feature_dtype = feature_dtype or torch.float32
    if feature_dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError("fpn_topdown: feature dtype must be float32 or bfloat16")
    prev, lat = _f32c(prev, "prev"), _f32c(lat, "lat")
    b, clat, h, w = lat.shape
    if tuple(prev.shape) != (b, 64, h // 2, w // 2) or h % 2 or w % 2:
        raise RuntimeError("fpn_topdown: prev %s does not match lat %s" % (tuple(prev.shape), tuple(lat.shape)))
    ns = len(w_out_slices)
    cout = 8 * ns
    feat = torch.empty((b, h, w, cout), device=lat.device, dtype=feature_dtype)
    intra = torch.empty((b, 64, h, w), device=lat.device, dtype=torch.float32) if (want_intra or ns > 1) else None
    lib = _lib.load()
    for t in (w_in_host, b_in_host) + tuple(w_out_slices):
        if t.device.type != "cpu" or t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError("fpn_topdown: weights must be contiguous CPU fp32 tensors")
    if tuple(w_in_host.shape) != (clat, 64) or b_in_host.numel() != 64:
        raise RuntimeError("fpn_topdown: w_in must be [Clat,64], b_in [64]")
    for i, ws in enumerate(w_out_slices):
        if tuple(ws.shape) != (3, 3, 64, 8):
            raise RuntimeError("fpn_topdown: output-conv slices must be [3,3,64,8]")
        first = i == 0
        _lib.check(lib.mvster_fpn_topdown_ex(
            _ptr(prev) if first else None, _ptr(lat) if first else None, None if first else _ptr(intra),
            _ptr(intra) if first else None, _ptr(feat), _dtype_code(feat), ctypes.c_void_p(ws.data_ptr()),
            ctypes.c_void_p(w_in_host.data_ptr()), ctypes.c_void_p(b_in_host.data_ptr()), b, clat, 8, cout, 8 * i, h, w,
            _stream(lat)))
    return feat, intra


def _matmul_fp32(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """cuBLAS fp32 GEMM with TF32 off whatever the process-wide switch says (the FPN parity bounds are fp32 ones)."""
    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        return torch.matmul(a, b)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32


def fpn_project(intra: torch.Tensor, wp_t: torch.Tensor) -> torch.Tensor:
    """``[B,64,h,w]`` planar -> ``[B,h,w,N]`` NHWC tap-wise projection ``intra^T wp_t`` (``wp_t`` [64,N] on the device):
    the plain library GEMM in front of :func:`fpn_lin_gather` / :func:`fpn_project_up`."""
    _require_cuda(intra, "intra")
    intra = _f32c(intra, "intra")
    b, c, h, w = intra.shape
    if wp_t.dim() != 2 or wp_t.shape[0] != c or wp_t.device != intra.device or wp_t.dtype != torch.float32:
        raise RuntimeError("fpn_project: wp_t must be fp32 [%d, N] on the features' device" % c)
    return _matmul_fp32(intra.flatten(2).transpose(1, 2), wp_t).view(b, h, w, wp_t.shape[1])


def fpn_lin_gather(proj: torch.Tensor, p_off: int, lat: torch.Tensor, wc_host: torch.Tensor, bc_host: torch.Tensor,
                   feature_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """``mvster_fpn_topdown_lin``: ``proj`` [B,H/2,W/2,Nc] NHWC fp32 (channels ``p_off .. p_off+9*Cout-1`` are this
    level's tap-wise projection), ``lat`` [B,Clat,H,W] planar, ``wc_host`` [9,Clat,Cout], ``bc_host`` [9,Cout] CPU.
    Returns ``feat`` NHWC [B,H,W,Cout] in ``feature_dtype`` (fp32 / bf16)."""
    _require_cuda(lat, "lat")
    feature_dtype = feature_dtype or torch.float32
    if feature_dtype not in (torch.float32, torch.bfloat16):
        raise RuntimeError("fpn_topdown_lin: feature dtype must be float32 or bfloat16")
    lat = _f32c(lat, "lat")
    b, clat, h, w = lat.shape
    if wc_host.dim() != 3 or bc_host.dim() != 2:
        raise RuntimeError("fpn_topdown_lin: wc / bc must be [9,Clat,Cout] / [9,Cout]")
    cout = wc_host.shape[2]
    for t, shape in ((wc_host, (9, clat, cout)), (bc_host, (9, cout))):
        if t.device.type != "cpu" or t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != shape:
            raise RuntimeError("fpn_topdown_lin: wc / bc must be contiguous CPU fp32 [9,Clat,Cout] / [9,Cout]")
    if h % 2 or w % 2 or h < 2 or w < 2 or proj.dim() != 4 or tuple(proj.shape[:3]) != (b, h // 2, w // 2) \
            or proj.dtype != torch.float32 or proj.device != lat.device or not proj.is_contiguous():
        raise RuntimeError("fpn_topdown_lin: proj %s must be contiguous fp32 [B,H/2,W/2,N] matching lat %s"
                           % (tuple(proj.shape), tuple(lat.shape)))
    feat = torch.empty((b, h, w, cout), device=lat.device, dtype=feature_dtype)
    _lib.check(_lib.load().mvster_fpn_topdown_lin(
        _ptr(proj), int(proj.shape[3]), int(p_off), _ptr(lat), _ptr(feat), _dtype_code(feat),
        ctypes.c_void_p(wc_host.data_ptr()), ctypes.c_void_p(bc_host.data_ptr()), b, clat, cout, h, w, _stream(lat)))
    return feat


def fpn_project_up(q: torch.Tensor, q_off: int, n_proj: int, lat: torch.Tensor, wl_host: torch.Tensor,
                   bl_host: torch.Tensor) -> torch.Tensor:
    """``mvster_fpn_project_up``: the projection of ``up2(prev) + inner(lat)`` from ``q = project(prev)`` (NHWC
    [B,H/2,W/2,Nq], channels ``q_off .. q_off+n_proj-1``): returns NHWC [B,H,W,n_proj] fp32."""
    _require_cuda(lat, "lat")
    lat = _f32c(lat, "lat")
    b, clat, h, w = lat.shape
    if h % 2 or w % 2 or h < 2 or w < 2 or q.dim() != 4 or tuple(q.shape[:3]) != (b, h // 2, w // 2) \
            or q.dtype != torch.float32 or q.device != lat.device or not q.is_contiguous():
        raise RuntimeError("fpn_project_up: q %s must be contiguous fp32 [B,H/2,W/2,N] matching lat %s"
                           % (tuple(q.shape), tuple(lat.shape)))
    for t, shape in ((wl_host, (clat, n_proj)), (bl_host, (n_proj,))):
        if t.device.type != "cpu" or t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != shape:
            raise RuntimeError("fpn_project_up: wl / bl must be contiguous CPU fp32 [Clat,N] / [N]")
    out = torch.empty((b, h, w, n_proj), device=lat.device, dtype=torch.float32)
    _lib.check(_lib.load().mvster_fpn_project_up(
        _ptr(q), int(q.shape[3]), int(q_off), _ptr(lat), _ptr(out), ctypes.c_void_p(wl_host.data_ptr()),
        ctypes.c_void_p(bl_host.data_ptr()), b, clat, int(n_proj), h, w, _stream(lat)))
    return out


def fpn_topdown_lin(prev: torch.Tensor, lat: torch.Tensor, wp_t: torch.Tensor, wc_host: torch.Tensor,
                    bc_host: torch.Tensor, feature_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """One FPN4 top-down level through its linearity, for a level whose ``intra`` no finer level needs:
    ``fpn_project`` (one fp32 cuBLAS GEMM, 64 -> 9*Cout at half resolution) + ``fpn_lin_gather``.
    ``prev`` [B,64,H/2,W/2], ``lat`` [B,Clat,H,W] planar CUDA fp32; ``wp_t`` [64, 9*Cout] CUDA
    (``wp_t[c, tap*Cout+co] = out_conv.weight[co,c,tap]``); ``wc_host`` [9,Clat,Cout], ``bc_host`` [9,Cout] CPU fp32.
    Returns ``feat`` NHWC [B,H,W,Cout]."""
    b, clat, h, w = lat.shape
    if tuple(prev.shape) != (b, 64, h // 2, w // 2):
        raise RuntimeError("fpn_topdown_lin: prev %s does not match lat %s" % (tuple(prev.shape), tuple(lat.shape)))
    return fpn_lin_gather(fpn_project(prev, wp_t), 0, lat, wc_host, bc_host, feature_dtype)


def bn_train_supported(x: torch.Tensor) -> bool:
    """Planar fp32 CUDA activations ``[N, C, ...]`` whose (sample, channel) plane is a multiple of four elements."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() >= 3 and x.is_contiguous()):
        return False
    plane = x[0, 0].numel()
    return plane > 0 and plane % 4 == 0 and x.shape[0] * x.shape[1] <= 65535 and x.data_ptr() % 16 == 0


def _bn_workspace(n: int, c: int, s: int, dev) -> torch.Tensor:
    nbytes = int(_lib.load().mvster_bn_train_workspace_bytes(n, c, s))
    return torch.empty(((nbytes + 7) // 8,), device=dev, dtype=torch.float64)


def bn_train_fwd(x: torch.Tensor, weight, bias, running_mean, running_var, momentum: float, eps: float, relu: bool):
    """Training-mode BatchNorm (+ ReLU) forward (``mvster_bn_train_fwd``).  Returns ``(y, mean [C], invstd [C])`` and
    updates ``running_mean`` / ``running_var`` in place like ``F.batch_norm(training=True)``."""
    if not bn_train_supported(x):
        raise RuntimeError("bn_train_fwd: needs a contiguous planar fp32 CUDA tensor [N,C,...] with a plane size % 4 == 0")
    n, c = x.shape[0], x.shape[1]
    s = x[0, 0].numel()
    y = torch.empty_like(x)
    mean = torch.empty((c,), device=x.device, dtype=torch.float32)
    invstd = torch.empty_like(mean)
    for t in (weight, bias, running_mean, running_var):
        if t is not None and (t.device != x.device or t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != c):
            raise RuntimeError("bn_train_fwd: weight / bias / running statistics must be contiguous fp32 [C] on x's device")
    ws = _bn_workspace(n, c, s, x.device)
    _lib.check(_lib.load().mvster_bn_train_fwd(
        _ptr(x), _ptr(weight), _ptr(bias), _ptr(y), _ptr(mean), _ptr(invstd), _ptr(running_mean), _ptr(running_var),
        float(momentum), float(eps), int(bool(relu)), n, c, s, _ptr(ws), _stream(x)))
    return y, mean, invstd


def bn_train_bwd(x: torch.Tensor, y: torch.Tensor, dy: torch.Tensor, weight, mean: torch.Tensor, invstd: torch.Tensor,
                 relu: bool):
    """Training-mode BatchNorm (+ ReLU) backward (``mvster_bn_train_bwd``).  Returns ``(dx, dweight [C], dbias [C])``."""
    n, c = x.shape[0], x.shape[1]
    s = x[0, 0].numel()
    dy = _f32c(dy, "grad_output")
    if dy.shape != x.shape or dy.data_ptr() % 16:
        raise RuntimeError("bn_train_bwd: grad_output must match x and be 16-byte aligned")
    dx = torch.empty_like(x)
    dw = torch.empty((c,), device=x.device, dtype=torch.float32)
    db = torch.empty_like(dw)
    ws = _bn_workspace(n, c, s, x.device)
    _lib.check(_lib.load().mvster_bn_train_bwd(
        _ptr(x), _ptr(y), _ptr(dy), _ptr(weight), _ptr(mean), _ptr(invstd), _ptr(dx), _ptr(dw), _ptr(db),
        int(bool(relu)), n, c, s, _ptr(ws), _stream(x)))
    return dx, dw, db


def conv3d_wgrad_supported(a: torch.Tensor, b: torch.Tensor, kd: int, stride: int) -> bool:
    """Planar fp32 CUDA ``[N,C,D,H,W]`` pair of ``conv3d_wgrad`` (A channels % 8 == 0, matching spatial sizes)."""
    if not (a.is_cuda and b.is_cuda and a.dtype == b.dtype == torch.float32 and a.dim() == b.dim() == 5):
        return False
    if kd not in (1, 3) or stride not in (1, 2) or a.shape[1] % 8 or a.shape[0] != b.shape[0] or a.shape[2] != b.shape[2]:
        return False
    ha, wa, hb, wb = a.shape[3], a.shape[4], b.shape[3], b.shape[4]
    if stride == 1:
        return (ha, wa) == (hb, wb)
    return ((hb + 1) // 2, (wb + 1) // 2) == (ha, wa)


def conv3d_wgrad(a: torch.Tensor, b: torch.Tensor, kd: int, stride: int) -> torch.Tensor:
    """``mvster_conv3d_wgrad``: ``dW [CA,CB,kd,3,3] = sum A[n,a,d,y,x] B[n,b,d+kd-kd//2, s*y+ky-1, s*x+kx-1]``.
    Conv3d: ``a = grad_output``, ``b = input``; ConvTranspose3d (stride 2): ``a = input``, ``b = grad_output``."""
    a, b = _f32c(a, "A"), _f32c(b, "B")
    if not conv3d_wgrad_supported(a, b, kd, stride):
        raise RuntimeError("conv3d_wgrad: unsupported shapes A %s B %s kd=%d stride=%d" % (tuple(a.shape), tuple(b.shape), kd, stride))
    n, ca, d, ha, wa = a.shape
    cb, hb, wb = b.shape[1], b.shape[3], b.shape[4]
    lib = _lib.load()
    ws = torch.empty(((int(lib.mvster_conv3d_wgrad_workspace_bytes(n, ca, cb, kd, d, ha, wa)) + 3) // 4,), device=a.device,
                     dtype=torch.float32)
    dw = torch.empty((ca, cb, kd, 3, 3), device=a.device, dtype=torch.float32)
    _lib.check(lib.mvster_conv3d_wgrad(_ptr(a), _ptr(b), _ptr(dw), n, ca, cb, kd, d, ha, wa, hb, wb, stride, _ptr(ws),
                                       _stream(a)))
    return dw


def conv1x1_wgrad(x: torch.Tensor, g: torch.Tensor):
    """``mvster_conv1x1_wgrad``: ``(dw [C], db [1])`` of a C -> 1 pointwise convolution, ``x`` [N,C,...], ``g`` [N,1,...]."""
    x, g = _f32c(x, "x"), _f32c(g, "grad_output")
    n, c = x.shape[0], x.shape[1]
    s = x[0, 0].numel()
    if g.shape[0] != n or g.shape[1] != 1 or g[0, 0].numel() != s:
        raise RuntimeError("conv1x1_wgrad: grad_output %s does not match x %s" % (tuple(g.shape), tuple(x.shape)))
    lib = _lib.load()
    ws = torch.empty(((int(lib.mvster_conv1x1_wgrad_workspace_bytes(n, c, s)) + 7) // 8,), device=x.device, dtype=torch.float64)
    dw = torch.empty((c,), device=x.device, dtype=torch.float32)
    db = torch.empty((1,), device=x.device, dtype=torch.float32)
    _lib.check(lib.mvster_conv1x1_wgrad(_ptr(x), _ptr(g), _ptr(dw), _ptr(db), n, c, s, _ptr(ws), _stream(x)))
    return dw, db


def tail_bwd(attn, hypo, depth, g_attn, g_depth, depth_mode: int) -> torch.Tensor:
    b, d, h, w = attn.shape
    g_attn = None if g_attn is None else _f32c(g_attn, "grad attn")
    g_depth = None if g_depth is None else _f32c(g_depth, "grad depth")
    g_logits = torch.empty_like(attn)
    _lib.check(_lib.load().mvster_tail_bwd(_ptr(attn), _ptr(hypo), _ptr(depth), _ptr(g_attn), _ptr(g_depth),
                                           int(depth_mode), _ptr(g_logits), b, d, h, w, _stream(attn)))
    return g_logits


def _dbl(a, n) -> "ctypes.Array":
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
    if arr.size != n:
        raise RuntimeError("expected %d camera values, got %d" % (n, arr.size))
    return arr


def depth2pts(depth: torch.Tensor, k, e) -> torch.Tensor:
    """World points [H*W,3] float64 of a depth map (reference depth2pts_np, test_mvs4.py:206-218)."""
    dz = _f32c(depth, "depth")
    h, w = dz.shape
    xyz = torch.empty((h * w, 3), device=dz.device, dtype=torch.float64)
    dp = ctypes.POINTER(ctypes.c_double)
    kk, ee = _dbl(k, 9), _dbl(e, 16)
    _lib.check(_lib.load().mvster_depth2pts(_ptr(dz), kk.ctypes.data_as(dp), ee.ctypes.data_as(dp), _ptr(xyz), h, w,
                                            _stream(dz)))
    return xyz


def geo_check_pair(depth_ref: torch.Tensor, k_ref, e_ref, depth_src: torch.Tensor, k_src, e_src,
                   condmask_pixel: float, condmask_depth: float):
    dr, ds = _f32c(depth_ref, "depth_ref"), _f32c(depth_src, "depth_src")
    h, w = dr.shape
    if ds.shape != dr.shape:
        raise RuntimeError("depth_ref and depth_src must have the same [H,W] shape")
    dev = dr.device
    mask = torch.empty((h, w), device=dev, dtype=torch.uint8)
    drep = torch.empty((h, w), device=dev, dtype=torch.float32)
    x2d = torch.empty_like(drep)
    y2d = torch.empty_like(drep)
    dp = ctypes.POINTER(ctypes.c_double)
    kr, er, ks, es = _dbl(k_ref, 9), _dbl(e_ref, 16), _dbl(k_src, 9), _dbl(e_src, 16)
    _lib.check(_lib.load().mvster_geo_check_pair(
        _ptr(dr), kr.ctypes.data_as(dp), er.ctypes.data_as(dp), _ptr(ds), ks.ctypes.data_as(dp), es.ctypes.data_as(dp),
        float(condmask_pixel), float(condmask_depth), _ptr(mask), _ptr(drep), _ptr(x2d), _ptr(y2d), h, w, _stream(dr)))
    return mask.bool(), drep, x2d, y2d


def geo_filter(depths: torch.Tensor, confs: torch.Tensor, ks, es, pairs, condmask_pixel: float, condmask_depth: float,
               photomask: float, geomask: int, want_geo_sum: bool = False):
    """K2b over a whole scene.  ``pairs`` [R, 1+S] int (ref, sources...; negative source = skip)."""
    dz, cf = _f32c(depths, "depths"), _f32c(confs, "confs")
    v, h, w = dz.shape
    pr = np.ascontiguousarray(np.asarray(pairs, dtype=np.int32))
    r, s1 = pr.shape
    dev = dz.device
    photo = torch.empty((r, h, w), device=dev, dtype=torch.uint8)
    geo = torch.empty_like(photo)
    final = torch.empty_like(photo)
    avg = torch.empty((r, h, w), device=dev, dtype=torch.float32)
    gsum = torch.empty((r, h, w), device=dev, dtype=torch.int32) if want_geo_sum else None
    kk, ee = _dbl(ks, v * 9), _dbl(es, v * 16)
    dp = ctypes.POINTER(ctypes.c_double)
    _lib.check(_lib.load().mvster_geo_filter(
        _ptr(dz), _ptr(cf), kk.ctypes.data_as(dp), ee.ctypes.data_as(dp),
        pr.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), v, r, s1 - 1, float(condmask_pixel), float(condmask_depth),
        float(photomask), int(geomask), _ptr(photo), _ptr(geo), _ptr(final), _ptr(avg), _ptr(gsum), h, w, _stream(dz)))
    return photo.bool(), geo.bool(), final.bool(), avg, gsum


# ------------------------------------------------------------------------------------------------------------------
def sinkhorn_fwd(gt_depth: torch.Tensor, hypo: torch.Tensor, attn: torch.Tensor, mask: torch.Tensor, iters: int,
                 eps: float, continuous: bool, inverse_depth: bool = False, want_grad: bool = True,
                 want_tmap: bool = False):
    """Fused Sinkhorn loss (K3).  Returns ``(stats [3], grad_px or None, tmap or None)``.

    ``stats`` = (mean loss over masked pixels, masked-pixel count, range_err_ratio); ``grad_px`` is
    d(per-pixel loss)/d(attn) (see :func:`sinkhorn_bwd`).  Reference: ``sinkhorn``, models/mvs4net_utils.py:1164-1210.
    """
    attn = _f32c(attn, "attn_weight")
    hypo = _f32c(hypo, "hypo_depth")
    gt = _f32c(gt_depth, "gt_depth")
    _require_cuda(mask, "mask")
    b, d, h, w = attn.shape
    if tuple(hypo.shape) != (b, d, h, w) or tuple(gt.shape) != (b, h, w) or tuple(mask.shape) != (b, h, w):
        raise RuntimeError("sinkhorn: expected gt/mask [B,H,W] and hypo/attn [B,D,H,W], got %s %s %s %s" % (
            tuple(gt.shape), tuple(mask.shape), tuple(hypo.shape), tuple(attn.shape)))
    m8 = (mask if mask.dtype == torch.bool else mask > 0.5).contiguous().view(torch.uint8)
    lib = _lib.load()
    nblocks = lib.mvster_sinkhorn_blocks(b, d, h, w, int(iters), int(bool(continuous)), int(bool(want_grad)))
    if nblocks <= 0:
        raise RuntimeError("sinkhorn: unsupported (D=%d, iters=%d); D must be 4 or 8" % (d, iters))
    dev = attn.device
    stats = torch.empty(3, device=dev, dtype=torch.float32)
    partials = torch.empty(3 * nblocks, device=dev, dtype=torch.float64)
    grad_px = torch.empty_like(attn) if want_grad else None
    tmap = torch.empty((b, h * w, d, d + (1 if continuous else 0)), device=dev, dtype=torch.float32) if want_tmap else None
    _lib.check(lib.mvster_sinkhorn_fwd(_ptr(gt), _ptr(hypo), _ptr(attn), _ptr(m8), int(iters), float(eps),
                                       int(bool(continuous)), int(bool(inverse_depth)), _ptr(stats), _ptr(grad_px),
                                       _ptr(tmap), _ptr(partials), b, d, h, w, _stream(attn)))
    return stats, grad_px, tmap


def sinkhorn_bwd(grad_px: torch.Tensor, stats: torch.Tensor, grad_loss: torch.Tensor) -> torch.Tensor:
    """``grad_attn = grad_px * grad_loss / count`` (the mean over masked pixels), on the device without a sync."""
    b, d, h, w = grad_px.shape
    gl = _f32c(grad_loss.reshape(1), "grad_loss")
    out = torch.empty_like(grad_px)
    _lib.check(_lib.load().mvster_sinkhorn_bwd(_ptr(grad_px), _ptr(stats), _ptr(gl), _ptr(out), b, d, h, w,
                                               _stream(grad_px)))
    return out
