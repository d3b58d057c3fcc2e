"""CPU oracle for the MVSTER cost-volume hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this file, and only as the checker / the timed CPU baseline (``scripts/bench_extra.py`` additionally times the
``*_port`` functions as the "reference op sequence in eager PyTorch" baseline of its secondary benchmarks - again as the
thing compared against, never as the product).  The product package never imports it and has no CPU fallback.

This is a from-scratch restatement of the reference algorithm (olivier-2018/Deep_reconstruction_with_epipolar_lines_MVSTER);
every function cites the reference lines it follows.  All ``file:line`` citations are relative to the reference tree.

Parity status: PINNED.  The reference ships no tests or golden vectors, so the pin is the reference
itself: ``tests/golden/make_golden.py`` imports the unmodified reference (``models.mvs4net_utils`` and, via ``ast``,
the filter functions of ``test_mvs4.py``) in the build container, runs it on seeded synthetic inputs and freezes
inputs+outputs as ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function below against
those files; ``tests/golden/make_golden_sinkhorn.py`` / ``make_golden_train.py`` do the same for the OT loss
(``sinkhorn``, ``MVS4net_loss``) and for one whole training step, checked by ``tests/test_sinkhorn_oracle.py``.

Two flavours are provided for the fused op:
  * ``*_np``   float64 NumPy, per-pixel "exact" math (Appendix A of SURVEY.md) - the numerical ground truth;
  * ``*_port`` fp32 torch-CPU, op-for-op the reference's eager sequence (materialised warped volume, softmax,
    accumulate) - the CPU baseline that is timed next to the GPU numbers ("port").
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np

try:  # torch is only needed by the *_port functions
    import torch
    import torch.nn.functional as F
except Exception:  # pragma: no cover
    torch = None
    F = None


# ---------------------------------------------------------------------------------------------------------------
# K0: projection composition                                                     models/mvs4net_utils.py:1047-1050,32-34
# ---------------------------------------------------------------------------------------------------------------
def compose_projection_np(proj_view: np.ndarray) -> np.ndarray:
    """``proj_view`` [..., 2, 4, 4] (E, K) -> P [..., 4, 4] = E with rows 0..2 replaced by K[:3,:3] @ E[:3,:4].

    Follows models/mvs4net_utils.py:1047-1050 (the bottom row of E is kept as is).
    """
    e = np.asarray(proj_view[..., 0, :, :], dtype=np.float64)
    k = np.asarray(proj_view[..., 1, :3, :3], dtype=np.float64)
    p = e.copy()
    p[..., :3, :4] = k @ e[..., :3, :4]
    return p


def relative_homography_np(proj: np.ndarray) -> np.ndarray:
    """``proj`` [B, N, 2, 4, 4] -> ``Rt`` [B, N-1, 3, 4] float64 with ``M = P_src @ inv(P_ref)``; Rt = M[:3, :4].

    Follows models/mvs4net_utils.py:32-34 (rot = M[:3,:3], trans = M[:3,3]).
    """
    p = compose_projection_np(proj)                       # [B,N,4,4]
    pref_inv = np.linalg.inv(p[:, 0])                     # [B,4,4]
    m = p[:, 1:] @ pref_inv[:, None]                      # [B,N-1,4,4]
    return m[:, :, :3, :4].copy()


# ---------------------------------------------------------------------------------------------------------------
# K1: homography warp + group correlation + epipolar attention + view aggregation
#                                                                             models/mvs4net_utils.py:21-67,1027-1102
# ---------------------------------------------------------------------------------------------------------------
def sample_coords_np(rt: np.ndarray, hypo: np.ndarray, origin: Tuple[int, int] = (0, 0)
                     ) -> Tuple[np.ndarray, np.ndarray]:
    """Source-image sample coordinates for every (pixel, hypothesis).

    ``rt`` [3,4] float64, ``hypo`` [D,H,W] -> (sx, sy) each [D,H,W] float64.  ``origin`` = (y0, x0) of ``hypo``'s
    top-left pixel in the full reference image (for checking a window of a large image).
    models/mvs4net_utils.py:36-48: p = R @ [x, y, 1]^T * d + t ; z == 0 -> 1e-9 ; (sx, sy) = p.xy / z.
    The normalise (:51-52) / un-normalise (grid_sample, align_corners=True) pair is the identity in exact math.
    """
    d, h, w = hypo.shape
    ys, xs = np.meshgrid(np.arange(h, dtype=np.float64) + origin[0], np.arange(w, dtype=np.float64) + origin[1],
                         indexing="ij")
    r = rt[:, :3]
    t = rt[:, 3]
    ax = r[0, 0] * xs + r[0, 1] * ys + r[0, 2]
    ay = r[1, 0] * xs + r[1, 1] * ys + r[1, 2]
    az = r[2, 0] * xs + r[2, 1] * ys + r[2, 2]
    hy = hypo.astype(np.float64)
    px = ax[None] * hy + t[0]
    py = ay[None] * hy + t[1]
    pz = az[None] * hy + t[2]
    pz = np.where(pz == 0.0, 1e-9, pz)
    return px / pz, py / pz


def bilinear_zeros_np(src: np.ndarray, sx: np.ndarray, sy: np.ndarray) -> np.ndarray:
    """``F.grid_sample(mode='bilinear', padding_mode='zeros', align_corners=True)`` in pixel coordinates.

    ``src`` [C,Hs,Ws]; ``sx``/``sy`` [...] -> [C, ...].  Taps at floor(), out-of-image taps contribute 0
    (models/mvs4net_utils.py:59; semantics of torch's GridSampler bilinear kernel).
    """
    c, hs, ws = src.shape
    src = src.astype(np.float64)
    finite = np.isfinite(sx) & np.isfinite(sy)
    sx = np.where(finite, sx, -10.0)
    sy = np.where(finite, sy, -10.0)
    # clamp far-away coordinates so the integer conversion below cannot overflow (they sample nothing either way)
    sx = np.clip(sx, -4.0, ws + 4.0)
    sy = np.clip(sy, -4.0, hs + 4.0)
    x0 = np.floor(sx)
    y0 = np.floor(sy)
    fx = sx - x0
    fy = sy - y0
    out = np.zeros((c,) + sx.shape, dtype=np.float64)
    for dy, wy in ((0, 1.0 - fy), (1, fy)):
        for dx, wx in ((0, 1.0 - fx), (1, fx)):
            xi = (x0 + dx).astype(np.int64)
            yi = (y0 + dy).astype(np.int64)
            valid = (xi >= 0) & (xi < ws) & (yi >= 0) & (yi < hs)
            xi = np.where(valid, xi, 0)
            yi = np.where(valid, yi, 0)
            out += src[:, yi, xi] * (wx * wy * valid)[None]
    return out


def homo_warping_np(src_fea: np.ndarray, src_proj: np.ndarray, ref_proj: np.ndarray,
                    depth_values: np.ndarray) -> np.ndarray:
    """models/mvs4net_utils.py:21-67 for 4-D ``src_fea`` [B,C,Hs,Ws] and ``depth_values`` [B,D,H,W] -> [B,C,D,H,W]."""
    b = src_fea.shape[0]
    outs = []
    for i in range(b):
        m = np.asarray(src_proj[i], np.float64) @ np.linalg.inv(np.asarray(ref_proj[i], np.float64))
        sx, sy = sample_coords_np(m[:3, :4], depth_values[i])
        outs.append(bilinear_zeros_np(src_fea[i], sx, sy))
    return np.stack(outs)


def epipolar_aggregate_np(ref: np.ndarray, srcs: Sequence[np.ndarray], proj: np.ndarray, hypo: np.ndarray,
                          groups: int, attn_temp: float, group_cor: bool = True, attn_fuse_d: bool = True,
                          rt: Optional[np.ndarray] = None, window: Optional[Tuple[int, int, int, int]] = None):
    """Float64 restatement of ``stagenet.forward`` steps 1-2 (models/mvs4net_utils.py:1030-1102).

    ``ref`` [B,C,H,W]; ``srcs`` list of N-1 arrays [B,C,Hs,Ws]; ``proj`` [B,N,2,4,4]; ``hypo`` [B,D,H,W].
    Returns ``(volume [B,G,D,H,W], weights [B,N-1,D,H,W] (or [B,N-1,H,W] when not attn_fuse_d), wsum)``.
    ``rt`` overrides the homographies (e.g. the fp32-rounded ones the CUDA path uses).
    ``window`` = (y0, y1, x0, x1) restricts the computation to that block of reference pixels (the sources stay
    whole), so that windows of full-size images can be checked in seconds.
    """
    origin = (0, 0)
    if window is not None:
        y0, y1, x0, x1 = window
        ref = ref[:, :, y0:y1, x0:x1]
        hypo = hypo[:, :, y0:y1, x0:x1]
        origin = (y0, x0)
    b, c, h, w = ref.shape
    d = hypo.shape[1]
    if rt is None:
        rt = relative_homography_np(proj)
    g = groups if group_cor else c
    ref64 = ref.astype(np.float64)
    vol = np.zeros((b, g, d, h, w))
    if attn_fuse_d:
        wsum = np.full((b, d, h, w), 1e-8)                                                   # :1037
        weights = np.zeros((b, len(srcs), d, h, w))
    else:
        wsum = np.full((b, h, w), 1e-8)
        weights = np.zeros((b, len(srcs), h, w))
    for i in range(b):
        for v, src in enumerate(srcs):
            sx, sy = sample_coords_np(np.asarray(rt[i, v], np.float64).reshape(3, 4), hypo[i], origin)
            warped = bilinear_zeros_np(src[i], sx, sy)                                       # [C,D,H,W]
            if group_cor:                                                                    # :1066-1069
                cor = (warped.reshape(g, c // g, d, h, w) * ref64[i].reshape(g, c // g, 1, h, w)).mean(1)
            else:                                                                            # :1071
                cor = (ref64[i][:, None] - warped) ** 2
            score = cor.sum(0)                                                               # [D,H,W]
            if attn_fuse_d:                                                                  # :1083
                s = score / attn_temp
                s = s - s.max(0, keepdims=True)
                e = np.exp(s)
                wgt = e / e.sum(0, keepdims=True) / math.sqrt(c)
                wsum[i] += wgt
                vol[i] += wgt[None] * cor
            else:                                                                            # :1079-1081
                s = score - score.max(0, keepdims=True)
                e = np.exp(s)
                wgt = (e / e.sum(0, keepdims=True)).max(0)
                wsum[i] += wgt
                vol[i] += wgt[None, None] * cor
            weights[i, v] = wgt
    if attn_fuse_d:
        vol = vol / wsum[:, None]                                                            # :1100
    else:
        vol = vol / wsum[:, None, None]                                                      # :1098
    return vol, weights, wsum


def compose_projection_port(proj_view):
    """torch fp32 version of models/mvs4net_utils.py:1047-1050 (one view, [B,2,4,4] -> [B,4,4])."""
    p = proj_view[:, 0].clone()
    p[:, :3, :4] = torch.matmul(proj_view[:, 1, :3, :3], proj_view[:, 0, :3, :4])
    return p


def homo_warping_port(src_fea, src_proj, ref_proj, depth_values):
    """fp32 torch port of models/mvs4net_utils.py:21-67 with the reference's own op sequence.

    This materialises the [B,C,D,H,W] warped volume exactly as the reference does; it is what the CPU baseline times.
    """
    b, c, hs, ws = src_fea.shape
    _, nd, h, w = depth_values.shape
    with torch.no_grad():
        m = torch.matmul(src_proj, torch.inverse(ref_proj))                                  # :32
        rot, trans = m[:, :3, :3], m[:, :3, 3:4]
        yy, xx = torch.meshgrid(torch.arange(h, dtype=torch.float32, device=src_fea.device),
                                torch.arange(w, dtype=torch.float32, device=src_fea.device), indexing="ij")
        pix = torch.stack((xx.reshape(-1), yy.reshape(-1), torch.ones(h * w, device=src_fea.device)))  # [3,HW]
        rays = torch.matmul(rot, pix.unsqueeze(0).expand(b, -1, -1))                         # :42
        pts = rays.unsqueeze(2) * depth_values.reshape(b, 1, nd, h * w) + trans.reshape(b, 3, 1, 1)  # :43-44
        z = pts[:, 2:3]
        z = torch.where(z == 0, torch.full_like(z, 1e-9), z)                                 # :46-47
        xy = pts[:, :2] / z                                                                  # :48
        gx = xy[:, 0] / ((ws - 1) / 2) - 1                                                   # :51
        gy = xy[:, 1] / ((hs - 1) / 2) - 1                                                   # :52
        grid = torch.stack((gx, gy), dim=3)                                                  # [B,D,HW,2]
    out = F.grid_sample(src_fea, grid.reshape(b, nd * h, w, 2), mode="bilinear", padding_mode="zeros",
                        align_corners=True)                                                  # :59
    return out.reshape(b, c, nd, h, w)


def epipolar_aggregate_port(features, proj_matrices, depth_hypo, groups, attn_temp, return_weights=False):
    """fp32 torch port of ``stagenet.forward`` steps 1-2 (models/mvs4net_utils.py:1030-1102), group_cor + attn_fuse_d.

    ``features`` list of N tensors [B,C,H,W]; ``proj_matrices`` [B,N,2,4,4]; returns volume [B,G,D,H,W]
    (and the per-view weights [N-1][B,D,H,W]).  Differentiable w.r.t. ``features`` like the reference.
    """
    views = torch.unbind(proj_matrices, 1)
    ref, srcs = features[0], features[1:]
    b, d, h, w = depth_hypo.shape
    c = ref.shape[1]
    ref_vol = ref.unsqueeze(2).repeat(1, 1, d, 1, 1).reshape(b, groups, c // groups, d, h, w)  # :1036,1068
    wsum = 1e-8
    acc = 0
    weights = []
    for src, pv in zip(srcs, views[1:]):
        warped = homo_warping_port(src, compose_projection_port(pv), compose_projection_port(views[0]), depth_hypo)
        cor = (warped.reshape(b, groups, c // groups, d, h, w) * ref_vol).mean(2)            # :1067-1069
        wgt = torch.softmax(cor.sum(1) / attn_temp, 1) / math.sqrt(c)                        # :1083
        wsum = wsum + wgt
        acc = acc + wgt.unsqueeze(1) * cor                                                   # :1085
        weights.append(wgt)
    vol = acc / wsum.unsqueeze(1)                                                            # :1100
    if return_weights:
        return vol, weights
    return vol


# ---------------------------------------------------------------------------------------------------------------
# hypothesis schedule                                                              models/mvs4net_utils.py:79-94
# ---------------------------------------------------------------------------------------------------------------
def init_inverse_range_np(cur_depth: np.ndarray, ndepths: int, h: int, w: int, dtype=np.float32) -> np.ndarray:
    """models/mvs4net_utils.py:79-85: D hypotheses uniform in 1/d from 1/d[:, -1] (index 0 = far) to 1/d[:, 0]."""
    cur = cur_depth.astype(dtype)
    inv_min = (dtype(1.0) / cur[:, 0])
    inv_max = (dtype(1.0) / cur[:, -1])
    itv = (np.arange(ndepths, dtype=dtype) / dtype(ndepths - 1)).reshape(1, -1, 1, 1)
    inv = inv_max[:, None, None, None] + (inv_min - inv_max)[:, None, None, None] * itv
    inv = np.broadcast_to(inv, (cur.shape[0], ndepths, h, w))
    return (dtype(1.0) / inv).astype(dtype)


def upsample_bilinear_align_corners_np(x: np.ndarray, h: int, w: int) -> np.ndarray:
    """``F.interpolate(x[:,None], [D,H,W], mode='trilinear', align_corners=True)`` with D unchanged.

    With the depth size unchanged the depth lerp weight is exactly (1, 0), so this is a per-plane bilinear
    upsample: source index = dst * (in-1)/(out-1), lower tap = int(), upper tap = lower + (lower < in-1).
    """
    dt = x.dtype.type
    hin, win = x.shape[-2:]
    sh = dt(hin - 1) / dt(h - 1) if h > 1 else dt(0)
    sw = dt(win - 1) / dt(w - 1) if w > 1 else dt(0)
    fy = (np.arange(h, dtype=x.dtype) * sh).astype(x.dtype)
    fx = (np.arange(w, dtype=x.dtype) * sw).astype(x.dtype)
    y0 = fy.astype(np.int64)
    x0 = fx.astype(np.int64)
    y1 = y0 + (y0 < hin - 1)
    x1 = x0 + (x0 < win - 1)
    ly = (fy - y0.astype(x.dtype)).astype(x.dtype)
    lx = (fx - x0.astype(x.dtype)).astype(x.dtype)
    top = x[..., y0, :][..., :, x0] * (dt(1) - lx) + x[..., y0, :][..., :, x1] * lx
    bot = x[..., y1, :][..., :, x0] * (dt(1) - lx) + x[..., y1, :][..., :, x1] * lx
    return (top * (dt(1) - ly)[:, None] + bot * ly[:, None]).astype(x.dtype)


def schedule_inverse_range_np(inverse_min_depth: np.ndarray, inverse_max_depth: np.ndarray, ndepths: int,
                              h: int, w: int) -> np.ndarray:
    """models/mvs4net_utils.py:87-94: per-pixel inverse-depth lerp at (H/2, W/2), bilinear x2 upsample, reciprocal."""
    dt = inverse_min_depth.dtype.type
    itv = (np.arange(ndepths, dtype=inverse_min_depth.dtype) / dt(ndepths - 1)).reshape(1, -1, 1, 1)
    inv = inverse_max_depth[:, None] + (inverse_min_depth - inverse_max_depth)[:, None] * itv
    inv = upsample_bilinear_align_corners_np(inv.astype(inverse_min_depth.dtype), h, w)
    return (dt(1.0) / inv).astype(inverse_min_depth.dtype)


# ---------------------------------------------------------------------------------------------------------------
# K2a: depth / confidence tail                                                 models/mvs4net_utils.py:1105-1162
# ---------------------------------------------------------------------------------------------------------------
def tail_np(logits: np.ndarray, hypo: np.ndarray, split_itv: float, training: bool = False,
            regress: bool = False):
    """``logits`` [B,D,H,W] (regnet output), ``hypo`` [B,D,H,W] -> dict like ``stagenet.forward``'s ``ret_dict``.

    conf = max_d L / sum_d L on the raw logits, eval only (:1109-1113,1138; 0.0 scalar in training, :1143-1144);
    attn = softmax_d L (:1126); depth = hypo[argmax_d attn] (first maximum, :1129-1130), or sum_d attn*hypo when
    ``regress`` (models/module.py:935-941 / the commented line :1133);
    itv = 1/hypo[:,2] - 1/hypo[:,1]; inverse_min/max_depth = 1/depth +- split_itv*itv (:1151-1156).
    """
    dt = logits.dtype.type
    m = logits.max(1, keepdims=True)
    e = np.exp(logits - m)
    attn = (e / e.sum(1, keepdims=True)).astype(logits.dtype)
    idx = attn.argmax(1)
    if regress:
        depth = (attn * hypo).sum(1).astype(logits.dtype)
    else:
        depth = np.take_along_axis(hypo, idx[:, None], 1)[:, 0]
    if training:
        conf = np.zeros((), dtype=np.float32)
    else:
        lidx = logits.argmax(1)
        conf = np.take_along_axis(logits, lidx[:, None], 1)[:, 0] / logits.sum(1)
    itv = dt(1.0) / hypo[:, 2] - dt(1.0) / hypo[:, 1]
    inv_min = dt(1.0) / depth + dt(split_itv) * itv
    inv_max = dt(1.0) / depth - dt(split_itv) * itv
    return {"depth": depth, "photometric_confidence": conf, "hypo_depth": hypo, "attn_weight": attn,
            "inverse_min_depth": inv_min, "inverse_max_depth": inv_max, "argmax": idx}


# ---------------------------------------------------------------------------------------------------------------
# K2a': last layers of reg2d (eval) + tail                                  models/mvs4net_utils.py:899-926,1109-1156
# ---------------------------------------------------------------------------------------------------------------
def reg2d_last_layers_np(low: np.ndarray, skip: np.ndarray, deconv_w: np.ndarray, bn_weight: np.ndarray,
                         bn_bias: np.ndarray, bn_mean: np.ndarray, bn_var: np.ndarray, prob_w: np.ndarray,
                         prob_b: float, eps: float = 1e-5) -> np.ndarray:
    """float64 restatement of ``x = conv0 + relu(bn(conv11(x))); logits = prob(x)`` (:923-926) in eval mode.

    ``low`` [B,Ci,D,h,w]; ``skip`` [B,Co,D,2h,2w]; ``deconv_w`` [Ci,Co,1,3,3] (ConvTranspose3d weight layout,
    stride (1,2,2), padding (0,1,1), output_padding (0,1,1): out[y] gathers in[iy] with y = 2*iy - 1 + ky).
    Returns the logits [B,D,2h,2w] as float64."""
    low = low.astype(np.float64)
    b, ci, d, h, w = low.shape
    co = deconv_w.shape[1]
    wt = deconv_w.astype(np.float64)[:, :, 0]
    out = np.zeros((b, co, d, 2 * h, 2 * w), dtype=np.float64)
    for ky in range(3):
        for kx in range(3):
            contrib = np.einsum("bcdyx,ck->bkdyx", low, wt[:, :, ky, kx])
            ys = 2 * np.arange(h) - 1 + ky
            xs = 2 * np.arange(w) - 1 + kx
            vy, vx = (ys >= 0) & (ys < 2 * h), (xs >= 0) & (xs < 2 * w)
            out[:, :, :, ys[vy][:, None], xs[vx][None, :]] += contrib[:, :, :, vy][:, :, :, :, vx]
    scale = bn_weight.astype(np.float64) / np.sqrt(bn_var.astype(np.float64) + eps)
    shift = bn_bias.astype(np.float64) - bn_mean.astype(np.float64) * scale
    out = out * scale[None, :, None, None, None] + shift[None, :, None, None, None]
    x = skip.astype(np.float64) + np.maximum(out, 0.0)
    return np.einsum("bcdyx,c->bdyx", x, prob_w.astype(np.float64).reshape(-1)) + float(prob_b)


# ---------------------------------------------------------------------------------------------------------------
# K2b: geometric consistency filter + mask fusion                                     test_mvs4.py:612-670,716-749
# ---------------------------------------------------------------------------------------------------------------
def remap_linear_np(img: np.ndarray, mapx: np.ndarray, mapy: np.ndarray) -> np.ndarray:
    """Emulation of ``cv2.remap(img, mapx, mapy, INTER_LINEAR)`` with BORDER_CONSTANT 0 (test_mvs4.py:632).

    OpenCV converts the float32 maps to fixed point with INTER_BITS = 5: ix = cvRound(x*32) (round half to even),
    integer tap = ix >> 5 (floor), fractional weight = (ix & 31)/32; taps outside the image read 0.
    The weights are applied in float32 as products of the two 1-D weights (OpenCV's bilinear table).
    """
    h, w = img.shape
    x = np.asarray(mapx, np.float32)
    y = np.asarray(mapy, np.float32)
    bad = ~(np.isfinite(x) & np.isfinite(y))
    xs = np.where(bad, np.float32(-1e6), x).astype(np.float64) * 32.0
    ys = np.where(bad, np.float32(-1e6), y).astype(np.float64) * 32.0
    # cvRound saturates through int conversion; far-away values sample nothing, clamp keeps the ints well-defined
    xs = np.clip(xs, -1e8, 1e8)
    ys = np.clip(ys, -1e8, 1e8)
    ix = np.rint(xs).astype(np.int64)
    iy = np.rint(ys).astype(np.int64)
    x0 = ix >> 5
    y0 = iy >> 5
    fx = ((ix & 31).astype(np.float32)) / np.float32(32.0)
    fy = ((iy & 31).astype(np.float32)) / np.float32(32.0)
    out = np.zeros(x.shape, dtype=np.float32)
    one = np.float32(1.0)
    for dy, wy in ((0, one - fy), (1, fy)):
        for dx, wx in ((0, one - fx), (1, fx)):
            xi = x0 + dx
            yi = y0 + dy
            valid = (xi >= 0) & (xi < w) & (yi >= 0) & (yi < h)
            v = img[np.where(valid, yi, 0), np.where(valid, xi, 0)].astype(np.float32)
            out += np.where(valid, v, np.float32(0)) * (wx * wy).astype(np.float32)
    return out


def reproject_with_depth_np(depth_ref, k_ref, e_ref, depth_src, k_src, e_src, use_cv2: bool = False):
    """test_mvs4.py:612-649 restated per pixel in float64; returns the same 5 float32 maps.

    ``use_cv2=True`` samples with the real ``cv2.remap`` (what the CPU baseline times); otherwise the 1/32-px
    fixed-point emulation above is used (what the CUDA kernel implements).
    """
    h, w = depth_ref.shape
    k_ref = np.asarray(k_ref, np.float64)
    k_src = np.asarray(k_src, np.float64)
    e_ref = np.asarray(e_ref, np.float64)
    e_src = np.asarray(e_src, np.float64)
    xs, ys = np.meshgrid(np.arange(w), np.arange(h))
    pix = np.vstack((xs.reshape(-1), ys.reshape(-1), np.ones(h * w, dtype=xs.dtype)))
    with np.errstate(divide="ignore", invalid="ignore"):
        xyz_ref = np.linalg.inv(k_ref) @ (pix * depth_ref.reshape(-1))                         # :619-620
        xyz_src = ((e_src @ np.linalg.inv(e_ref)) @ np.vstack((xyz_ref, np.ones(h * w))))[:3]   # :622-623
        kx = k_src @ xyz_src                                                                    # :625
        xy_src = kx[:2] / kx[2:3]                                                               # :626
        x_src = xy_src[0].reshape(h, w).astype(np.float32)                                      # :630-631
        y_src = xy_src[1].reshape(h, w).astype(np.float32)
        if use_cv2:
            import cv2
            sampled = cv2.remap(depth_src, x_src, y_src, interpolation=cv2.INTER_LINEAR)        # :632
        else:
            sampled = remap_linear_np(depth_src, x_src, y_src)
        xyz_s = np.linalg.inv(k_src) @ (np.vstack((xy_src, np.ones(h * w))) * sampled.reshape(-1))      # :637-638
        xyz_r = ((e_ref @ np.linalg.inv(e_src)) @ np.vstack((xyz_s, np.ones(h * w))))[:3]       # :640-641
        depth_rep = xyz_r[2].reshape(h, w).astype(np.float32)                                   # :643
        kr = k_ref @ xyz_r                                                                      # :644
        xy_rep = kr[:2] / kr[2:3]                                                               # :645
    x_rep = xy_rep[0].reshape(h, w).astype(np.float32)
    y_rep = xy_rep[1].reshape(h, w).astype(np.float32)
    return depth_rep, x_rep, y_rep, x_src, y_src


def check_geometric_consistency_np(depth_ref, k_ref, e_ref, depth_src, k_src, e_src, condmask_pixel: float,
                                   condmask_depth: float, use_cv2: bool = False):
    """test_mvs4.py:653-670: returns (mask bool, depth_reprojected (0 where ~mask), x2d_src, y2d_src)."""
    h, w = depth_ref.shape
    xs, ys = np.meshgrid(np.arange(w), np.arange(h))
    d_rep, x_rep, y_rep, x_src, y_src = reproject_with_depth_np(depth_ref, k_ref, e_ref, depth_src, k_src, e_src,
                                                                use_cv2)
    with np.errstate(divide="ignore", invalid="ignore"):
        dist = np.sqrt((x_rep - xs) ** 2 + (y_rep - ys) ** 2)                                  # :661 (float64)
        rel = np.abs(d_rep - depth_ref) / depth_ref                                            # :664-665 (float32)
        mask = np.logical_and(dist < condmask_pixel, rel < condmask_depth)                    # :667
    d_rep = d_rep.copy()
    d_rep[~mask] = 0                                                                           # :668
    return mask, d_rep, x_src, y_src


def depth2pts_np(depth_map: np.ndarray, cam_intrinsic: np.ndarray, cam_extrinsic: np.ndarray) -> np.ndarray:
    """World points [H*W,3] float64 of a depth map: pixel-centre grid (x+0.5, y+0.5, 1), ``K^-1``, times depth, then
    ``R^-1 (X - t)`` (test_mvs4.py:206-229)."""
    h, w = depth_map.shape
    xs, ys = np.meshgrid(np.linspace(0.5, w - 0.5, w), np.linspace(0.5, h - 0.5, h))
    grid = np.stack([xs.reshape(-1), ys.reshape(-1), np.ones(h * w)], 0)
    cam = (np.linalg.inv(cam_intrinsic) @ grid) * depth_map.reshape(1, -1)
    r, t = cam_extrinsic[:3, :3], cam_extrinsic[:3, 3:4]
    return (np.linalg.inv(r) @ (cam - t)).T


def filter_fuse_np(depths: np.ndarray, confs: np.ndarray, ks: np.ndarray, es: np.ndarray, pairs: np.ndarray,
                   condmask_pixel: float, condmask_depth: float, photomask: float, geomask: int,
                   use_cv2: bool = False):
    """Mask fusion of ``filter_depth`` (test_mvs4.py:716,725-749) for every reference view of ``pairs`` [R, 1+S].

    ``pairs[r] = (ref, src_1 .. src_S)``; ``depths``/``confs`` [V,H,W] float32; ``ks`` [V,3,3], ``es`` [V,4,4] float64.
    Returns (photo [R,H,W] bool, geo, final, depth_avg [R,H,W] float32, geo_sum int32).
    """
    r = pairs.shape[0]
    h, w = depths.shape[1:]
    photo = np.zeros((r, h, w), bool)
    geo = np.zeros((r, h, w), bool)
    final = np.zeros((r, h, w), bool)
    avg = np.zeros((r, h, w), np.float32)
    gsum = np.zeros((r, h, w), np.int32)
    for i in range(r):
        ref = int(pairs[i, 0])
        photo[i] = confs[ref] > photomask                                                     # :716
        acc = 0
        cnt = 0
        for s in pairs[i, 1:]:
            s = int(s)
            m, d_rep, _, _ = check_geometric_consistency_np(depths[ref], ks[ref], es[ref], depths[s], ks[s], es[s],
                                                            condmask_pixel, condmask_depth, use_cv2)
            cnt = cnt + m.astype(np.int32)                                                     # :738
            acc = acc + d_rep                                                                  # sum() at :744
        avg[i] = (acc + depths[ref]) / (cnt + 1)                                               # :744
        gsum[i] = cnt
        geo[i] = cnt >= geomask                                                                # :746
        final[i] = np.logical_and(photo[i], geo[i])                                            # :749
    return photo, geo, final, avg, gsum


# ---------------------------------------------------------------------------------------------------------------
# whole hot path of one depth map on the CPU (what bench.py times as the CPU baseline / reference arm)
# ---------------------------------------------------------------------------------------------------------------
def cascade_port(features, projs, depth_values, logits, groups, ndepths, split_itv, attn_temp=2.0):
    """The hot-path slice of ``MVS4net.forward`` (models/MVS4Net.py:99-136) on the CPU with the reference's own
    op sequence: per stage the hypothesis schedule (:109-118), ``stagenet`` steps 1-2 (fp32 torch port above) and the
    tail (:1109-1156, torch ops as in the reference).  ``regnet`` is out of scope: ``logits[s]`` stands in for its
    output, exactly as in the GPU benchmark.

    ``features[s]`` = list of N tensors [B,C,H,W]; ``projs[s]`` [B,N,2,4,4]; ``depth_values`` [B,2].
    Returns the last stage's (depth, confidence).
    """
    out = None
    for s in range(len(features)):
        b, c, h, w = features[s][0].shape
        d = ndepths[s]
        if s == 0:                                                                          # MVS4Net.py:109-111
            inv_min = 1.0 / depth_values[:, 0]
            inv_max = 1.0 / depth_values[:, -1]
            itv = torch.arange(0, d, dtype=torch.float32).reshape(1, -1, 1, 1).repeat(1, 1, h, w) / (d - 1)
            hypo = 1.0 / (inv_max[:, None, None, None] + (inv_min - inv_max)[:, None, None, None] * itv)
        else:                                                                               # MVS4Net.py:115-116
            itv = torch.arange(0, d, dtype=torch.float32).reshape(1, -1, 1, 1).repeat(1, 1, h // 2, w // 2) / (d - 1)
            inv = out["inverse_max_depth"][:, None] + (out["inverse_min_depth"] - out["inverse_max_depth"])[:, None] * itv
            inv = F.interpolate(inv.unsqueeze(1), [d, h, w], mode="trilinear", align_corners=True).squeeze(1)
            hypo = 1.0 / inv
        vol = epipolar_aggregate_port(features[s], projs[s], hypo, groups[s], attn_temp)
        lg = logits[s](hypo) if callable(logits[s]) else logits[s]   # stand-in for regnet(vol)
        idx = lg.max(1, keepdim=True)[1]                                                    # :1109-1113
        conf = torch.gather(lg, 1, idx).squeeze(1) / lg.sum(1)
        attn = F.softmax(lg, dim=1)                                                         # :1126
        depth = torch.gather(hypo, 1, attn.max(1, keepdim=True)[1]).squeeze(1)              # :1129-1130
        last_itv = 1.0 / hypo[:, 2] - 1.0 / hypo[:, 1]                                      # :1152
        out = {"depth": depth, "photometric_confidence": conf, "attn_weight": attn, "volume": vol,
               "inverse_min_depth": 1 / depth + split_itv[s] * last_itv,
               "inverse_max_depth": 1 / depth - split_itv[s] * last_itv}
    return out["depth"], out["photometric_confidence"]


# ---------------------------------------------------------------------------------------------------------------
# K3: Sinkhorn / Wasserstein depth loss                     models/mvs4net_utils.py:1164-1210, models/MVS4Net.py:225-234
# ---------------------------------------------------------------------------------------------------------------
def _logsumexp_np(x: np.ndarray, axis: int) -> np.ndarray:
    m = np.max(x, axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0.0)
    return (np.log(np.sum(np.exp(x - m), axis=axis, keepdims=True)) + m).squeeze(axis)


def sinkhorn_np(gt_depth: np.ndarray, hypo_depth: np.ndarray, attn_weight: np.ndarray, mask: np.ndarray, iters: int,
                eps: float = 1.0, continuous: bool = False, inverse_depth: bool = False, dtype=np.float64):
    """float64 restatement of ``sinkhorn`` (models/mvs4net_utils.py:1164-1210) with the analytic gradient.

    Inputs keep the reference's layouts: gt/mask [B,H,W], hypo/attn [B,D,H,W].  Returns a dict with
      ``T_map`` [B,HW,D,NC], ``loss`` (mean over masked pixels, :1208), ``grad_attn`` [B,D,H,W] = d loss / d attn_weight
      (what autograd yields for the reference), ``count`` and ``range_err_ratio`` (models/MVS4Net.py:225-232).
    Quirk kept: the cost enters the exponent with a positive sign (``D_map/eps``, :1200-1204).
    """
    gt = np.asarray(gt_depth, dtype=np.float32)
    hypo32 = np.asarray(hypo_depth, dtype=np.float32)
    b, d, h, w = hypo32.shape
    msk = np.asarray(mask).astype(bool)
    ar = np.arange(d)
    if not continuous:
        cost = np.abs(ar[:, None] - ar[None, :]).astype(dtype)                              # :1173
        cost = np.broadcast_to(cost, (b, h * w, d, d)).copy()                               # :1174
        gi = np.abs(hypo32 - gt[:, None]).argmin(1).reshape(b, h * w)                       # :1175 (fp32 like the ref)
        gt_dist = np.zeros((b, h * w, d), dtype=dtype)                                      # :1176-1178
        np.put_along_axis(gt_dist, gi[..., None], 1.0, axis=2)
    else:
        gt_dist = np.zeros((b, h * w, d + 1), dtype=dtype)                                  # :1180-1181
        gt_dist[:, :, -1] = 1
        cost = np.zeros((b, h, w, d, d + 1), dtype=dtype)                                   # :1182-1184
        cost[..., :d, :d] = np.abs(ar[:, None] - ar[None, :])
        with np.errstate(divide="ignore", invalid="ignore"):
            itv = np.float32(1) / hypo32[:, 2] - np.float32(1) / hypo32[:, 1]               # :1185 (fp32 like the ref)
            gbd = (np.float32(1) / gt - np.float32(1) / hypo32[:, 0]) / itv                 # :1186
        gbd[~msk] = 10                                                                      # :1188
        cost[..., -1] = np.abs(gbd[..., None].astype(dtype) - ar)                           # :1190-1191
        cost = cost.reshape(b, h * w, d, d + 1)                                             # :1192
    pred = np.transpose(np.asarray(attn_weight, dtype=np.float32), (0, 2, 3, 1)).reshape(b, h * w, d)   # :1194
    # the +1e-12 is an fp32 addition in the reference; keep it so that log(1 + 1e-12) == 0 exactly
    log_mu = np.log((gt_dist.astype(np.float32) + np.float32(1e-12)).astype(dtype))         # :1197
    at = (pred + np.float32(1e-12)).astype(dtype)
    log_nu = np.log(at)                                                                     # :1198
    kmat = cost / dtype(eps)
    u = np.zeros_like(log_nu)
    v = np.zeros_like(log_mu)
    lus, lvs = [], []
    for _ in range(iters):                                                                  # :1201-1204
        lv = _logsumexp_np(kmat + u[..., :, None], axis=2)
        v = log_mu - lv
        lu = _logsumexp_np(kmat + v[..., None, :], axis=3)
        u = log_nu - lu
        lus.append(lu)
        lvs.append(lv)
    tmap = np.exp(kmat + u[..., :, None] + v[..., None, :])                                 # :1207
    per_px = (tmap * cost).reshape(b * h * w, -1).sum(-1)
    sel = msk.reshape(-1)
    count = int(sel.sum())
    loss = per_px[sel].mean() if count else float("nan")                                    # :1208
    # reverse sweep of the unrolled iterations
    g = tmap * cost
    gu, gv = g.sum(3), g.sum(2)
    gl = np.zeros_like(log_nu)
    for t in range(iters - 1, -1, -1):
        gl += gu
        vt = log_mu - lvs[t]
        gv = gv - np.einsum("bpi,bpij->bpj", gu, np.exp(kmat + vt[..., None, :] - lus[t][..., :, None]))
        up = (log_nu - lus[t - 1]) if t > 0 else np.zeros_like(log_nu)
        gu = -np.einsum("bpj,bpij->bpi", gv, np.exp(kmat + up[..., :, None] - lvs[t][..., None, :]))
        gv = np.zeros_like(gv)
    grad = gl / at
    grad = grad * sel.reshape(b, h * w, 1) / max(count, 1)
    grad = np.transpose(grad.reshape(b, h, w, d), (0, 3, 1, 2))
    # out-of-range statistic                                                              models/MVS4Net.py:225-232
    with np.errstate(divide="ignore", invalid="ignore"):
        if inverse_depth:
            ditv = np.abs(np.float32(1) / hypo32[:, 2] - np.float32(1) / hypo32[:, 1])
            oor = (np.abs(np.float32(1) / hypo32 - (np.float32(1) / gt)[:, None]) <= ditv[:, None]).sum(1) == 0
        else:
            ditv = np.abs(hypo32[:, 2] - hypo32[:, 1])
            oor = (np.abs(hypo32 - gt[:, None]) <= ditv[:, None]).sum(1) == 0
    ratio = float(oor[msk].astype(np.float32).mean()) if count else float("nan")
    return {"T_map": tmap, "loss": float(loss), "grad_attn": grad, "count": count, "range_err_ratio": ratio}


def sinkhorn_port(gt_depth, hypo_depth, attn_weight, mask, iters, eps=1.0, continuous=False):
    """fp32 torch-CPU port with the reference's own op sequence (materialised [B,HW,D,D(+1)] tensors, 2*iters
    ``torch.logsumexp`` passes, autograd for the gradient) - the timed CPU baseline of K3.  Returns (T_map, loss)."""
    b, d, h, w = attn_weight.shape
    dev = attn_weight.device
    ar = torch.arange(d, dtype=torch.float32, device=dev)
    base = (ar[:, None] - ar[None, :]).abs()
    if not continuous:
        cost = base[None, None].repeat(b, h * w, 1, 1)
        gi = (hypo_depth - gt_depth[:, None]).abs().min(1)[1].reshape(b * h * w, 1)
        gt_dist = torch.zeros(b * h * w, d, device=dev).scatter_add_(1, gi, torch.ones(b * h * w, 1, device=dev)).reshape(b, h * w, d)
    else:
        gt_dist = torch.zeros(b, h * w, d + 1, device=dev)
        gt_dist[:, :, -1] = 1
        cost = torch.zeros(b, d, d + 1, device=dev)
        cost[:, :d, :d] = base
        cost = cost[:, None, None].repeat(1, h, w, 1, 1)
        itv = 1 / hypo_depth[:, 2] - 1 / hypo_depth[:, 1]
        gbd = (1 / gt_depth - 1 / hypo_depth[:, 0]) / itv
        gbd[~mask] = 10
        cost[..., -1] = torch.stack([(gbd - i).abs() for i in range(d)], dim=1).permute(0, 2, 3, 1)
        cost = cost.reshape(b, h * w, d, d + 1)
    pred = attn_weight.permute(0, 2, 3, 1).reshape(b, h * w, d)
    log_mu, log_nu = (gt_dist + 1e-12).log(), (pred + 1e-12).log()
    u, v = torch.zeros_like(log_nu), torch.zeros_like(log_mu)
    for _ in range(iters):
        v = log_mu - torch.logsumexp(cost / eps + u.unsqueeze(3), dim=2)
        u = log_nu - torch.logsumexp(cost / eps + v.unsqueeze(2), dim=3)
    tmap = (cost / eps + u.unsqueeze(3) + v.unsqueeze(2)).exp()
    loss = (tmap * cost).reshape(b * h * w, -1)[mask.reshape(-1)].sum(-1).mean()
    return tmap, loss
