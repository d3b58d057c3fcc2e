"""Geometric / photometric consistency filter (reference test_mvs4.py:612-670 and :716-749) on the GPU.

``check_geometric_consistency`` keeps the reference's call signature and return tuple; the thresholds the reference
reads from its module-global ``args`` (test_mvs4.py:667,716,746) live in ``FilterConfig`` here.
``filter_scene`` fuses all reference views of a scene in one launch, with the depth/confidence stack resident on the
device (the reference round-trips through .pfm files between depth generation and filtering).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import numpy as np
import torch

from . import ops


@dataclass
class FilterConfig:
    """Defaults follow the reference's live launch config (.vscode/launch.json:206-211)."""
    condmask_pixel: float = 1.0
    condmask_depth: float = 0.01
    photomask: float = 0.75
    geomask: int = 2


args = FilterConfig()  # module-level, like the reference's global ``args``


def _dev_f32(a, device):
    if isinstance(a, torch.Tensor):
        return a.to(device=device, dtype=torch.float32)
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(device)


def check_geometric_consistency(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src,
                                config: FilterConfig = None, device="cuda"):
    """test_mvs4.py:653.  Depth maps may be numpy arrays (as in the reference) or CUDA tensors; returns
    ``(mask, depth_reprojected, x2d_src, y2d_src)`` in the same kind (numpy in -> numpy out)."""
    cfg = config or args
    as_numpy = not isinstance(depth_ref, torch.Tensor)
    dr, ds = _dev_f32(depth_ref, device), _dev_f32(depth_src, device)
    mask, drep, x2d, y2d = ops.geo_check_pair(dr, intrinsics_ref, extrinsics_ref, ds, intrinsics_src, extrinsics_src,
                                              cfg.condmask_pixel, cfg.condmask_depth)
    if as_numpy:
        return mask.cpu().numpy(), drep.cpu().numpy(), x2d.cpu().numpy(), y2d.cpu().numpy()
    return mask, drep, x2d, y2d


def filter_scene(depths, confs, intrinsics, extrinsics, pairs: Sequence, config: FilterConfig = None, device="cuda",
                 want_geo_sum: bool = False):
    """Mask fusion of ``filter_depth`` (test_mvs4.py:694-749) for all reference views of ``pairs`` at once.

    ``depths``/``confs`` [V,H,W]; ``intrinsics`` [V,3,3], ``extrinsics`` [V,4,4] (float64); ``pairs`` either an int
    array [R, 1+S] or the reference's ``read_pair_file`` structure ``[(ref, [src...]), ...]`` (truncate the source
    lists to NviewFilter-1 beforehand, as test_mvs4.py:698 does).
    Returns ``(photo_mask, geo_mask, final_mask, depth_est_averaged, geo_mask_sum or None)`` as CUDA tensors.
    """
    cfg = config or args
    if not isinstance(pairs, np.ndarray):
        s = max(len(src) for _, src in pairs)
        arr = -np.ones((len(pairs), 1 + s), dtype=np.int32)
        for i, (ref, src) in enumerate(pairs):
            arr[i, 0] = ref
            arr[i, 1:1 + len(src)] = src
        pairs = arr
    dz, cf = _dev_f32(depths, device), _dev_f32(confs, device)
    return ops.geo_filter(dz, cf, np.asarray(intrinsics, np.float64), np.asarray(extrinsics, np.float64), pairs,
                          cfg.condmask_pixel, cfg.condmask_depth, cfg.photomask, cfg.geomask, want_geo_sum)
