"""Depth-map I/O and point-cloud fusion around the filter.

The reference writes every depth / confidence map to a ``.pfm`` file after depth generation
(``test_mvs4.py:454-495``, ``datasets/data_io.py:44-71``), reads them back for the filter (``:684-700``), builds the
world points of each reference view with NumPy (``depth2pts_np``, ``:206-229``, ``:786-793``) and writes one ``.ply``.
Here the depth / confidence stack stays on the GPU from ``MVS4net`` to the filter (``filter_scene``) to the
back-projection (``mvster_depth2pts``); files are written only if asked for, in the reference's formats:

  * ``save_pfm`` / ``read_pfm``  - byte-compatible with ``datasets/data_io.py:6-71``
  * ``write_ply``               - binary little-endian ``vertex`` element (x, y, z float32; red, green, blue uint8),
                                  what the reference's ``PlyData([PlyElement.describe(vertex_all, 'vertex')])`` emits
"""
from __future__ import annotations

import re
import sys
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from .filter import FilterConfig, filter_scene


# ----------------------------------------------------------------------------------------------------------------------
# file formats (host side)
# ----------------------------------------------------------------------------------------------------------------------
def save_pfm(filename: str, image: np.ndarray, scale: float = 1) -> None:
    """reference ``datasets/data_io.py:44-71``: rows bottom-up, ``Pf``/``PF`` header, negative scale = little endian."""
    image = np.asarray(image)
    if image.dtype.name != "float32":
        raise Exception("Image dtype must be float32.")
    if image.ndim == 3 and image.shape[2] == 3:
        color = True
    elif image.ndim == 2 or (image.ndim == 3 and image.shape[2] == 1):
        color = False
    else:
        raise Exception("Image must have H x W x 3, H x W x 1 or H x W dimensions.")
    image = np.flipud(image)
    endian = image.dtype.byteorder
    if endian == "<" or (endian == "=" and sys.byteorder == "little"):
        scale = -scale
    with open(filename, "wb") as f:
        f.write(b"PF\n" if color else b"Pf\n")
        f.write(("%d %d\n" % (image.shape[1], image.shape[0])).encode("utf-8"))
        f.write(("%f\n" % scale).encode("utf-8"))
        image.tofile(f)


def read_pfm(filename: str) -> Tuple[np.ndarray, float]:
    """reference ``datasets/data_io.py:6-41``: returns ``(data, scale)`` with rows top-down."""
    with open(filename, "rb") as f:
        header = f.readline().decode("utf-8").rstrip()
        if header not in ("PF", "Pf"):
            raise Exception("Not a PFM file.")
        m = re.match(r"^(\d+)\s(\d+)\s$", f.readline().decode("utf-8"))
        if not m:
            raise Exception("Malformed PFM header.")
        width, height = map(int, m.groups())
        scale = float(f.readline().rstrip())
        endian = "<" if scale < 0 else ">"
        data = np.fromfile(f, endian + "f")
    shape = (height, width, 3) if header == "PF" else (height, width)
    return np.flipud(np.reshape(data, shape)), abs(scale)


def write_ply(filename: str, xyz: np.ndarray, rgb: np.ndarray) -> None:
    """Binary little-endian PLY with one ``vertex`` element: x, y, z (float32), red, green, blue (uint8)."""
    xyz = np.asarray(xyz, dtype=np.float32).reshape(-1, 3)
    rgb = np.asarray(rgb, dtype=np.uint8).reshape(-1, 3)
    if len(xyz) != len(rgb):
        raise ValueError("write_ply: %d points but %d colours" % (len(xyz), len(rgb)))
    vert = np.empty(len(xyz), dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("red", "u1"), ("green", "u1"),
                                      ("blue", "u1")])
    vert["x"], vert["y"], vert["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    vert["red"], vert["green"], vert["blue"] = rgb[:, 0], rgb[:, 1], rgb[:, 2]
    header = ("ply\nformat binary_little_endian 1.0\nelement vertex %d\nproperty float x\nproperty float y\n"
              "property float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n" % len(vert))
    with open(filename, "wb") as f:
        f.write(header.encode("ascii"))
        vert.tofile(f)


# ----------------------------------------------------------------------------------------------------------------------
# GPU-resident fusion
# ----------------------------------------------------------------------------------------------------------------------
def depth2pts(depth_map, cam_intrinsic, cam_extrinsic, device="cuda"):
    """reference ``depth2pts_np`` (test_mvs4.py:206-218): world points ``[H*W, 3]`` float64 of a depth map (pixel
    centres at +0.5).  NumPy in -> NumPy out, CUDA tensor in -> CUDA tensor out."""
    as_numpy = not isinstance(depth_map, torch.Tensor)
    dz = torch.from_numpy(np.ascontiguousarray(depth_map, dtype=np.float32)).to(device) if as_numpy else depth_map
    xyz = ops.depth2pts(dz, cam_intrinsic, cam_extrinsic)
    return xyz.cpu().numpy() if as_numpy else xyz


def fuse_scene(depths, confs, intrinsics, extrinsics, pairs: Sequence, images: Optional[Sequence] = None,
               config: FilterConfig = None, device="cuda"):
    """The body of the reference's ``filter_depth`` loop (test_mvs4.py:694-793) for a whole scene, GPU-resident:
    photometric + geometric masks and averaged depth for every reference view (one launch), then per view the world
    points of the averaged depth and the masked selection (row-major pixel order, as ``xyz_world[final_mask.flatten()]``).

    ``images`` (optional) ``[V,H,W,3]`` in [0,1] (the reference's ``ref_img``): colours ``(img[mask] * 255).astype(uint8)``.
    Returns ``(vertices float64 [P,3], colors uint8 [P,3] or None, masks dict)`` as CUDA tensors; ``P`` = kept points of
    all reference views, concatenated in ``pairs`` order like the reference's ``vertices`` list."""
    photo, geo, final, avg, _ = filter_scene(depths, confs, intrinsics, extrinsics, pairs, config, device)
    refs = [int(p[0]) for p in pairs]
    ks, es = np.asarray(intrinsics, np.float64), np.asarray(extrinsics, np.float64)
    verts, cols = [], []
    for i, r in enumerate(refs):
        m = final[i].reshape(-1).bool()
        verts.append(ops.depth2pts(avg[i], ks[r], es[r])[m])
        if images is not None:
            img = images[r]
            img = img if isinstance(img, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(img))
            img = img.to(device=avg.device, dtype=torch.float32).reshape(-1, 3)[m]
            cols.append((img * 255).to(torch.uint8))       # (xyz_color_masked * 255).astype(np.uint8), :793
    vertices = torch.cat(verts, 0) if verts else torch.empty((0, 3), dtype=torch.float64, device=device)
    colors = torch.cat(cols, 0) if cols else None
    return vertices, colors, {"photo": photo, "geo": geo, "final": final, "depth_est_averaged": avg, "refs": refs}
