"""Synthetic DTU-shaped inputs for the cost-volume hot path.

There is no dataset in the build or benchmark environment, so every test, the benchmark and the golden
fixtures use this rig: pinhole cameras on an arc around the world point (0, 0, 680) looking at a depth range of
425..935 (the DTU range, reference ``datasets/dataloaderWIP.py:28-29``), with the per-stage intrinsics scaling of
the reference eval loader (``datasets/dataloader_eval.py:269-287``: rows 0-1 of K divided by 8/4/2/1).

Only numpy/torch are used here; nothing in this file touches ``oracle/``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch

DTU_DEPTH_MIN = 425.0
DTU_DEPTH_MAX = 935.0

# per-stage (channels, groups, hypotheses) of the shipped configuration
# channels: reference models/mvs4net_utils.py:456-477 (FPN4, base 8); G, D: .vscode/launch.json:27,30
STAGE_CHANNELS = (64, 32, 16, 8)
STAGE_GROUPS = (8, 8, 4, 4)
STAGE_NDEPTHS = (8, 8, 4, 4)
STAGE_SPLIT_ITV = (0.5, 0.5, 0.5, 1.0)


def stage_shape(h0: int, w0: int, stage: int) -> Tuple[int, int]:
    """(H, W) of cascade stage ``stage`` (0..3) for a full-resolution H0 x W0 image."""
    s = 2 ** (3 - stage)
    return h0 // s, w0 // s


def intrinsics(h0: int, w0: int, stage: int = 3) -> np.ndarray:
    """3x3 K for ``stage`` (0 = 1/8 res .. 3 = full res); f follows the DTU 1600-px-wide calibration."""
    s = 2.0 ** (3 - stage)
    f = 2892.33 * (w0 / 1600.0) / s
    return np.array([[f, 0.0, (w0 / 2.0) / s], [0.0, f, (h0 / 2.0) / s], [0.0, 0.0, 1.0]], dtype=np.float64)


def extrinsics(view: int, step_rad: float = 0.06, centre: Sequence[float] = (0.0, 0.0, 680.0),
               tilt_rad: float = 0.0) -> np.ndarray:
    """4x4 world->camera for view ``view``: rotation about Y (and optionally X) around ``centre``."""
    a = step_rad * view
    ca, sa = math.cos(a), math.sin(a)
    ry = np.array([[ca, 0.0, sa], [0.0, 1.0, 0.0], [-sa, 0.0, ca]], dtype=np.float64)
    b = tilt_rad * view
    cb, sb = math.cos(b), math.sin(b)
    rx = np.array([[1.0, 0.0, 0.0], [0.0, cb, -sb], [0.0, sb, cb]], dtype=np.float64)
    r = rx @ ry
    c = np.asarray(centre, dtype=np.float64)
    e = np.eye(4, dtype=np.float64)
    e[:3, :3] = r
    e[:3, 3] = c - r @ c
    return e


def grid_extrinsics(idx: int, cols: int = 7, step_rad: float = 0.05) -> np.ndarray:
    """BDS8-style rig: cameras on a cols x cols grid of (yaw, pitch) angles around the scene centre."""
    i, j = divmod(idx, cols)
    yaw = (j - (cols - 1) / 2.0) * step_rad
    pitch = (i - (cols - 1) / 2.0) * step_rad
    cy, sy = math.cos(yaw), math.sin(yaw)
    cp, sp = math.cos(pitch), math.sin(pitch)
    ry = np.array([[cy, 0.0, sy], [0.0, 1.0, 0.0], [-sy, 0.0, cy]])
    rx = np.array([[1.0, 0.0, 0.0], [0.0, cp, -sp], [0.0, sp, cp]])
    r = rx @ ry
    c = np.array([0.0, 0.0, 680.0])
    e = np.eye(4)
    e[:3, :3] = r
    e[:3, 3] = c - r @ c
    return e


def proj_matrices(batch: int, nviews: int, h0: int, w0: int, stage: int, step_rad: float = 0.06,
                  per_batch_jitter: float = 0.0, tilt_rad: float = 0.0) -> np.ndarray:
    """``[B, N, 2, 4, 4]`` float32 in the reference loader's layout.

    ``[:, v, 0]`` is the 4x4 extrinsic, ``[:, v, 1, :3, :3]`` the intrinsic, all other entries zero
    (reference ``datasets/dataloader_eval.py:269-271``).
    """
    out = np.zeros((batch, nviews, 2, 4, 4), dtype=np.float32)
    k = intrinsics(h0, w0, stage)
    for b in range(batch):
        for v in range(nviews):
            out[b, v, 0] = extrinsics(v, step_rad * (1.0 + per_batch_jitter * b), tilt_rad=tilt_rad)
            out[b, v, 1, :3, :3] = k
    return out


def proj_matrices_all_stages(batch: int, nviews: int, h0: int, w0: int, **kw) -> Dict[str, np.ndarray]:
    return {"stage%d" % (s + 1): proj_matrices(batch, nviews, h0, w0, s, **kw) for s in range(4)}


def smooth_features(batch: int, channels: int, h: int, w: int, seed: int, device="cpu",
                    dtype=torch.float32) -> torch.Tensor:
    """``[B, C, H, W]`` FPN-like raw conv maps: N(0, 0.5^2) noise, 3x3 box-smoothed."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    x = torch.randn(batch, channels, h, w, generator=g, dtype=torch.float32) * 0.5
    x = torch.nn.functional.avg_pool2d(x, 3, stride=1, padding=1, count_include_pad=False) * 1.7
    return x.to(device=device, dtype=dtype)


def depth_values(batch: int) -> np.ndarray:
    return np.tile(np.array([[DTU_DEPTH_MIN, DTU_DEPTH_MAX]], dtype=np.float32), (batch, 1))


def smooth_depth_map(h: int, w: int, seed: int = 0, lo: float = 600.0, hi: float = 760.0) -> np.ndarray:
    """A smooth analytic surface (tilted plane + bump) in [lo, hi], float32 ``[H, W]``."""
    ys, xs = np.meshgrid(np.linspace(-1, 1, h), np.linspace(-1, 1, w), indexing="ij")
    rng = np.random.RandomState(seed)
    a, b = rng.uniform(-0.25, 0.25, size=2)
    bump = np.exp(-((xs - 0.2) ** 2 + (ys + 0.1) ** 2) / 0.18)
    z = 0.5 + a * xs + b * ys - 0.25 * bump
    z = (z - z.min()) / max(z.max() - z.min(), 1e-9)
    return (lo + (hi - lo) * z).astype(np.float32)


def render_surface_depths(k: np.ndarray, es: List[np.ndarray], h: int, w: int, noise_mm: float = 0.3,
                          seed: int = 0) -> np.ndarray:
    """Depth maps ``[V, H, W]`` float32 of the world surface z = 680 + 40*sin(x/90)*cos(y/70) seen from each camera.

    Ray/surface intersection by fixed-point iteration (the surface is nearly fronto-parallel); Gaussian noise of
    ``noise_mm`` is added, as a depth estimator would.
    """
    rng = np.random.RandomState(seed)
    ys, xs = np.meshgrid(np.arange(h, dtype=np.float64), np.arange(w, dtype=np.float64), indexing="ij")
    pix = np.stack([xs.ravel(), ys.ravel(), np.ones(h * w)])
    kinv = np.linalg.inv(k)
    out = np.zeros((len(es), h, w), dtype=np.float32)
    for v, e in enumerate(es):
        einv = np.linalg.inv(e)
        rays_c = kinv @ pix                       # camera-frame rays with z = 1
        d = np.full(h * w, 680.0)
        for _ in range(12):
            pc = rays_c * d                       # camera frame
            pw = einv[:3, :3] @ pc + einv[:3, 3:4]
            zs = 680.0 + 40.0 * np.sin(pw[0] / 90.0) * np.cos(pw[1] / 70.0)
            # move along the ray so that world z matches the surface
            dz = zs - pw[2]
            d = d + dz / np.maximum((einv[:3, :3] @ rays_c)[2], 1e-3)
        out[v] = (d + rng.normal(0.0, noise_mm, size=d.shape)).reshape(h, w).astype(np.float32)
    return out


def pair_list(nviews: int, nsrc: int) -> np.ndarray:
    """``[V, nsrc]`` int32: for each ref view the ``nsrc`` nearest other views by index distance."""
    pairs = np.zeros((nviews, nsrc), dtype=np.int32)
    for r in range(nviews):
        others = sorted((v for v in range(nviews) if v != r), key=lambda v: (abs(v - r), v))
        pairs[r] = others[:nsrc]
    return pairs


def fill_state_dict(state_dict, seed: int = 0):
    """Deterministic, framework-independent synthetic weights for a (reference or B200) ``MVS4net``: every entry is
    drawn from a NumPy stream seeded by ``crc32(key) + seed``, so the same recipe applied to two models with the same
    parameter names gives identical weights - no checkpoint has to be stored with the golden fixtures.
    Conv weights ~ N(0, 2/fan_in) (N(0, 0.5/fan_in) for the layers without ReLU); BatchNorm weight ~ U(0.8,1.2), bias / running_mean ~ N(0,0.1),
    running_var ~ U(0.5,1.5); conv biases ~ N(0,0.05)."""
    import zlib

    import torch
    out = {}
    for key, val in state_dict.items():
        rng = np.random.RandomState((zlib.crc32(key.encode()) + seed) % (2 ** 31))
        shape = tuple(val.shape)
        if key.endswith("num_batches_tracked"):
            arr = np.zeros(shape, dtype=np.int64)
        elif key.endswith("running_var"):
            arr = rng.uniform(0.5, 1.5, size=shape)
        elif key.endswith("running_mean"):
            arr = rng.normal(0.0, 0.1, size=shape)
        elif key.endswith("weight") and len(shape) == 1:
            arr = rng.uniform(0.8, 1.2, size=shape)
        elif key.endswith("weight"):
            fan_in = int(np.prod(shape[1:]))
            # layers without a ReLU behind them (FPN lateral / output convs, prob) get gain 1/2 so that features and
            # logits stay O(1) through the whole network
            linear = any(t in key for t in (".inner", ".out", ".prob."))
            arr = rng.normal(0.0, np.sqrt((0.5 if linear else 2.0) / fan_in), size=shape)
        elif ".bn." in key or key.split(".")[-2].isdigit():
            arr = rng.normal(0.0, 0.1, size=shape)
        else:
            arr = rng.normal(0.0, 0.05, size=shape)
        out[key] = torch.from_numpy(np.asarray(arr)).to(val.dtype)
    return out


def filter_fixture_512(views: int = 3, h: int = 512, w: int = 640):
    """Deterministic inputs of the full-size filter golden (``tests/golden/filter512.npz`` holds only the reference's
    OUTPUTS for these inputs plus a checksum of them): ``views`` cameras of the 0.05-rad rig looking at the analytic
    surface, depth noise 0.3 mm, a hole in every depth map and one inconsistent patch."""
    k = intrinsics(h, w, 3)
    es = [extrinsics(i, 0.05, tilt_rad=0.01) for i in range(views)]
    depths = render_surface_depths(k, es, h, w, noise_mm=0.3, seed=7)
    for v in range(views):
        depths[v, 40 + 30 * v:60 + 30 * v, 100 + 50 * v:140 + 50 * v] = 0.0
    depths[1, 300:340, 200:280] += 20.0
    conf = np.random.RandomState(5).uniform(0, 1, size=(views, h, w)).astype(np.float32)
    pairs = np.concatenate([np.arange(views)[:, None], pair_list(views, views - 1)], 1).astype(np.int32)
    return depths, conf, np.stack([k] * views), np.stack(es), pairs
