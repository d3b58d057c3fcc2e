#!/usr/bin/env python
"""FPN4 (5 views, 832x1152) and the whole MVS4net.forward with the finest top-down level evaluated directly
(ops.fpn_topdown) and through its linearity (ops.fpn_topdown_lin).  One JSON line.

    python scripts/bench_fpn.py [--iters 20]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import deep_reconstruction_with_epipolar_lines_mvster_b200 as mv  # noqa: E402
from deep_reconstruction_with_epipolar_lines_mvster_b200 import synthetic as syn  # noqa: E402
from scripts.bench_extra import NET_CFG, timed  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    h0, w0, n, b = 832, 1152, 5, 1
    model = mv.MVS4net(**NET_CFG).eval()
    model.load_state_dict(syn.fill_state_dict(model.state_dict(), seed=7))
    model = model.to(dev)
    gen = torch.Generator(device=dev).manual_seed(0)
    imgs = [torch.rand((b, 3, h0, w0), device=dev, generator=gen) for _ in range(n)]
    proj = {k: torch.from_numpy(v).to(dev) for k, v in syn.proj_matrices_all_stages(b, n, h0, w0).items()}
    dv = torch.from_numpy(syn.depth_values(b)).to(dev)
    res = {"bench": "fpn4_topdown_832x1152_n5"}
    feats = {}
    for lin in (False, True):
        model.feature.linear_topdown = lin
        tag = "linear" if lin else "direct"
        with torch.no_grad():
            res["fpn_%s_ms" % tag] = timed(lambda: model.extract_features(imgs), args.iters)
            res["forward_%s_ms" % tag] = timed(lambda: model(imgs, proj, dv), args.iters)
            feats[tag] = model.extract_features(imgs)
            out = model(imgs, proj, dv)
            res["depth_mean_%s" % tag] = float(out["stage4"]["depth"].mean())
    for k in ("stage3", "stage4"):
        f0, f1 = feats["direct"][0][k], feats["linear"][0][k]
        res[k + "_feature_max_abs_diff"] = float((f0 - f1).abs().max())
        res[k + "_feature_abs_max"] = float(f0.abs().max())
    print(json.dumps(res))


if __name__ == "__main__":
    main()
