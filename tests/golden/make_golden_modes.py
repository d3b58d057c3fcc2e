#!/usr/bin/env python
"""Golden volumes and autograd gradients of the reference's ``group_cor=False`` / ``attn_fuse_d=False`` options
(models/mvs4net_utils.py:1071,1078-1081,1098), generated from the UNMODIFIED reference.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_modes.py

Inputs are those already frozen in ``k1_stage2.npz`` / ``k1_stage4.npz`` / ``k1_oob.npz`` (features, projections,
hypotheses); for every fixture and each of the three option combinations the reference ``stagenet`` is run with a
recording regnet, a seeded upstream gradient is back-propagated through its autograd graph, and volume + gradients
are stored in ``k1_modes.npz`` under ``<fixture>/<gc><fd>/...`` keys.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MVSTER_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import models.mvs4net_utils as U  # noqa: E402  (the reference)


class RecordingRegnet(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.volume = None

    def forward(self, x):
        self.volume = x
        return x.sum(1)


def main():
    out = {}
    for name in ["k1_stage2", "k1_stage4", "k1_oob"]:
        g = np.load(os.path.join(HERE, name + ".npz"))
        groups, temp = int(g["groups"]), float(g["attn_temp"])
        for gc, fd in [(False, True), (True, False), (False, False)]:
            feats = [torch.from_numpy(g["ref"]).clone().requires_grad_(True)]
            feats += [torch.from_numpy(g["srcs"][:, v]).clone().requires_grad_(True) for v in range(g["srcs"].shape[1])]
            net = U.stagenet(inverse_depth=True, mono=False, attn_fuse_d=fd, vis_ETA=False, attn_temp=temp).eval()
            reg = RecordingRegnet()
            net(feats, torch.from_numpy(g["proj"]), torch.from_numpy(g["hypo"]), reg, 1, group_cor=gc,
                group_cor_dim=groups, split_itv=1.0)
            vol = reg.volume
            gout = torch.randn(vol.shape, generator=torch.Generator().manual_seed(7 + 2 * int(gc) + int(fd)))
            (vol * gout).sum().backward()
            key = "%s/%d%d/" % (name, int(gc), int(fd))
            out[key + "volume"] = vol.detach().numpy()
            out[key + "gout"] = gout.numpy()
            out[key + "grad_ref"] = feats[0].grad.numpy()
            out[key + "grad_srcs"] = np.stack([f.grad.numpy() for f in feats[1:]], 1)
            print(key, tuple(vol.shape), "absmax %.4f" % vol.abs().max().item(),
                  "grad_ref absmax %.4f" % feats[0].grad.abs().max().item())
    np.savez_compressed(os.path.join(HERE, "k1_modes.npz"), **out)


if __name__ == "__main__":
    main()
