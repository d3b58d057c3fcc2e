#!/usr/bin/env python
"""Golden fixture of one TRAINING step of the unmodified reference: ``MVS4net(...).train()`` forward (BatchNorm on batch
statistics), ``MVS4net_loss`` (Sinkhorn OT loss per stage, models/MVS4Net.py:195-240) and autograd backward, on CPU.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_train.py      ->  tests/golden/train.npz

Pins the whole differentiable chain of the B200 path at once: K1 backward, tail backward and the fused Sinkhorn
gradient, through cuDNN's FPN4 / reg2d.
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

from deep_reconstruction_with_epipolar_lines_mvster_b200 import synthetic as syn  # noqa: E402
from models.MVS4Net import MVS4net, MVS4net_loss  # noqa: E402  (the reference)

CFG = dict(arch_mode="fpn", reg_net="reg2d", num_stage=4, fpn_base_channel=8, reg_channel=8,
           stage_splits=[8, 8, 4, 4], depth_interals_ratio=[0.5, 0.5, 0.5, 1.0], group_cor=True,
           group_cor_dim=[8, 8, 4, 4], inverse_depth=True, agg_type="ConvBnReLU3D", attn_temp=2.0, attn_fuse_d=True)
LOSS_KW = dict(stage_lw=[1, 1, 1, 1], l1ot_lw=[0, 1], inverse_depth=True, ot_iter=10, ot_eps=1, ot_continous=False,
               mono=False)
WATCH = ["feature.conv0.0.conv.weight", "feature.conv2.1.conv.weight", "feature.out3.weight", "feature.inner1.bias",
         "reg.0.conv0.conv.weight", "reg.1.conv5.conv.weight", "reg.3.prob.weight", "reg.3.conv11.0.weight"]


def main():
    h0, w0, n, b = 64, 128, 3, 2
    model = MVS4net(**CFG).train()
    model.load_state_dict(syn.fill_state_dict(model.state_dict(), seed=7))
    names = dict(model.named_parameters())
    missing = [w for w in WATCH if w not in names]
    assert not missing, (missing, list(names)[:80])
    g = torch.Generator().manual_seed(33)
    imgs = [torch.rand(b, 3, h0, w0, generator=g) for _ in range(n)]
    proj = {k: torch.from_numpy(v) for k, v in syn.proj_matrices_all_stages(b, n, h0, w0).items()}
    dv = torch.from_numpy(syn.depth_values(b))
    gts, masks = {}, {}
    for s in range(4):
        h, w = h0 >> (3 - s), w0 >> (3 - s)
        yy, xx = torch.meshgrid(torch.linspace(0, 1, h), torch.linspace(0, 1, w), indexing="ij")
        gts["stage%d" % (s + 1)] = (560 + 200 * xx + 90 * yy)[None].repeat(b, 1, 1).contiguous()
        masks["stage%d" % (s + 1)] = (torch.rand(b, h, w, generator=g) > 0.25).float()
    out = model(imgs, proj, dv)
    total, l1s, ots, ratios = MVS4net_loss(out, gts, masks, **LOSS_KW)
    total.backward()
    rec = {"imgs": torch.stack(imgs, 0).numpy(), "depth_values": dv.numpy(), "total": total.detach().numpy(),
           "ots": np.array([float(o) for o in ots]), "ratios": np.array([float(r) for r in ratios]),
           "watch": np.array(WATCH)}
    for k in gts:
        rec["gt_" + k] = gts[k].numpy()
        rec["mask_" + k] = masks[k].numpy()
        rec["attn_" + k] = out[k]["attn_weight"].detach().numpy()
        rec["depth_" + k] = out[k]["depth"].detach().numpy()
    for wname in WATCH:
        rec["grad/" + wname] = names[wname].grad.numpy()
    gn = torch.sqrt(sum((p.grad ** 2).sum() for p in model.parameters() if p.grad is not None))
    rec["grad_norm"] = gn.numpy()
    np.savez_compressed(os.path.join(HERE, "train.npz"), **rec)
    print("train step: total %.6f, ot %s, ratios %s, |grad| %.4g, %.2f MB" % (
        float(total), rec["ots"], rec["ratios"], float(gn), os.path.getsize(os.path.join(HERE, "train.npz")) / 1e6))


if __name__ == "__main__":
    main()
