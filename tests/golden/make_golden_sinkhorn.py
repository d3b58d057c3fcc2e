#!/usr/bin/env python
"""Golden fixtures for the fused Sinkhorn loss (K3) from the UNMODIFIED reference.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_sinkhorn.py

Imports ``sinkhorn`` (models/mvs4net_utils.py:1164) and ``MVS4net_loss`` (models/MVS4Net.py:195) from /root/reference,
runs them on seeded synthetic stage outputs and freezes inputs, T_map, loss, the autograd gradient w.r.t.
``attn_weight`` and the loss function's statistics in ``sinkhorn.npz``.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MVSTER_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, REF)

import models.mvs4net_utils as U  # noqa: E402  (the reference)
from models.MVS4Net import MVS4net_loss  # noqa: E402


def stage_inputs(seed, b, d, h, w, sharp=1.0):
    """hypotheses uniform in inverse depth around a smooth surface, gt near it, softmax attention, ragged mask."""
    g = torch.Generator().manual_seed(seed)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, h), torch.linspace(0, 1, w), indexing="ij")
    surf = 600 + 120 * xx + 60 * yy + 15 * torch.sin(6 * xx)
    centre = surf[None].repeat(b, 1, 1) * (1 + 0.01 * torch.randn(b, h, w, generator=g))
    itv = 2.5e-5 * (1 + 0.2 * torch.rand(b, h, w, generator=g))
    steps = torch.arange(d, dtype=torch.float32).view(1, d, 1, 1) - (d - 1) / 2
    hypo = 1.0 / (1.0 / centre[:, None] - steps * itv[:, None])          # index 0 = farthest, like the reference
    gt = 1.0 / (1.0 / centre + (torch.rand(b, h, w, generator=g) - 0.5) * (d + 4) * itv)   # some pixels out of range
    attn = torch.softmax(sharp * torch.randn(b, d, h, w, generator=g), dim=1)
    mask = torch.rand(b, h, w, generator=g) > 0.3
    gt = torch.where(mask, gt, torch.zeros_like(gt))                      # unmasked gt = 0, as in the DTU loader
    return gt, hypo, attn, mask


def main():
    out = {}
    cases = [  # name, D, iters, eps, continuous
        ("d4_it3_e1", 4, 3, 1.0, False),
        ("d8_it3_e1", 8, 3, 1.0, False),
        ("d8_it10_e01", 8, 10, 0.1, False),
        ("d4_it10_e1_cont", 4, 10, 1.0, True),
        ("d8_it3_e1_cont", 8, 3, 1.0, True),
        ("d4_it0_e1", 4, 0, 1.0, False),
    ]
    for k, (name, d, iters, eps, cont) in enumerate(cases):
        gt, hypo, attn, mask = stage_inputs(100 + k, 2, d, 12, 20, sharp=2.0 if k % 2 else 1.0)
        a = attn.clone().requires_grad_(True)
        tmap, loss = U.sinkhorn(gt, hypo, a, mask, iters=iters, eps=eps, continuous=cont)
        grad = torch.autograd.grad(loss, a)[0] if loss.requires_grad else torch.zeros_like(a)  # iters=0: no dependence
        for key, val in (("gt", gt), ("hypo", hypo), ("attn", attn), ("mask", mask), ("tmap", tmap.detach()),
                         ("loss", loss.detach()), ("grad", grad)):
            out["%s/%s" % (name, key)] = val.numpy()
        out["%s/cfg" % name] = np.array([d, iters, eps, float(cont)], dtype=np.float64)
        print(name, float(loss), float(grad.abs().max()))
    out["cases"] = np.array([c[0] for c in cases])

    # the whole loss function over four stages (default kwargs of train_mvs4.py: ot_iter=10, inverse_depth=True)
    inputs, gts, masks = {}, {}, {}
    for s, (d, h, w) in enumerate([(8, 6, 8), (8, 12, 16), (4, 24, 32), (4, 48, 64)]):
        gt, hypo, attn, mask = stage_inputs(200 + s, 1, d, h, w)
        key = "stage%d" % (s + 1)
        inputs[key] = {"depth": torch.gather(hypo, 1, attn.argmax(1, keepdim=True)).squeeze(1), "hypo_depth": hypo,
                       "attn_weight": attn.clone().requires_grad_(True)}
        gts[key], masks[key] = gt, mask.float()
    kw = dict(stage_lw=[1, 1, 1, 1], l1ot_lw=[0, 1], inverse_depth=True, ot_iter=10, ot_eps=1, ot_continous=False,
              mono=False)
    total, l1s, ots, ratios = MVS4net_loss(inputs, gts, masks, **kw)
    grads = torch.autograd.grad(total, [inputs["stage%d" % (s + 1)]["attn_weight"] for s in range(4)])
    for s in range(4):
        key = "stage%d" % (s + 1)
        out["loss4/%s/gt" % key] = gts[key].numpy()
        out["loss4/%s/mask" % key] = masks[key].numpy()
        out["loss4/%s/hypo" % key] = inputs[key]["hypo_depth"].numpy()
        out["loss4/%s/attn" % key] = inputs[key]["attn_weight"].detach().numpy()
        out["loss4/%s/depth" % key] = inputs[key]["depth"].numpy()
        out["loss4/%s/grad" % key] = grads[s].numpy()
        out["loss4/%s/ot" % key] = ots[s].detach().numpy()
        out["loss4/%s/ratio" % key] = ratios[s].numpy()
    out["loss4/total"] = total.detach().numpy()
    print("MVS4net_loss", float(total), [float(o) for o in ots], [float(r) for r in ratios])
    np.savez_compressed(os.path.join(HERE, "sinkhorn.npz"), **out)


if __name__ == "__main__":
    main()
