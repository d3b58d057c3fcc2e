#!/usr/bin/env python
"""Golden fixture for the widened scope: FPN4, reg2d and a whole ``MVS4net`` forward of the UNMODIFIED
reference on CPU, with weights from the deterministic recipe ``synthetic.fill_state_dict`` (so no checkpoint is stored).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_network.py      ->  tests/golden/network.npz
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
from models.MVS4Net import MVS4net  # noqa: E402  (the reference)

CFG = dict(arch_mode="fpn", reg_net="reg2d", num_stage=4, fpn_base_channel=8, reg_channel=8,
           stage_splits=[8, 8, 4, 4], depth_interals_ratio=[0.5, 0.5, 0.5, 1.0], group_cor=True,
           group_cor_dim=[8, 8, 4, 4], inverse_depth=True, agg_type="ConvBnReLU3D", attn_temp=2.0, attn_fuse_d=True)


def main():
    h0, w0, n = 64, 128, 3
    model = MVS4net(**CFG).eval()
    model.load_state_dict(syn.fill_state_dict(model.state_dict(), seed=7))
    rec = {"state_keys": np.array(sorted(model.state_dict().keys()))}
    g = torch.Generator().manual_seed(21)
    imgs = [torch.rand(1, 3, h0, w0, generator=g) for _ in range(n)]
    proj = {k: torch.from_numpy(v) for k, v in syn.proj_matrices_all_stages(1, n, h0, w0).items()}
    dv = torch.from_numpy(syn.depth_values(1))
    rec["imgs"] = torch.stack(imgs, 0).numpy()

    # regnet taps: input volume and logits of every stage, plus the input of conv11 / the conv0 skip of stage 4
    taps = {}
    for s, reg in enumerate(model.reg):
        def hook(mod, inp, out, s=s):
            taps["s%d_volume" % (s + 1)] = inp[0].detach().numpy().copy()
            taps["s%d_logits" % (s + 1)] = out.detach().numpy().copy()
        reg.register_forward_hook(hook)
    model.reg[3].conv11.register_forward_hook(lambda m, i, o: taps.__setitem__("s4_low", i[0].detach().numpy().copy()))
    model.reg[3].conv0.register_forward_hook(lambda m, i, o: taps.__setitem__("s4_skip", o.detach().numpy().copy()))
    with torch.no_grad():
        feats = model.feature(imgs[1])
        for k, v in feats.items():
            rec["fpn_view1_" + k] = v.numpy()
        out = model(imgs, proj, dv)
    rec.update(taps)
    for st, d in out.items():
        for name, val in d.items():
            rec["%s_%s" % (st, name)] = val.numpy()
    rec["depth_values"] = dv.numpy()
    np.savez_compressed(os.path.join(HERE, "network.npz"), **rec)
    print("network: %d arrays, %.2f MB; stage-4 depth %.1f..%.1f, conf %.3f..%.3f" % (
        len(rec), os.path.getsize(os.path.join(HERE, "network.npz")) / 1e6, float(rec["stage4_depth"].min()),
        float(rec["stage4_depth"].max()), float(np.nanmin(rec["stage4_photometric_confidence"])),
        float(np.nanmax(rec["stage4_photometric_confidence"]))))
    srt = np.sort(rec["stage4_attn_weight"], 1)
    print("stage-4 top-2 attention gap: median %.3g, frac < 1e-6: %.4f" % (
        float(np.median(srt[:, -1] - srt[:, -2])), float(((srt[:, -1] - srt[:, -2]) < 1e-6).mean())))


if __name__ == "__main__":
    main()
