"""The reference itself as a timed baseline.  TEST / BASELINE INFRASTRUCTURE - only ``bench.py`` (``--impl reference``,
``cpu_baseline``, ``gpu_reference``) and ``tests/`` import this; the product never does.

``ReferenceCascade`` runs the hot-path slice of the UNMODIFIED reference (``oracle/_ref``, see ``make_ref.py``) through
its own public entry point ``MVS4net.forward`` (models/MVS4Net.py:60-186): the stage loop, ``init_inverse_range`` /
``schedule_inverse_range``, ``stagenet.forward`` with ``homo_warping``, the group correlation, the epipolar attention,
the aggregation and the depth / confidence tail are the reference's code, executed as the reference executes them.
The two sub-networks that are out of scope of the metric are swapped for stand-ins exactly as in the GPU arm:
``model.feature`` (FPN4) returns the pre-computed synthetic feature maps of the view, ``model.reg[k]`` (reg2d) returns
the pre-computed stand-in logits of stage k.  Nothing of this repository's kernels or oracle port is on that path.
"""
from __future__ import annotations

import importlib
import os
import sys
from typing import List, Optional, Sequence

import torch as _torch
import torch.nn as _nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "models", "mvs4net_utils.py"))


def load_reference():
    """Import ``models`` from ``oracle/_ref`` (the copy made by make_ref.py); None when it is absent."""
    if not available():
        return None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    sys.dont_write_bytecode = True
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):  # the reference prints a DCN notice at import
        return importlib.import_module("models")


class _Features(_nn.Module):
    """Stand-in for FPN4: ``imgs[v]`` is a [B, 1] tensor holding the view index; returns that view's feature dict."""

    def __init__(self, per_view, index_of):
        super().__init__()
        self.per_view = per_view
        self.index_of = index_of   # id(img tensor) -> view index: no device read-back inside the timed region

    def forward(self, img):
        return self.per_view[self.index_of[id(img)]]


class _Logits(_nn.Module):
    """Stand-in for reg2d: returns the stage's pre-computed logits (or, in the set-up pass, builds them from the
    hypotheses that ``stagenet`` was just called with)."""

    def __init__(self):
        super().__init__()
        self.fixed = None
        self.make = None
        self.hypo = None

    def forward(self, cor_feats):
        if self.fixed is None:
            self.fixed = self.make(self.hypo)
        return self.fixed


class ReferenceCascade:
    def __init__(self, torch, features, projs, depth_values, make_logits, groups: Sequence[int], ndepths: Sequence[int],
                 split_itv: Sequence[float], attn_temp: float = 2.0, device="cpu"):
        """features[s][v]: [B,C,H,W]; projs[s]: [B,N,2,4,4]; make_logits[s](hypo) -> [B,D,H,W] stand-in logits."""
        models = load_reference()
        if models is None:
            raise RuntimeError("oracle/_ref is missing: run python oracle/make_ref.py where /root/reference exists")
        self.torch = torch
        nstage, nviews = len(features), len(features[0])
        dev = torch.device(device)
        model = models.MVS4net(reg_net="reg2d", num_stage=nstage, stage_splits=list(ndepths),
                               depth_interals_ratio=list(split_itv), group_cor=True, group_cor_dim=list(groups),
                               inverse_depth=True, attn_temp=attn_temp).eval()
        per_view = [{"stage%d" % (s + 1): features[s][v].to(dev) for s in range(nstage)} for v in range(nviews)]
        self.imgs = [torch.full((features[0][0].shape[0], 1), float(v), device=dev) for v in range(nviews)]
        model.feature = _Features(per_view, {id(t): v for v, t in enumerate(self.imgs)})
        self.stubs = [_Logits() for _ in range(nstage)]
        for s, st in enumerate(self.stubs):
            st.make = make_logits[s]
        model.reg = _nn.ModuleList(self.stubs)  # indexed like the ModuleList it replaces (models/MVS4Net.py:121)
        self.model = model
        self.proj = {"stage%d" % (s + 1): projs[s].to(dev) for s in range(nstage)}
        self.dv = depth_values.to(dev)
        # set-up pass (untimed): stagenet is wrapped once so that each stub sees the hypotheses of its stage and
        # freezes its logits; afterwards the unwrapped reference runs on fixed logits, like the GPU arm
        sn = model.stagenet
        orig = sn.forward

        def recording(features, proj_matrices, depth_hypo, regnet, stage_idx, **kw):
            regnet.hypo = depth_hypo
            return orig(features, proj_matrices, depth_hypo=depth_hypo, regnet=regnet, stage_idx=stage_idx, **kw)

        sn.forward = recording
        with torch.no_grad():
            self.model(self.imgs, self.proj, self.dv)
        del sn.forward  # back to the class's own forward

    def run(self):
        with self.torch.no_grad():
            out = self.model(self.imgs, self.proj, self.dv)
        last = out["stage%d" % len(self.stubs)]
        return last["depth"], last["photometric_confidence"]
