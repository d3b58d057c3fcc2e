#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the UNMODIFIED reference implementation.

Run in the build container (where /root/reference is mounted):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference is imported, never copied: ``models.mvs4net_utils`` / ``models.MVS4Net`` as modules, and the filter
functions of ``test_mvs4.py`` (which is not importable: it needs open3d/plyfile and parses argv at import time) are
lifted at run time with ``ast`` and executed unmodified.  Inputs are seeded synthetic tensors from the package's rig
(``synthetic.py``); inputs and reference outputs are frozen as ``*.npz`` so that the CPU oracle tests and the GPU
parity tests can run where the reference does not exist (the GPU box).
"""
from __future__ import annotations

import ast
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("MVSTER_REFERENCE", "/root/reference")
sys.dont_write_bytecode = True
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

from deep_reconstruction_with_epipolar_lines_mvster_b200 import synthetic as syn  # noqa: E402

import models.mvs4net_utils as U  # noqa: E402  (the reference)
from models.MVS4Net import MVS4net  # noqa: E402


class RecordingRegnet(torch.nn.Module):
    """Stands in for ``regnet``: stores the aggregated volume it is given, returns fixed or derived logits."""

    def __init__(self, logits=None):
        super().__init__()
        self.volume = None
        self.logits = logits

    def forward(self, x):
        self.volume = x
        return x.sum(1) if self.logits is None else self.logits


def reference_weights(features, proj, hypo, attn_temp, groups):
    """Per-view cor_weight by re-running the reference's own loop body (mvs4net_utils.py:1047-1083)."""
    views = torch.unbind(proj, 1)
    ref, srcs = features[0], features[1:]
    b, d, h, w = hypo.shape
    c = ref.shape[1]
    ref_volume = ref.unsqueeze(2).repeat(1, 1, d, 1, 1).reshape(b, groups, c // groups, d, h, w)
    out = []
    rp = views[0][:, 0].clone()
    rp[:, :3, :4] = torch.matmul(views[0][:, 1, :3, :3], views[0][:, 0, :3, :4])
    for src, pv in zip(srcs, views[1:]):
        sp = pv[:, 0].clone()
        sp[:, :3, :4] = torch.matmul(pv[:, 1, :3, :3], pv[:, 0, :3, :4])
        warped = U.homo_warping(src, sp, rp, hypo)
        cor = (warped.reshape(b, groups, c // groups, d, h, w) * ref_volume).mean(2)
        out.append(torch.softmax(cor.sum(1) / attn_temp, 1) / np.sqrt(c))
    return torch.stack(out, 1)


def hypotheses(b, d, h, w, stage, seed):
    dv = torch.from_numpy(syn.depth_values(b))
    if stage == 0:
        return U.init_inverse_range(dv, d, "cpu", torch.float32, h, w)
    lo = np.stack([syn.smooth_depth_map(h // 2, w // 2, seed + i) for i in range(b)])
    inv = torch.from_numpy(1.0 / lo)
    half = 0.5 * (1.0 / 425.0 - 1.0 / 935.0) / (2.0 ** stage) / 4.0
    return U.schedule_inverse_range(inv + half, inv - half, d, h, w)


def k1_case(name, b, n, c, g, d, h, w, stage, h0, w0, hs=None, ws=None, step=0.06, jitter=0.0, tilt=0.0,
            attn_temp=2.0, seed=0):
    hs = hs or h
    ws = ws or w
    torch.manual_seed(seed)
    feats = [syn.smooth_features(b, c, h, w, 1234 + seed)]
    feats += [syn.smooth_features(b, c, hs, ws, 4321 + seed + 17 * v) for v in range(1, n)]
    feats = [f.clone().requires_grad_(True) for f in feats]
    proj = torch.from_numpy(syn.proj_matrices(b, n, h0, w0, stage, step_rad=step, per_batch_jitter=jitter,
                                              tilt_rad=tilt))
    hypo = hypotheses(b, d, h, w, stage, seed)
    net = U.stagenet(inverse_depth=True, mono=False, attn_fuse_d=True, vis_ETA=False, attn_temp=attn_temp).eval()
    reg = RecordingRegnet()
    net(feats, proj, hypo, reg, stage, group_cor=True, group_cor_dim=g, split_itv=1.0)
    vol = reg.volume
    gout = torch.randn(vol.shape, generator=torch.Generator().manual_seed(99 + seed))
    (vol * gout).sum().backward()
    with torch.no_grad():
        wts = reference_weights([f.detach() for f in feats], proj, hypo, attn_temp, g)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        ref=feats[0].detach().numpy(), srcs=np.stack([f.detach().numpy() for f in feats[1:]], 1),
        proj=proj.numpy(), hypo=hypo.numpy(), groups=g, attn_temp=attn_temp,
        volume=vol.detach().numpy(), weights=wts.numpy(), gout=gout.numpy(),
        grad_ref=feats[0].grad.numpy(), grad_srcs=np.stack([f.grad.numpy() for f in feats[1:]], 1))
    print(name, "volume", tuple(vol.shape), "absmax %.4f" % vol.abs().max().item())


def warp_case():
    b, c, h, w, d = 2, 8, 12, 16, 4
    src = syn.smooth_features(b, c, h + 3, w + 5, 7)
    proj = torch.from_numpy(syn.proj_matrices(b, 2, 12, 16, 3, per_batch_jitter=0.3))
    p = []
    for v in range(2):
        pv = proj[:, v, 0].clone()
        pv[:, :3, :4] = torch.matmul(proj[:, v, 1, :3, :3], proj[:, v, 0, :3, :4])
        p.append(pv)
    hypo = hypotheses(b, d, h, w, 3, 5)
    out = U.homo_warping(src, p[1], p[0], hypo)
    np.savez_compressed(os.path.join(HERE, "warp.npz"), src=src.numpy(), src_proj=p[1].numpy(), ref_proj=p[0].numpy(),
                        hypo=hypo.numpy(), warped=out.numpy())
    print("warp", tuple(out.shape))


def tail_case():
    b, d, h, w = 2, 8, 10, 12
    g = torch.Generator().manual_seed(5)
    logits = torch.randn(b, d, h, w, generator=g) * 2.0 + 0.3
    logits[0, :, 0, 0] = 1.0                      # exact tie -> first index
    logits[0, 3, 0, 1] = logits[0, 5, 0, 1] = 9.0  # two-way tie
    hypo = hypotheses(b, d, h, w, 0, 0) * (1.0 + 0.01 * torch.rand(b, d, h, w, generator=g))
    feats = [syn.smooth_features(b, 8, h, w, s) for s in range(2)]
    proj = torch.from_numpy(syn.proj_matrices(b, 2, h, w, 3))
    out = {}
    for mode in ("eval", "train"):
        net = U.stagenet(inverse_depth=True, mono=True, attn_fuse_d=True, attn_temp=2.0)
        net = net.eval() if mode == "eval" else net.train()
        ret = net(feats, proj, hypo, RecordingRegnet(logits), 1, group_cor=True, group_cor_dim=4, split_itv=0.5)
        for k, v in ret.items():
            out[mode + "_" + k] = v.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "tail.npz"), logits=logits.numpy(), hypo=hypo.numpy(), split_itv=0.5, **out)
    print("tail keys", sorted(out))


def schedule_case():
    dv = torch.tensor([[425.0, 935.0], [300.0, 500.0, 1200.0]][0:1] * 2)
    dv[1] = torch.tensor([310.0, 1210.0])
    init = U.init_inverse_range(dv, 8, "cpu", torch.float32, 6, 7)
    g = torch.Generator().manual_seed(3)
    h, w = 14, 18
    centre = 1.0 / (500.0 + 300.0 * torch.rand(2, h // 2, w // 2, generator=g))
    itv = 2e-5 * (1 + torch.rand(2, h // 2, w // 2, generator=g))
    sched4 = U.schedule_inverse_range(centre + itv, centre - itv, 4, h, w)
    sched8 = U.schedule_inverse_range(centre + itv, centre - itv, 8, h, w)
    np.savez_compressed(os.path.join(HERE, "schedule.npz"), depth_values=dv.numpy(), init=init.numpy(),
                        inv_min=(centre + itv).numpy(), inv_max=(centre - itv).numpy(), sched4=sched4.numpy(),
                        sched8=sched8.numpy())
    print("schedule", tuple(init.shape), tuple(sched4.shape))


def lift_filter_functions(condmask_pixel, condmask_depth):
    import cv2
    src = open(os.path.join(REF, "test_mvs4.py")).read()
    tree = ast.parse(src)
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef)
            and n.name in ("reproject_with_depth", "check_geometric_consistency")]
    ns = {"np": np, "cv2": cv2,
          "args": types.SimpleNamespace(condmask_pixel=condmask_pixel, condmask_depth=condmask_depth)}
    exec(compile(ast.Module(body=keep, type_ignores=[]), "test_mvs4.py", "exec"), ns)
    return ns


def filter_case():
    h, w, v = 48, 64, 5
    k = syn.intrinsics(h, w, 3)
    es = [syn.extrinsics(i, 0.05, tilt_rad=0.01) for i in range(v)]
    depths = syn.render_surface_depths(k, es, h, w, noise_mm=0.6, seed=1)
    depths[1, 5:9, 7:12] = 0.0                       # holes in a source depth map
    depths[0, 20:23, 30:33] = 0.0                    # and in the reference (division by zero -> nan/inf)
    depths[2, 40:44, 10:20] += 25.0                  # an inconsistent patch
    conf = np.random.RandomState(0).uniform(0, 1, size=(v, h, w)).astype(np.float32)
    pairs = np.concatenate([np.arange(v)[:, None], syn.pair_list(v, 3)], 1).astype(np.int32)
    thr = dict(condmask_pixel=1.0, condmask_depth=0.01, photomask=0.75, geomask=2)
    ns = lift_filter_functions(thr["condmask_pixel"], thr["condmask_depth"])
    ks = np.stack([k] * v)
    es = np.stack(es)
    masks, dreps, xs, ys = [], [], [], []
    photo, geo, final, avg, gsum = [], [], [], [], []
    with np.errstate(all="ignore"):
        for row in pairs:
            r = int(row[0])
            acc, cnt, per = [], 0, []
            for s in row[1:]:
                m, d, x2, y2 = ns["check_geometric_consistency"](depths[r], ks[r], es[r], depths[int(s)], ks[int(s)],
                                                                 es[int(s)])
                per.append((m, d.copy(), x2, y2))
                cnt = cnt + m.astype(np.int32)
                acc.append(d)
            # mask fusion exactly as test_mvs4.py:716,744,746,749
            photo_mask = conf[r] > thr["photomask"]
            depth_est_averaged = (sum(acc) + depths[r]) / (cnt + 1)
            geo_mask = cnt >= thr["geomask"]
            final_mask = np.logical_and(photo_mask, geo_mask)
            masks.append(np.stack([p[0] for p in per]))
            dreps.append(np.stack([p[1] for p in per]))
            xs.append(np.stack([p[2] for p in per]))
            ys.append(np.stack([p[3] for p in per]))
            photo.append(photo_mask)
            geo.append(geo_mask)
            final.append(final_mask)
            avg.append(depth_est_averaged)
            gsum.append(cnt)
    np.savez_compressed(os.path.join(HERE, "filter.npz"), depths=depths, conf=conf, ks=ks, es=es, pairs=pairs,
                        pair_mask=np.stack(masks), pair_depth_reprojected=np.stack(dreps), pair_x2d_src=np.stack(xs),
                        pair_y2d_src=np.stack(ys), photo=np.stack(photo), geo=np.stack(geo), final=np.stack(final),
                        depth_avg=np.stack(avg), geo_sum=np.stack(gsum), **thr)
    print("filter: geo %.3f final %.3f" % (np.stack(geo).mean(), np.stack(final).mean()))


def cascade_case():
    """One full reference MVS4net forward (eval) with every stagenet call's inputs/outputs recorded."""
    h0, w0, n = 64, 128, 3
    torch.manual_seed(0)
    model = MVS4net(arch_mode="fpn", reg_net="reg2d", num_stage=4, fpn_base_channel=8, reg_channel=8,
                    stage_splits=[8, 8, 4, 4], depth_interals_ratio=[0.5, 0.5, 0.5, 1.0], group_cor=True,
                    group_cor_dim=[8, 8, 4, 4], inverse_depth=True, agg_type="ConvBnReLU3D", attn_temp=2.0,
                    attn_fuse_d=True).eval()
    rec = {}
    inner = model.stagenet

    class Tap(torch.nn.Module):
        def forward(self, features, proj_matrices, depth_hypo, regnet, stage_idx, **kw):
            store = {}

            class R(torch.nn.Module):
                def forward(s, x):
                    store["volume"] = x
                    store["logits"] = regnet(x)
                    return store["logits"]

            ret = inner(features, proj_matrices, depth_hypo, R(), stage_idx, **kw)
            k = "s%d_" % (stage_idx + 1)
            rec[k + "features"] = torch.stack(features, 1).numpy()
            rec[k + "proj"] = proj_matrices.numpy()
            rec[k + "hypo"] = depth_hypo.numpy()
            rec[k + "volume"] = store["volume"].numpy()
            rec[k + "logits"] = store["logits"].numpy()
            rec[k + "groups"] = kw["group_cor_dim"]
            rec[k + "split_itv"] = kw["split_itv"]
            for name, val in ret.items():
                rec[k + "out_" + name] = val.numpy()
            return ret

    model.stagenet = Tap()
    g = torch.Generator().manual_seed(11)
    imgs = [torch.rand(1, 3, h0, w0, generator=g) for _ in range(n)]
    proj = {k: torch.from_numpy(v) for k, v in syn.proj_matrices_all_stages(1, n, h0, w0).items()}
    dv = torch.from_numpy(syn.depth_values(1))
    with torch.no_grad():
        model(imgs, proj, dv)
    np.savez_compressed(os.path.join(HERE, "cascade.npz"), depth_values=dv.numpy(), **rec)
    print("cascade: recorded", len(rec), "arrays; stage-4 depth range",
          float(rec["s4_out_depth"].min()), float(rec["s4_out_depth"].max()))


def main():
    with torch.no_grad():
        pass
    k1_case("k1_stage1", b=1, n=3, c=64, g=8, d=8, h=8, w=10, stage=0, h0=64, w0=80)
    k1_case("k1_stage2", b=2, n=4, c=32, g=8, d=8, h=12, w=16, stage=1, h0=48, w0=64, jitter=0.4, seed=1)
    k1_case("k1_stage3", b=1, n=5, c=16, g=4, d=4, h=16, w=20, stage=2, h0=32, w0=40, tilt=0.02, seed=2)
    k1_case("k1_stage4", b=1, n=5, c=8, g=4, d=4, h=24, w=33, stage=3, h0=24, w0=33, seed=3)
    k1_case("k1_oob", b=1, n=3, c=8, g=4, d=4, h=10, w=12, stage=3, h0=10, w0=12, hs=7, ws=15, step=0.5,
            attn_temp=1.0, seed=4)
    warp_case()
    tail_case()
    schedule_case()
    filter_case()
    cascade_case()


if __name__ == "__main__":
    main()
