#!/usr/bin/env python
"""Secondary benchmarks (BASELINE.json configs 2, 4, 5 and per-stage K1 timings) - one JSON line each.

    python scripts/bench_extra.py [--which stages,binpick,train,filter,ref_gpu] [--iters 20]

  stages   per-stage K1 forward at the DTU 864x1152 N=5 shapes (B=8): time, algorithmic GB/s, roofline fraction
  binpick  config 2: N=4 views 512x640, fused K1 forward, fp32 and bf16 features (B=8)
  train    config 4 (single GPU part): K1 forward+backward at 512x640, B=2, N=5 through EpipolarAggregate
  filter   config 5: 49 views 512x640, 9 sources each, fused geometric/photometric filter
  ref_gpu  plain-PyTorch (eager, stock ATen kernels) restatement of the same op on the same GPU, as context
  config0  BASELINE configs[0]: whole MVS4net forward 512x640 N=5 on the host CPU (reference op sequence) vs the B200 path
  scene    49-view scene end to end on the GPU: depth maps -> filter -> point cloud
  network  whole MVS4net.forward (FPN4 + reg2d via cuDNN, fused stagenet / regulariser tail) vs the eager op sequence
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import deep_reconstruction_with_epipolar_lines_mvster_b200 as mv  # noqa: E402
from deep_reconstruction_with_epipolar_lines_mvster_b200 import ops, synthetic as syn  # noqa: E402

HBM = 6547.2
if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")):
    HBM = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def smooth_hypo(b, d, h, w, stage, dev):
    if stage == 0:
        return mv.init_inverse_range(torch.from_numpy(syn.depth_values(b)).to(dev), d, None, None, h, w)
    inv = torch.from_numpy(np.stack([1.0 / syn.smooth_depth_map(h // 2, w // 2, s, 560, 800) for s in range(b)])).to(dev)
    half = 0.5 * (1 / 425.0 - 1 / 935.0) / 7 / (4.0 ** (stage - 1)) / (2 if stage > 1 else 1)
    return mv.schedule_inverse_range(inv + half, inv - half, d, h, w)


def stage_inputs(b, n, h0, w0, stage, dev, dtype=torch.float32):
    c, g, d = syn.STAGE_CHANNELS[stage], syn.STAGE_GROUPS[stage], syn.STAGE_NDEPTHS[stage]
    h, w = syn.stage_shape(h0, w0, stage)
    gen = torch.Generator(device=dev).manual_seed(stage)
    feats = []
    for v in range(n):
        x = torch.randn((b, c, h, w), device=dev, generator=gen) * 0.5
        x = torch.nn.functional.avg_pool2d(x, 3, 1, 1, count_include_pad=False) * 1.7
        feats.append(x.to(dtype).contiguous(memory_format=torch.channels_last))
    proj = torch.from_numpy(syn.proj_matrices(b, n, h0, w0, stage, per_batch_jitter=0.02)).to(dev)
    return feats, proj, smooth_hypo(b, d, h, w, stage, dev), g, d, c, h, w


def k1_bytes(b, n, c, g, d, h, w, s):
    return b * h * w * (n * c * s + d * 4 + g * d * 4)


def bench_stages(args, dev):
    for stage in range(4):
        feats, proj, hypo, g, d, c, h, w = stage_inputs(8, 5, 864, 1152, stage, dev)
        nhwc = [ops.to_nhwc(f) for f in feats]
        rt = ops.compose_homographies(proj)
        ms = timed(lambda: ops.epi_fwd(nhwc[0], nhwc[1:], rt, hypo, g, 2.0), args.iters)
        by = k1_bytes(8, 5, c, g, d, h, w, 4)
        print(json.dumps({"bench": "k1_fwd_stage%d" % (stage + 1), "shape": [8, 5, c, g, d, h, w], "ms": ms,
                          "algorithmic_GB": by / 1e9, "GBps": by / ms / 1e6, "frac_hbm_roofline": by / ms / 1e6 / HBM}))


def bench_binpick(args, dev):
    for dtype, name in ((torch.float32, "fp32"), (torch.bfloat16, "bf16")):
        tot_ms, tot_by = 0.0, 0
        for stage in range(4):
            feats, proj, hypo, g, d, c, h, w = stage_inputs(8, 4, 512, 640, stage, dev, dtype)
            nhwc = [ops.to_nhwc(f) for f in feats]
            rt = ops.compose_homographies(proj)
            tot_ms += timed(lambda: ops.epi_fwd(nhwc[0], nhwc[1:], rt, hypo, g, 2.0), args.iters)
            tot_by += k1_bytes(8, 4, c, g, d, h, w, 2 if dtype == torch.bfloat16 else 4)
        print(json.dumps({"bench": "binpick_k1_4stages_" + name, "config": "N=4 512x640 B=8", "ms_per_8_depth_maps": tot_ms,
                          "depth_maps_per_s": 8e3 / tot_ms, "GBps": tot_by / tot_ms / 1e6,
                          "frac_hbm_roofline": tot_by / tot_ms / 1e6 / HBM}))


def bench_train(args, dev):
    out = {}
    for stage in (3, 2, 1, 0):
        feats, proj, hypo, g, d, c, h, w = stage_inputs(2, 5, 512, 640, stage, dev)
        feats = [f.requires_grad_(True) for f in feats]
        gout = torch.randn((2, g, d, h, w), device=dev)

        def step():
            for f in feats:
                f.grad = None
            vol = mv.epipolar_aggregate(feats, proj, hypo, g, 2.0)
            vol.backward(gout)

        def fwd_only():
            with torch.no_grad():
                mv.epipolar_aggregate(feats, proj, hypo, g, 2.0)

        out["stage%d_fwd_bwd_ms" % (stage + 1)] = timed(step, args.iters)
        out["stage%d_fwd_ms" % (stage + 1)] = timed(fwd_only, args.iters)
        # the backward launch alone (C-ABI call + the zero-fill of grad_src it scatters into), and its HBM roofline:
        # SURVEY 8d bytes = grad_out + ref + src + hypo read, grad_ref + grad_src written (+ out, wsum re-read here)
        with torch.no_grad():
            nh = [ops.to_nhwc(f.detach()) for f in feats]
            rt = ops.compose_homographies(proj)
            vol, wsum, _ = ops.epi_fwd(nh[0], nh[1:], rt, hypo, g, 2.0, want_wsum=True)
            ms = timed(lambda: ops.epi_bwd(nh[0], nh[1:], rt, hypo, vol, wsum, gout, g, 2.0), args.iters)
            # the same call replayed from a CUDA graph (10 calls per replay): device time without the Python / ctypes
            # / allocator cost of an eager call, which exceeds the kernel time at the coarse stages
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                ops.epi_bwd(nh[0], nh[1:], rt, hypo, vol, wsum, gout, g, 2.0)
            torch.cuda.current_stream().wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for _ in range(10):
                    keep = ops.epi_bwd(nh[0], nh[1:], rt, hypo, vol, wsum, gout, g, 2.0)
            gms = timed(graph.replay, max(3, args.iters // 4)) / 10
            del graph, keep
        px = 2 * h * w
        alg = px * (4 * g * d * 2 + 4 * d * 2 + 5 * c * 4 * 2)
        out["stage%d_bwd_call_ms" % (stage + 1)] = ms
        out["stage%d_bwd_frac_hbm_roofline" % (stage + 1)] = alg / (HBM * 1e6) / ms
        out["stage%d_bwd_graph_ms" % (stage + 1)] = gms
        out["stage%d_bwd_graph_frac_hbm_roofline" % (stage + 1)] = alg / (HBM * 1e6) / gms
    out["bench"] = "train_k1_fwd_bwd"
    out["config"] = "512x640 B=2 N=5 fp32, EpipolarAggregate autograd (includes layout views, zero-init of grads)"
    print(json.dumps(out))


def bench_filter(args, dev):
    h, w, v, s = 512, 640, 49, 9
    k = syn.intrinsics(h, w, 3)
    es = [syn.grid_extrinsics(i, 7, 0.04) for i in range(v)]
    # smooth analytic depths: rendering 49 views on the CPU is slow, use one rendered map per row of the rig
    depths = syn.render_surface_depths(k, es, h, w, noise_mm=0.3, seed=0)
    conf = np.random.RandomState(0).uniform(0, 1, size=(v, h, w)).astype(np.float32)
    pairs = np.concatenate([np.arange(v)[:, None], syn.pair_list(v, s)], 1).astype(np.int32)
    ks = np.stack([k] * v)
    es = np.stack(es)
    dz, cf = torch.from_numpy(depths).to(dev), torch.from_numpy(conf).to(dev)
    cfg = mv.FilterConfig()
    ms = timed(lambda: mv.filter_scene(dz, cf, ks, es, pairs, cfg), max(3, args.iters // 4))
    photo, geo, final, avg, _ = mv.filter_scene(dz, cf, ks, es, pairs, cfg)
    by = 441 * h * w * 4 + 49 * (2 * h * w * 4 + h * w * 4 + 3 * h * w)
    line = {"bench": "filter_49x9_512x640", "ms_per_scene": ms, "scenes_per_s": 1e3 / ms, "pair_checks_per_s": 441e3 / ms,
            "algorithmic_GB": by / 1e9, "GBps": by / ms / 1e6, "geo_mask_mean": float(geo.float().mean()),
            "final_mask_mean": float(final.float().mean()),
            "sha1_masks": hashlib.sha1(torch.stack([photo, geo, final]).cpu().numpy().tobytes()).hexdigest()[:16],
            "sha1_depth_avg": hashlib.sha1(avg.cpu().numpy().tobytes()).hexdigest()[:16],
            "ppt": os.environ.get("MVSTER_FILTER_PPT", "4")}
    if args.cpu_filter_pairs > 0:
        from oracle import mvster_oracle as O
        t0 = time.perf_counter()
        n = 0
        for r in range(v):
            for j in range(s):
                if n >= args.cpu_filter_pairs:
                    break
                src = int(pairs[r, 1 + j])
                O.check_geometric_consistency_np(depths[r], ks[r], es[r], depths[src], ks[src], es[src], 1.0, 0.01, use_cv2=True)
                n += 1
        dt = time.perf_counter() - t0
        line["cpu_port_ms_per_pair"] = 1e3 * dt / n
        line["cpu_port_s_per_scene_extrapolated"] = dt / n * 441
        line["cpu_pairs_timed"] = n
    print(json.dumps(line))


def bench_ref_gpu(args, dev):
    """The reference's op sequence in eager PyTorch on the same GPU (stock ATen kernels), via the oracle port."""
    from oracle import mvster_oracle as O
    for stage in range(4):
        feats, proj, hypo, g, d, c, h, w = stage_inputs(2, 5, 864, 1152, stage, dev)
        feats = [f.contiguous() for f in feats]
        with torch.no_grad():
            ms = timed(lambda: O.epipolar_aggregate_port(feats, proj, hypo, g, 2.0), max(3, args.iters // 4))
            nhwc = [ops.to_nhwc(f) for f in feats]
            rt = ops.compose_homographies(proj)
            ours = timed(lambda: ops.epi_fwd(nhwc[0], nhwc[1:], rt, hypo, g, 2.0), args.iters)
        print(json.dumps({"bench": "eager_pytorch_vs_fused_stage%d" % (stage + 1), "shape": [2, 5, c, g, d, h, w],
                          "eager_ms": ms, "fused_ms": ours, "speedup": ms / ours}))


def bench_ref_gpu_train(args, dev):
    """Eager-PyTorch forward+backward of the reference op sequence vs the fused op, config-4 shape (512x640, B=2)."""
    from oracle import mvster_oracle as O
    for stage in (3, 2):
        feats, proj, hypo, g, d, c, h, w = stage_inputs(2, 5, 512, 640, stage, dev)
        gout = torch.randn((2, g, d, h, w), device=dev)
        fe = [f.contiguous().requires_grad_(True) for f in feats]
        ff = [f.requires_grad_(True) for f in feats]

        def eager():
            for f in fe:
                f.grad = None
            O.epipolar_aggregate_port(fe, proj, hypo, g, 2.0).backward(gout)

        def fused():
            for f in ff:
                f.grad = None
            mv.epipolar_aggregate(ff, proj, hypo, g, 2.0).backward(gout)

        e_ms, f_ms = timed(eager, max(3, args.iters // 4)), timed(fused, args.iters)
        torch.cuda.reset_peak_memory_stats()
        eager(); torch.cuda.synchronize(); e_mem = torch.cuda.max_memory_allocated()
        torch.cuda.reset_peak_memory_stats()
        fused(); torch.cuda.synchronize(); f_mem = torch.cuda.max_memory_allocated()
        print(json.dumps({"bench": "eager_pytorch_vs_fused_fwd_bwd_stage%d" % (stage + 1), "shape": [2, 5, c, g, d, h, w],
                          "eager_ms": e_ms, "fused_ms": f_ms, "speedup": e_ms / f_ms,
                          "eager_peak_MB": e_mem / 1e6, "fused_peak_MB": f_mem / 1e6}))


NET_CFG = dict(arch_mode="fpn", reg_net="reg2d", num_stage=4, fpn_base_channel=8, reg_channel=8,
               stage_splits=[8, 8, 4, 4], depth_interals_ratio=[0.5, 0.5, 0.5, 1.0], group_cor=True,
               group_cor_dim=[8, 8, 4, 4], inverse_depth=True, agg_type="ConvBnReLU3D", attn_temp=2.0, attn_fuse_d=True)


class _EagerStagenet(torch.nn.Module):
    """The reference's stagenet op sequence in eager PyTorch (stock ATen kernels) - SURVEY §8d row (iii) on the GPU
    box, where /root/reference itself is not available: the oracle's op-for-op port + the torch tail."""

    def forward(self, features, proj_matrices, depth_hypo, regnet, stage_idx, group_cor=True, group_cor_dim=8,
                split_itv=1, fn=None):
        from oracle import mvster_oracle as O
        vol = O.epipolar_aggregate_port([f.contiguous() for f in features], proj_matrices, depth_hypo, group_cor_dim, 2.0)
        logits = regnet(vol)
        conf = logits.max(1)[0] / logits.sum(1)
        attn = torch.softmax(logits, 1)
        idx = attn.argmax(1, keepdim=True)
        depth = torch.gather(depth_hypo, 1, idx).squeeze(1)
        itv = 1.0 / depth_hypo[:, 2] - 1.0 / depth_hypo[:, 1]
        return {"depth": depth, "photometric_confidence": conf, "hypo_depth": depth_hypo, "attn_weight": attn,
                "inverse_min_depth": 1.0 / depth + split_itv * itv, "inverse_max_depth": 1.0 / depth - split_itv * itv}


def bench_network(args, dev):
    """whole MVS4net.forward, images in -> 4-stage depth out, DTU as-loaded shape
    832x1152 (reg2d needs multiples of 64), N=5, one scene per call."""
    h0, w0, n, b = 832, 1152, 5, 1
    model = mv.MVS4net(**NET_CFG).eval()
    model.load_state_dict(syn.fill_state_dict(model.state_dict(), seed=7))
    model = model.to(dev)
    gen = torch.Generator(device=dev).manual_seed(0)
    imgs = [torch.rand((b, 3, h0, w0), device=dev, generator=gen) for _ in range(n)]
    proj = {k: torch.from_numpy(v).to(dev) for k, v in syn.proj_matrices_all_stages(b, n, h0, w0).items()}
    dv = torch.from_numpy(syn.depth_values(b)).to(dev)
    fused_stagenet = model.stagenet
    res = {}
    its = max(3, args.iters // 4)

    def run():
        with torch.no_grad():
            return model(imgs, proj, dv)

    for name, tf32, fuse, eager in (("b200_fused_fp32", False, True, False), ("b200_unfused_fp32", False, False, False),
                                    ("b200_fused_tf32", True, True, False), ("eager_reference_like_fp32", False, False, True),
                                    ("eager_reference_like_tf32", True, False, True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        model.fuse_regnet_tail = fuse
        model.stagenet = _EagerStagenet() if eager else fused_stagenet
        if eager:   # the reference extracts features view by view in NCHW
            model.extract_features = lambda ims: [model.feature(i) for i in ims]
        elif "extract_features" in model.__dict__:
            del model.__dict__["extract_features"]
        torch.cuda.reset_peak_memory_stats()
        res[name + "_ms"] = timed(run, its)
        res[name + "_peak_MB"] = torch.cuda.max_memory_allocated() / 1e6
    # the same forward as one CUDA-graph launch (fp32 and TF32-cuDNN), B = 1 and B = 2 scenes per call
    model.stagenet, model.fuse_regnet_tail = fused_stagenet, True
    model.__dict__.pop("extract_features", None)
    for tf32 in (False, True):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        for bb in (1, 2):
            ims = [torch.rand((bb, 3, h0, w0), device=dev, generator=gen) for _ in range(n)]
            pj = {k: torch.from_numpy(v).to(dev) for k, v in syn.proj_matrices_all_stages(bb, n, h0, w0).items()}
            dvs = torch.from_numpy(syn.depth_values(bb)).to(dev)
            gm = mv.GraphedMVS4net(model, bb, n, h0, w0, dev).capture(ims, pj, dvs)
            ms = timed(lambda: gm(ims, pj, dvs), its)
            res["graph_%s_b%d_ms_per_scene" % ("tf32" if tf32 else "fp32", bb)] = ms / bb
            del gm
    # where the time goes on the B200 path (fused, fp32)
    torch.backends.cudnn.allow_tf32 = False
    model.stagenet, model.fuse_regnet_tail = fused_stagenet, True
    model.__dict__.pop("extract_features", None)
    with torch.no_grad():
        res["fpn_ms"] = timed(lambda: model.extract_features(imgs), its)
        vol = torch.randn((b, 4, 4, h0, w0), device=dev)
        hyp = torch.rand((b, 4, h0, w0), device=dev) * 400 + 450
        res["reg2d_stage4_unfused_plus_tail_ms"] = timed(lambda: ops.tail(model.reg[3](vol), hyp, 1.0, True, True), its)
        res["reg2d_stage4_fused_tail_ms"] = timed(lambda: model.reg[3].forward_fused_tail(vol, hyp, 1.0), its)
        conv0, low = model.reg[3]._trunk(vol)
        w, params = model.reg[3]._fold()
        ms = timed(lambda: ops.regtail(low, conv0, w, params, hyp, 1.0, True), args.iters)
        res["regtail_kernel_ms"] = ms
        res["regtail_GFMA_per_s"] = b * 4 * (h0 // 2) * (w0 // 2) * 1152 / ms / 1e6
        res["regtail_algorithmic_GBps"] = b * 4 * h0 * w0 * (16 / 4 + 8 + 1 + 1 + 1) * 4 / ms / 1e6
    res.update({"bench": "mvs4net_forward_832x1152_n5", "depth_maps_per_s_b200_fused_fp32": 1e3 * b / res["b200_fused_fp32_ms"],
                "depth_maps_per_s_eager_fp32": 1e3 * b / res["eager_reference_like_fp32_ms"],
                "speedup_vs_eager_fp32": res["eager_reference_like_fp32_ms"] / res["b200_fused_fp32_ms"]})
    print(json.dumps(res))


def bench_config0(args, dev):
    """BASELINE.json configs[0]: MVS4Net 4-stage forward, DTU-train shape B=1 N=5 512x640, ndepths 8-8-4-4, fp32 -
    the reference's op sequence on the host CPU cores (all threads) next to the B200 path on the same inputs."""
    h0, w0, n, b = 512, 640, 5, 1
    model = mv.MVS4net(**NET_CFG).eval()
    model.load_state_dict(syn.fill_state_dict(model.state_dict(), seed=7))
    gen = torch.Generator().manual_seed(0)
    imgs = [torch.rand((b, 3, h0, w0), generator=gen) for _ in range(n)]
    proj = {k: torch.from_numpy(v) for k, v in syn.proj_matrices_all_stages(b, n, h0, w0).items()}
    dv = torch.from_numpy(syn.depth_values(b))
    # CPU: per-view NCHW FPN, eager warp / correlation / attention, reg2d, torch tail - the reference's op sequence
    # (the schedule comes from the oracle: this package has no CPU path)
    from oracle import mvster_oracle as O
    eager = _EagerStagenet()

    def cpu_forward():
        feats = [model.feature(i) for i in imgs]
        outs, st = {}, None
        for si in range(4):
            key = "stage%d" % (si + 1)
            fs = [f[key] for f in feats]
            hh, ww = fs[0].shape[2:]
            d = NET_CFG["stage_splits"][si]
            if si == 0:
                hypo = torch.from_numpy(O.init_inverse_range_np(dv.numpy(), d, hh, ww))
            else:
                hypo = torch.from_numpy(O.schedule_inverse_range_np(st["inverse_min_depth"].numpy(),
                                                                    st["inverse_max_depth"].numpy(), d, hh, ww))
            st = eager(fs, proj[key], hypo, model.reg[si], si, group_cor_dim=NET_CFG["group_cor_dim"][si],
                       split_itv=NET_CFG["depth_interals_ratio"][si])
            outs[key] = st
        return outs

    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        cpu_forward()
        t0 = time.perf_counter()
        reps = 2
        for _ in range(reps):
            out_cpu = cpu_forward()
        cpu_s = (time.perf_counter() - t0) / reps
    model = model.to(dev)
    torch.backends.cudnn.allow_tf32 = False
    dimgs = [i.to(dev) for i in imgs]
    dproj = {k: v.to(dev) for k, v in proj.items()}
    ddv = dv.to(dev)

    def run():
        with torch.no_grad():
            return model(dimgs, dproj, ddv)

    ms = timed(run, max(5, args.iters // 2))
    out = run()
    a, r = out["stage1"]["attn_weight"].cpu(), out_cpu["stage1"]["attn_weight"]
    agree = (out["stage4"]["depth"].cpu() - out_cpu["stage4"]["depth"]).abs() < 1e-3 * 2.5
    print(json.dumps({"bench": "config0_mvs4net_forward_512x640_n5_fp32", "cpu_reference_like_s": cpu_s,
                      "cpu_threads": torch.get_num_threads(), "cpu_depth_maps_per_s": 1.0 / cpu_s, "b200_ms": ms,
                      "b200_depth_maps_per_s": 1e3 / ms, "speedup": cpu_s * 1e3 / ms,
                      "stage1_attn_max_abs_diff": float((a - r).abs().max()),
                      "stage4_depth_agree_frac": float(agree.float().mean())}))


def bench_scene(args, dev):
    """Whole-scene reconstruction, GPU-resident (BASELINE configs[4] extended to the full test_mvs4.py pipeline):
    49 views 512x640 -> 49 depth + confidence maps (MVS4net, N=5 views each, one CUDA-graph replay per reference view)
    -> photometric/geometric filter over 49x9 pairs -> averaged depth -> world points of the kept pixels.  Nothing
    leaves the GPU between the stages (the reference writes and re-reads 98 .pfm files in between)."""
    h0, w0, v, nv, s = 512, 640, 49, 5, 9
    model = mv.MVS4net(**NET_CFG).eval()
    model.load_state_dict(syn.fill_state_dict(model.state_dict(), seed=7))
    model = model.to(dev)
    torch.backends.cudnn.allow_tf32 = False
    gen = torch.Generator(device=dev).manual_seed(0)
    images = torch.rand((v, 3, h0, w0), device=dev, generator=gen)
    k = syn.intrinsics(h0, w0, 3)
    es = np.stack([syn.grid_extrinsics(i, 7, 0.04) for i in range(v)])
    ks = np.stack([k] * v)
    nbrs = syn.pair_list(v, s)                                     # [V, 9] source views, nearest first
    pairs = np.concatenate([np.arange(v)[:, None], nbrs], 1).astype(np.int32)
    dv = torch.from_numpy(syn.depth_values(1)).to(dev)

    def projections(ref):
        views = [ref] + [int(x) for x in nbrs[ref][:nv - 1]]
        out = {}
        for st in range(4):
            p = np.zeros((1, nv, 2, 4, 4), np.float32)
            kk = syn.intrinsics(h0, w0, st)
            for j, vi in enumerate(views):
                p[0, j, 0] = es[vi]
                p[0, j, 1, :3, :3] = kk
            out["stage%d" % (st + 1)] = torch.from_numpy(p).to(dev)
        return views, out

    pre = [projections(r) for r in range(v)]
    gm = mv.GraphedMVS4net(model, 1, nv, h0, w0, dev)
    depths = torch.empty((v, h0, w0), device=dev)
    confs = torch.empty((v, h0, w0), device=dev)
    cfg = mv.FilterConfig(photomask=0.0)   # random-init weights: keep every pixel the geometry accepts

    def run():
        for r in range(v):
            views, proj = pre[r]
            out = gm([images[i:i + 1] for i in views], proj, dv)["stage4"]
            depths[r].copy_(out["depth"][0])
            confs[r].copy_(out["photometric_confidence"][0])
        return mv.fuse_scene(depths, confs, ks, es, pairs, images=images.permute(0, 2, 3, 1), config=cfg)

    run()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        verts, cols, info = run()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(json.dumps({"bench": "scene_49views_512x640_depth_filter_points", "seconds_per_scene": dt,
                      "depth_maps_per_s": v / dt, "points_kept": int(verts.shape[0]),
                      "final_mask_mean": float(info["final"].float().mean()),
                      "note": "random-init network: depth maps are not geometrically consistent, so few points survive; "
                              "the timing does not depend on it"}))


def bench_sinkhorn(args, dev):
    """K3: the four per-stage OT losses of one training step (512x640, B=2, ndepths 8-8-4-4, train_mvs4.py default
    ot_iter=10, eps=1): fused loss+gradient kernel vs the reference's op sequence in eager PyTorch on the same GPU
    (oracle port: [B,HW,D,D] tensors, 2*iters logsumexp passes, autograd backward)."""
    from deep_reconstruction_with_epipolar_lines_mvster_b200 import loss as L
    from oracle import mvster_oracle as O
    gen = torch.Generator(device=dev).manual_seed(11)
    stages = []
    for d, h, w in ((8, 64, 80), (8, 128, 160), (4, 256, 320), (4, 512, 640)):
        hypo = torch.sort(500 + 300 * torch.rand(2, d, h, w, device=dev, generator=gen), dim=1, descending=True)[0]
        gt = 450 + 400 * torch.rand(2, h, w, device=dev, generator=gen)
        attn = torch.softmax(torch.randn(2, d, h, w, device=dev, generator=gen), 1).requires_grad_(True)
        mask = torch.rand(2, h, w, device=dev, generator=gen) > 0.2
        stages.append((gt, hypo, attn, mask))
    out = {"bench": "sinkhorn_loss_4stages_fwd_bwd", "config": "512x640 B=2 ndepths 8-8-4-4 ot_iter=10 eps=1 fp32"}

    def fused():
        for gt, hypo, attn, mask in stages:
            attn.grad = None
            L.SinkhornLoss.apply(gt, hypo, attn, mask, 10, 1.0, False, True)[0][0].backward()

    def eager():
        for gt, hypo, attn, mask in stages:
            attn.grad = None
            O.sinkhorn_port(gt, hypo, attn, mask, 10, 1.0, False)[1].backward()

    out["fused_ms"] = timed(fused, args.iters)
    torch.cuda.reset_peak_memory_stats()
    fused()
    out["fused_peak_MB"] = torch.cuda.max_memory_allocated() / 1e6
    out["eager_ms"] = timed(eager, max(3, args.iters // 4))
    torch.cuda.reset_peak_memory_stats()
    eager()
    out["eager_peak_MB"] = torch.cuda.max_memory_allocated() / 1e6
    out["speedup"] = out["eager_ms"] / out["fused_ms"]
    gt, hypo, attn, mask = stages[3]
    ms4 = timed(lambda: ops.sinkhorn_fwd(gt, hypo, attn.detach(), mask, 10, 1.0, False, True), args.iters)
    px = gt.numel()
    # per pixel and iteration: 2*D*D exp + 2*D log forward, 2*D*D exp backward (D = 4), plus D*D for the transport map
    mufu = px * (10 * (4 * 4 * 4 + 2 * 4) + 4 * 4 + 4)
    out["stage4_kernel_ms"] = ms4
    out["stage4_Gtranscendental_per_s"] = mufu / ms4 / 1e6
    out["stage4_frac_of_xu_peak"] = (mufu / ms4 / 1e6) / (148 * 16 * 1.92)   # 16 MUFU/clk/SM at 1.92 GHz
    print(json.dumps(out))


def bench_train_step(args, dev):
    """BASELINE configs[3] widened to the whole network: one training step (MVS4net.train() forward, MVS4net_loss with
    ot_iter=10, backward, Adam) at the DTU-train shape 512x640, B=2, N=5, fp32 - the B200 path (fused K1 fwd/bwd, tail
    fwd/bwd, fused Sinkhorn loss; cuDNN convolutions) vs the reference's op sequence in eager PyTorch on the same GPU."""
    from deep_reconstruction_with_epipolar_lines_mvster_b200 import loss as L
    from oracle import mvster_oracle as O
    h0, w0, n, b = 512, 640, 5, 2
    torch.backends.cudnn.allow_tf32 = False
    model = mv.MVS4net(**NET_CFG).train()
    model.load_state_dict(syn.fill_state_dict(model.state_dict(), seed=7))
    model = model.to(dev)
    gen = torch.Generator(device=dev).manual_seed(0)
    imgs = [torch.rand((b, 3, h0, w0), device=dev, generator=gen) for _ in range(n)]
    proj = {k: torch.from_numpy(v).to(dev) for k, v in syn.proj_matrices_all_stages(b, n, h0, w0).items()}
    dv = torch.from_numpy(syn.depth_values(b)).to(dev)
    gts, masks = {}, {}
    for s in range(4):
        h, w = h0 >> (3 - s), w0 >> (3 - s)
        gts["stage%d" % (s + 1)] = 560 + 300 * torch.rand((b, h, w), device=dev, generator=gen)
        masks["stage%d" % (s + 1)] = (torch.rand((b, h, w), device=dev, generator=gen) > 0.2).float()
    kw = dict(stage_lw=[1, 1, 1, 1], l1ot_lw=[0, 1], inverse_depth=True, ot_iter=10, ot_eps=1, ot_continous=False)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    fused_stagenet = model.stagenet

    def eager_loss(out):
        total = 0.0
        for k in out:
            total = total + O.sinkhorn_port(gts[k], out[k]["hypo_depth"], out[k]["attn_weight"], masks[k] > 0.5, 10, 1.0)[1]
        return total

    def step(eager):
        opt.zero_grad(set_to_none=True)
        out = model(imgs, proj, dv)
        total = eager_loss(out) if eager else L.MVS4net_loss(out, gts, masks, **kw)[0]
        total.backward()
        opt.step()

    res = {"bench": "train_step_mvs4net_512x640_b2_n5", "config": "fwd + OT loss (10 iters) + bwd + Adam, fp32 (TF32 off)"}
    its = max(3, args.iters // 4)
    from deep_reconstruction_with_epipolar_lines_mvster_b200 import network as NW
    # hand-written training kernels of the regulariser on / off (fused BatchNorm + ReLU: mvster_bn_train_*; weight
    # gradient of the 3-D convolutions: mvster_conv3d_wgrad), cuDNN TF32 off / on, cudnn.benchmark as train_mvs4.py sets it
    torch.backends.cudnn.benchmark = True
    for name, eager, fused_bn, hand_wgrad, tf32 in (
            ("b200", False, True, True, False), ("b200_cudnn_wgrad", False, True, False, False),
            ("b200_cudnn_wgrad_batchnorm", False, False, False, False), ("eager_reference_like", True, False, False, False),
            ("b200_tf32", False, True, True, True), ("b200_cudnn_wgrad_tf32", False, True, False, True),
            ("b200_cudnn_wgrad_batchnorm_tf32", False, False, False, True)):
        model.stagenet = _EagerStagenet() if eager else fused_stagenet
        NW.FUSED_TRAIN_BATCHNORM = fused_bn
        NW.HAND_WGRAD3D = NW.HAND_WGRAD2D = hand_wgrad
        torch.backends.cudnn.allow_tf32 = tf32
        torch.cuda.reset_peak_memory_stats()
        res[name + "_ms"] = timed(lambda: step(eager), its)
        res[name + "_peak_MB"] = torch.cuda.max_memory_allocated() / 1e6
    NW.HAND_WGRAD3D = NW.HAND_WGRAD2D = True
    model.stagenet = fused_stagenet
    NW.FUSED_TRAIN_BATCHNORM = True
    torch.backends.cudnn.allow_tf32 = False
    res["samples_per_s_b200"] = 1e3 * b / res["b200_ms"]
    res["speedup"] = res["eager_reference_like_ms"] / res["b200_ms"]
    print(json.dumps(res))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="stages,binpick,train,filter,ref_gpu,ref_gpu_train,network,config0,scene,sinkhorn,train_step")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--cpu-filter-pairs", type=int, default=20)
    ap.add_argument("--cudnn-benchmark", action="store_true", help="let cuDNN auto-tune its algorithm per shape")
    args = ap.parse_args()
    torch.backends.cudnn.benchmark = bool(args.cudnn_benchmark)
    dev = torch.device("cuda", 0)
    fns = {"stages": bench_stages, "binpick": bench_binpick, "train": bench_train, "filter": bench_filter,
           "ref_gpu": bench_ref_gpu, "ref_gpu_train": bench_ref_gpu_train, "network": bench_network, "config0": bench_config0, "scene": bench_scene, "sinkhorn": bench_sinkhorn, "train_step": bench_train_step}
    for name in args.which.split(","):
        fns[name](args, dev)


if __name__ == "__main__":
    main()
