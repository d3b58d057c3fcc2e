"""Full-size (BASELINE.json shapes) checks of the CUDA path: windows of the 864x1152 problem against the float64
oracle, and size-independent properties (view-permutation invariance, batch independence, partition of unity of the
attention, idempotence of the filter under identical cameras).  Run on a B200 with ``-m gpu``."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import deep_reconstruction_with_epipolar_lines_mvster_b200 as mv
from deep_reconstruction_with_epipolar_lines_mvster_b200 import ops, synthetic as syn
from oracle import mvster_oracle as O

DEV = "cuda"
H0, W0, N = 864, 1152, 5


def _stage(stage, batch=1, seed=0):
    c, g, d = syn.STAGE_CHANNELS[stage], syn.STAGE_GROUPS[stage], syn.STAGE_NDEPTHS[stage]
    h, w = syn.stage_shape(H0, W0, stage)
    feats = [syn.smooth_features(batch, c, h, w, 31 * stage + v + seed) for v in range(N)]
    proj = syn.proj_matrices(batch, N, H0, W0, stage, per_batch_jitter=0.1)
    if stage == 0:
        hypo = O.init_inverse_range_np(syn.depth_values(batch), d, h, w)
    else:
        inv = 1.0 / np.stack([syn.smooth_depth_map(h // 2, w // 2, s + seed, 560, 800) for s in range(batch)])
        half = np.float32(0.5 * (1 / 425.0 - 1 / 935.0) / 7 / 4.0 ** (stage - 1))
        hypo = O.schedule_inverse_range_np((inv + half).astype(np.float32), (inv - half).astype(np.float32), d, h, w)
    return feats, proj, hypo, g, d, h, w


@pytest.mark.parametrize("stage", [0, 1, 2, 3])
def test_k1_full_size_windows_match_oracle(stage):
    feats, proj, hypo, g, d, h, w = _stage(stage)
    vol, wts = mv.epipolar_weights([f.to(DEV) for f in feats], torch.from_numpy(proj).to(DEV),
                                   torch.from_numpy(hypo).to(DEV), g, 2.0)
    vol, wts = vol.cpu().numpy(), wts.cpu().numpy()
    assert np.isfinite(vol).all()
    # partition of unity: per view the weights sum to 1/sqrt(C) over D
    c = feats[0].shape[1]
    assert np.abs(wts.sum(2) - 1.0 / np.sqrt(c)).max() < 1e-5
    wh, ww = min(12, h), min(20, w)
    for (y0, x0) in [(0, 0), (h - wh, w - ww), (h // 2 - wh // 2, w // 3), (h - wh, 0)]:
        win = (y0, y0 + wh, x0, x0 + ww)
        ref64, w64, _ = O.epipolar_aggregate_np(feats[0].numpy(), [f.numpy() for f in feats[1:]], proj, hypo, g, 2.0,
                                                window=win)
        assert np.abs(vol[:, :, :, y0:y0 + wh, x0:x0 + ww] - ref64).max() < 1e-4, (stage, win)
        assert np.abs(wts[:, :, :, y0:y0 + wh, x0:x0 + ww] - w64).max() < 1e-4, (stage, win)


@pytest.mark.parametrize("stage", [1, 3])
def test_k1_view_permutation_and_batch_independence(stage):
    feats, proj, hypo, g, d, h, w = _stage(stage, batch=2, seed=5)
    dfe = [f.to(DEV) for f in feats]
    dproj, dhyp = torch.from_numpy(proj).to(DEV), torch.from_numpy(hypo).to(DEV)
    vol = mv.epipolar_aggregate(dfe, dproj, dhyp, g, 2.0)
    # source views are summed: any order gives the same volume up to fp32 summation order
    perm = [0, 3, 1, 4, 2]
    vol_p = mv.epipolar_aggregate([dfe[i] for i in perm], dproj[:, perm].contiguous(), dhyp, g, 2.0)
    assert (vol - vol_p).abs().max().item() < 2e-6
    # batch items are independent: running item 1 alone is bit-identical
    one = mv.epipolar_aggregate([f[1:2].contiguous() for f in dfe], dproj[1:2].contiguous(), dhyp[1:2].contiguous(), g, 2.0)
    assert torch.equal(one[0], vol[1])


@pytest.mark.parametrize("c,g,d", [(8, 4, 4), (16, 4, 4), (32, 8, 8), (32, 16, 4)])
def test_k1_large_footprint_falls_back_to_direct_gather(c, g, d):
    """A 35 % scale change between the views makes every tile's footprint exceed the TMA box: the kernel must take
    its per-view direct-gather path and still match the oracle (swizzled-box kernels C=8/16 and the line kernel C=32)."""
    h, w = 64, 96
    feats = [syn.smooth_features(1, c, h, w, 77 + v) for v in range(3)]
    proj = syn.proj_matrices(1, 3, h, w, 3)
    proj[:, 1, 1, :2, :2] *= 1.35            # zoomed source camera
    proj[:, 2, 1, :2, :2] *= 0.7
    hypo = O.init_inverse_range_np(syn.depth_values(1), d, h, w)
    vol = mv.epipolar_aggregate([f.to(DEV) for f in feats], torch.from_numpy(proj).to(DEV),
                                torch.from_numpy(hypo).to(DEV), g, 2.0).cpu().numpy()
    ref64, _, _ = O.epipolar_aggregate_np(feats[0].numpy(), [f.numpy() for f in feats[1:]], proj, hypo, g, 2.0)
    assert np.abs(vol - ref64).max() < 1e-4


def test_tail_and_schedule_full_size_vs_torch():
    b, d, h, w = 2, 4, H0, W0
    g = torch.Generator(device=DEV).manual_seed(0)
    logits = torch.randn((b, d, h, w), device=DEV, generator=g) * 3
    inv = torch.rand((b, h // 2, w // 2), device=DEV, generator=g) * 5e-4 + 1.2e-3
    hypo = mv.schedule_inverse_range(inv + 3e-6, inv - 3e-6, d, h, w)
    itv = torch.arange(d, device=DEV, dtype=torch.float32).reshape(1, -1, 1, 1) / (d - 1)
    ref = (inv - 3e-6)[:, None] + ((inv + 3e-6) - (inv - 3e-6))[:, None] * itv
    ref = 1.0 / torch.nn.functional.interpolate(ref.unsqueeze(1), [d, h, w], mode="trilinear", align_corners=True).squeeze(1)
    assert ((hypo - ref).abs() / ref).max().item() < 1e-6
    attn, depth, conf, lo, hi = ops.tail(logits, hypo, 1.0, True, True)
    sm = torch.softmax(logits, 1)
    assert (attn - sm).abs().max().item() < 1e-6
    assert (attn.sum(1) - 1).abs().max().item() < 1e-5
    idx = sm.max(1, keepdim=True)[1]
    same = depth == torch.gather(hypo, 1, idx).squeeze(1)
    assert same.float().mean().item() > 0.9999          # fp32 ties in the softmax only
    mxl = logits.max(1)[0]
    ok = logits.sum(1).abs() > 1e-2
    assert torch.allclose(conf[ok], (mxl / logits.sum(1))[ok], rtol=1e-4, atol=1e-6)
    assert torch.allclose(lo - hi, 2.0 * (1.0 / hypo[:, 2] - 1.0 / hypo[:, 1]), rtol=1e-3, atol=1e-10)


def test_filter_512x640_properties():
    h, w, v = 512, 640, 6
    k = syn.intrinsics(h, w, 3)
    es = [syn.grid_extrinsics(i, 3, 0.05) for i in range(v)]
    depths = syn.render_surface_depths(k, es, h, w, noise_mm=0.2, seed=3)
    conf = np.random.RandomState(1).uniform(0, 1, size=(v, h, w)).astype(np.float32)
    ks, es = np.stack([k] * v), np.stack(es)
    pairs = np.concatenate([np.arange(v)[:, None], syn.pair_list(v, 4)], 1).astype(np.int32)
    cfg = mv.FilterConfig()
    photo, geo, final, avg, gsum = mv.filter_scene(depths, conf, ks, es, pairs, cfg, want_geo_sum=True)
    assert torch.equal(final, photo & geo)
    assert torch.equal(photo.cpu(), torch.from_numpy(conf > cfg.photomask))
    assert geo.float().mean().item() > 0.9               # consistent synthetic surface
    # source order does not matter for the vote count
    pairs2 = pairs.copy()
    pairs2[:, 1:] = pairs[:, :0:-1]
    _, _, _, avg2, gsum2 = mv.filter_scene(depths, conf, ks, es, pairs2, cfg, want_geo_sum=True)
    assert torch.equal(gsum, gsum2)
    assert (avg - avg2).abs().max().item() < 1e-3
    # one pair of the fused result against the per-pair entry point and the CPU oracle (1/32-px remap emulation)
    r, s = int(pairs[2, 0]), int(pairs[2, 1])
    m, drep, _, _ = mv.check_geometric_consistency(depths[r], ks[r], es[r], depths[s], ks[s], es[s], cfg)
    mo, do, _, _ = O.check_geometric_consistency_np(depths[r], ks[r], es[r], depths[s], ks[s], es[s], 1.0, 0.01)
    assert (m == mo).mean() >= 0.9999
    both = m & mo
    assert np.abs(drep[both] - do[both]).max() < 2e-3


def test_k1_randomised_shapes_vs_oracle():
    """Seeded sweep over ragged shapes (odd sizes, Hs != H, B > 1, every channel/group/depth combination the
    library compiles) against the float64 oracle."""
    rng = np.random.RandomState(123)
    combos = [(8, 4, 4), (8, 8, 8), (8, 1, 4), (16, 4, 4), (16, 2, 8), (32, 8, 8), (32, 4, 4), (32, 16, 8), (32, 32, 4),
              (64, 8, 8), (64, 16, 4)]
    for i, (c, g, d) in enumerate(combos):
        b, n = int(rng.randint(1, 3)), int(rng.randint(2, 6))
        h, w = int(rng.randint(5, 41)), int(rng.randint(7, 70))
        hs, ws = (h, w) if i % 3 else (h + int(rng.randint(-3, 4)), w + int(rng.randint(-4, 5)))
        feats = [syn.smooth_features(b, c, h, w, 1000 + i)]
        feats += [syn.smooth_features(b, c, hs, ws, 2000 + 10 * i + v) for v in range(1, n)]
        proj = syn.proj_matrices(b, n, h * 4, w * 4, 1, step_rad=float(rng.uniform(0.02, 0.2)),
                                 per_batch_jitter=0.3, tilt_rad=float(rng.uniform(0, 0.05)))
        lo = rng.uniform(430, 700)
        hypo = np.sort(rng.uniform(lo, lo + 200, size=(b, d, h, w)).astype(np.float32), 1)[:, ::-1].copy()
        vol, wts = mv.epipolar_weights([f.to(DEV) for f in feats], torch.from_numpy(proj).to(DEV),
                                       torch.from_numpy(hypo).to(DEV), g, 1.7)
        ref64, w64, _ = O.epipolar_aggregate_np(feats[0].numpy(), [f.numpy() for f in feats[1:]], proj, hypo, g, 1.7)
        assert np.abs(vol.cpu().numpy() - ref64).max() < 1e-4, (c, g, d, b, n, h, w, hs, ws)
        assert np.abs(wts.cpu().numpy() - w64).max() < 1e-4, (c, g, d, b, n, h, w, hs, ws)


def test_k1_line_kernel_equals_direct_kernel(monkeypatch):
    """C=32 (whole-line texels): the TMA-staged rotated-chunk kernel and the direct-gather kernel (MVSTER_NO_LINE=1)
    compute the same sums in a different order - they must agree to fp32 rounding at the full stage-2 size."""
    feats, proj, hypo, g, d, h, w = _stage(1, batch=2, seed=3)
    dfe = [f.to(DEV) for f in feats]
    dproj, dhyp = torch.from_numpy(proj).to(DEV), torch.from_numpy(hypo).to(DEV)
    vol, wts = mv.epipolar_weights(dfe, dproj, dhyp, g, 2.0)
    monkeypatch.setenv("MVSTER_NO_LINE", "1")
    vol_d, wts_d = mv.epipolar_weights(dfe, dproj, dhyp, g, 2.0)
    assert (vol - vol_d).abs().max().item() < 5e-6
    assert (wts - wts_d).abs().max().item() < 5e-6


def test_cascade_free_running_vs_cpu_port():
    """The whole 4-stage hot path (CascadePlan: schedule -> K1 -> stand-in regnet -> tail, stage after stage, nothing
    teacher-forced) against the oracle's CPU port of the same cascade, at 128x192, N=4."""
    from deep_reconstruction_with_epipolar_lines_mvster_b200.pipeline import CascadePlan
    h0, w0, n = 128, 192, 4
    plan = CascadePlan(1, n, h0, w0, device=DEV)
    feats, projs, gts = [], [], []
    for s in range(4):
        h, w = syn.stage_shape(h0, w0, s)
        fs = [syn.smooth_features(1, syn.STAGE_CHANNELS[s], h, w, 50 * s + v) for v in range(n)]
        feats.append(fs)
        projs.append(torch.from_numpy(syn.proj_matrices(1, n, h0, w0, s)))
        gts.append(torch.from_numpy((1.0 / syn.smooth_depth_map(h, w, 9, 560, 800))[None].astype(np.float32)))
        for v in range(n):
            plan.features[s][v].copy_(fs[v].permute(0, 2, 3, 1))
        plan.proj[s].copy_(projs[s])
    dv = torch.from_numpy(syn.depth_values(1))
    plan.depth_values.copy_(dv)

    def logits_of(hypo, gt):   # deterministic stand-in for regnet: peaked at the hypothesis nearest the surface
        inv = 1.0 / hypo
        itv = (inv[:, 1:2] - inv[:, 0:1]).abs().clamp_min(1e-12)
        return -4.0 * (inv - gt[:, None].to(hypo.device)).abs() / itv

    plan.regnet = lambda s, vol: logits_of(plan.hypo[s], gts[s]).contiguous()
    depth, conf = plan.run()
    want_d, want_c = O.cascade_port(feats, projs, dv, [lambda hy, g=g: logits_of(hy, g) for g in gts],
                                    syn.STAGE_GROUPS, syn.STAGE_NDEPTHS, syn.STAGE_SPLIT_ITV, 2.0)
    depth, conf = depth.cpu(), conf.cpu()
    # an arg-max flip anywhere in the cascade moves a pixel to a neighbouring hypothesis: allow a small fraction
    close = (depth - want_d).abs() < 1e-3 * 2.5
    assert close.float().mean().item() > 0.995, close.float().mean().item()
    ok = close & torch.isfinite(want_c) & (want_c.abs() < 1e3)
    assert torch.allclose(conf[ok], want_c[ok], rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("stage", [2, 3])
def test_k1_bf16_tma_path_full_size_equals_direct_path(stage, monkeypatch):
    """bf16 features at the fine stages (16/32-byte texels) go through the TMA-staged swizzled box as well; the
    direct-gather kernel (MVSTER_NO_TMA=1) computes the same sums in the same order per pixel."""
    feats, proj, hypo, g, d, h, w = _stage(stage, batch=1, seed=11)
    dfe = [f.to(DEV).bfloat16() for f in feats]
    dproj, dhyp = torch.from_numpy(proj).to(DEV), torch.from_numpy(hypo).to(DEV)
    vol = mv.epipolar_aggregate(dfe, dproj, dhyp, g, 2.0)
    monkeypatch.setenv("MVSTER_NO_TMA", "1")
    vol_d = mv.epipolar_aggregate(dfe, dproj, dhyp, g, 2.0)
    assert (vol - vol_d).abs().max().item() < 5e-6
    ref64, _, _ = O.epipolar_aggregate_np(feats[0].bfloat16().float().numpy(), [f.bfloat16().float().numpy() for f in feats[1:]],
                                          proj, hypo, g, 2.0, window=(0, 16, 0, 24))
    assert np.abs(vol[:, :, :, :16, :24].cpu().numpy() - ref64).max() < 1e-4


def test_run_from_host_overlapped_copies_equal_resident_run():
    """CascadePlan.run_from_host (pinned host buffers in, stage k+1 copied on a side stream while stage k computes,
    depth + confidence out) gives exactly what run() gives on resident inputs - also on the second call, when the
    device buffers still hold the previous step's data."""
    from deep_reconstruction_with_epipolar_lines_mvster_b200.pipeline import CascadePlan
    h0, w0, n = 128, 192, 3
    plan = CascadePlan(2, n, h0, w0, device=DEV)
    plan.make_host_buffers()
    gen = torch.Generator().manual_seed(4)
    results = []
    for step in range(2):
        for s in range(4):
            for hf in plan.h_features[s]:
                hf.copy_(torch.randn(hf.shape, generator=gen) * 0.5)
            plan.h_proj[s].copy_(torch.from_numpy(syn.proj_matrices(2, n, h0, w0, s, per_batch_jitter=0.05 * (step + 1))))
            plan.logits[s].copy_(torch.randn(plan.logits[s].shape, generator=gen).to(DEV))
        plan.h_depth_values.copy_(torch.from_numpy(syn.depth_values(2)))
        hd, hc = plan.run_from_host()
        hd, hc = hd.clone(), hc.clone()
        depth, conf = plan.run()     # the inputs are resident now: same launches, same stream
        torch.cuda.synchronize()
        assert torch.equal(hd, depth.cpu()) and torch.equal(hc, conf.cpu()), step
        results.append(hd)
    assert not torch.equal(results[0], results[1])


def _grads(feats, proj, hypo, g, gout):
    fs = [f.detach().clone().requires_grad_(True) for f in feats]
    vol = mv.epipolar_aggregate(fs, proj, hypo, g, 2.0)
    vol.backward(gout)
    return [f.grad for f in fs]


@pytest.mark.parametrize("stage", [2, 3])
def test_k1_tma_backward_equals_direct_backward_full_size(stage, monkeypatch):
    """Fine stages (C=16 / C=8, fp32): the TMA-staged backward kernel and the direct-gather backward
    (MVSTER_NO_TMA_BWD=1) compute the same gradients; grad_ref is summed in registers in the same order (tight bound),
    grad_src goes through floating-point atomics in both (bound = atomic reordering noise)."""
    feats, proj, hypo, g, d, h, w = _stage(stage, batch=2, seed=5)
    dfe = [f.to(DEV) for f in feats]
    dproj, dhyp = torch.from_numpy(proj).to(DEV), torch.from_numpy(hypo).to(DEV)
    gout = torch.randn((2, g, d, h, w), device=DEV, generator=torch.Generator(device=DEV).manual_seed(stage))
    tma = _grads(dfe, dproj, dhyp, g, gout)
    monkeypatch.setenv("MVSTER_NO_TMA_BWD", "1")
    direct = _grads(dfe, dproj, dhyp, g, gout)
    scale = max(1.0, float(direct[0].abs().max()))
    assert (tma[0] - direct[0]).abs().max().item() < 2e-5 * scale
    for a, b in zip(tma[1:], direct[1:]):
        assert (a - b).abs().max().item() < 1e-4 * max(1.0, float(b.abs().max()))
        assert a.abs().sum().item() > 0


@pytest.mark.parametrize("c,g,d,h,w", [(8, 4, 4, 64, 96), (16, 4, 4, 37, 53), (8, 8, 4, 5, 33), (16, 16, 4, 40, 31),
                                       (8, 1, 4, 9, 70)])
def test_k1_backward_fallback_and_ragged_shapes_match_autograd_of_the_port(c, g, d, h, w):
    """Zoomed source cameras push every tile's footprint over the TMA box (per-view direct gather inside the staged
    backward kernel) and the ragged sizes exercise the dead-lane handling; the check is the autograd of the oracle's
    op-for-op port on the same inputs."""
    feats = [syn.smooth_features(1, c, h, w, 91 + v) for v in range(4)]
    proj = syn.proj_matrices(1, 4, h, w, 3)
    proj[:, 1, 1, :2, :2] *= 1.35            # footprint larger than the box
    proj[:, 2, 1, :2, :2] *= 0.7
    hypo = O.init_inverse_range_np(syn.depth_values(1), d, h, w)
    gen = torch.Generator().manual_seed(c + h)
    gout = torch.randn((1, g, d, h, w), generator=gen)
    cpu = [f.clone().requires_grad_(True) for f in feats]
    O.epipolar_aggregate_port(cpu, torch.from_numpy(proj), torch.from_numpy(hypo), g, 2.0).backward(gout)
    got = _grads([f.to(DEV) for f in feats], torch.from_numpy(proj).to(DEV), torch.from_numpy(hypo).to(DEV), g,
                 gout.to(DEV))
    for a, b in zip(got, cpu):
        assert (a.cpu() - b.grad).abs().max().item() < 2e-4 * max(1.0, float(b.grad.abs().max()))


# ---------------------------------------------------------------------------------------------------------------------
# parity holes named by the round-1 review: other sizes / view counts, full-size backward against the reference op
# sequence (not only TMA-vs-direct), full-size filter against the lifted reference
# ---------------------------------------------------------------------------------------------------------------------
def _stage_at(h0, w0, n, stage, batch=1, seed=0):
    c, g, d = syn.STAGE_CHANNELS[stage], syn.STAGE_GROUPS[stage], syn.STAGE_NDEPTHS[stage]
    h, w = syn.stage_shape(h0, w0, stage)
    feats = [syn.smooth_features(batch, c, h, w, 17 * stage + v + seed) for v in range(n)]
    proj = syn.proj_matrices(batch, n, h0, w0, stage, per_batch_jitter=0.1)
    if stage == 0:
        hypo = O.init_inverse_range_np(syn.depth_values(batch), d, h, w)
    else:
        inv = 1.0 / np.stack([syn.smooth_depth_map(h // 2, w // 2, s + seed, 560, 800) for s in range(batch)])
        half = np.float32(0.5 * (1 / 425.0 - 1 / 935.0) / 7 / 4.0 ** (stage - 1))
        hypo = O.schedule_inverse_range_np((inv + half).astype(np.float32), (inv - half).astype(np.float32), d, h, w)
    return feats, proj, hypo, g, d, h, w


@pytest.mark.parametrize("h0,w0,n", [(832, 1152, 5), (832, 1152, 2), (512, 640, 4), (512, 640, 2)])
@pytest.mark.parametrize("stage", [0, 1, 2, 3])
def test_k1_windows_at_other_sizes_and_view_counts(h0, w0, n, stage):
    """the four stage shapes at {512x640, 832x1152} with N in {2, 4, 5} (864x1152 N=5 is covered above):
    volume and per-view attention weights of windows against the float64 oracle, 1e-4."""
    feats, proj, hypo, g, d, h, w = _stage_at(h0, w0, n, stage, seed=h0 + n)
    vol, wts = mv.epipolar_weights([f.to(DEV) for f in feats], torch.from_numpy(proj).to(DEV),
                                   torch.from_numpy(hypo).to(DEV), g, 2.0)
    vol, wts = vol.cpu().numpy(), wts.cpu().numpy()
    assert np.isfinite(vol).all()
    wh, ww = min(10, h), min(16, w)
    for (y0, x0) in [(0, w - ww), (h // 2, w // 2 - ww // 2), (h - wh, 3)]:
        ref64, w64, _ = O.epipolar_aggregate_np(feats[0].numpy(), [f.numpy() for f in feats[1:]], proj, hypo, g, 2.0,
                                                window=(y0, y0 + wh, x0, x0 + ww))
        assert np.abs(vol[:, :, :, y0:y0 + wh, x0:x0 + ww] - ref64).max() < 1e-4, (stage, y0, x0)
        assert np.abs(wts[:, :, :, y0:y0 + wh, x0:x0 + ww] - w64).max() < 1e-4, (stage, y0, x0)


@pytest.mark.parametrize("stage", [2, 3])
def test_k1_backward_full_training_size_matches_autograd_of_the_port(stage):
    """Config 4's shape (512x640, B=2, N=5), fine stages: gradients w.r.t. the reference and every source feature map
    against torch autograd of the oracle's op-for-op port of the reference sequence (homo_warping -> grid_sample ->
    group correlation -> softmax -> aggregation) on the host - the CUDA backward against the reference's math, not
    against another CUDA kernel.  2e-4 x max(1, |g|) as for the small goldens."""
    feats, proj, hypo, g, d, h, w = _stage_at(512, 640, 5, stage, batch=2, seed=3)
    gen = torch.Generator().manual_seed(100 + stage)
    gout = torch.randn((2, g, d, h, w), generator=gen)
    cpu = [f.clone().requires_grad_(True) for f in feats]
    O.epipolar_aggregate_port(cpu, torch.from_numpy(proj), torch.from_numpy(hypo), g, 2.0).backward(gout)
    got = _grads([f.to(DEV) for f in feats], torch.from_numpy(proj).to(DEV), torch.from_numpy(hypo).to(DEV), g,
                 gout.to(DEV))
    for i, (a, b) in enumerate(zip(got, cpu)):
        scale = max(1.0, float(b.grad.abs().max()))
        assert (a.cpu() - b.grad).abs().max().item() < 2e-4 * scale, (stage, i)
        assert float(b.grad.abs().sum()) > 0


def _unpack(bits, w):
    return np.unpackbits(bits, axis=-1)[..., :w].astype(bool)


def test_filter_512x640_matches_the_lifted_reference():
    """Three 512x640 depth maps, every pair checked by the UNMODIFIED reference functions (tests/golden/
    make_golden_filter512.py): masks identical on >= 99.99 % of the pixels, reprojected / averaged depth within 2e-3
    where both agree."""
    import hashlib, os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "filter512.npz"))
    depths, conf, ks, es, pairs = syn.filter_fixture_512()
    sha = hashlib.sha256()
    for a in (depths, conf, ks, es, pairs):
        sha.update(np.ascontiguousarray(a).tobytes())
    assert sha.hexdigest() == str(g["inputs_sha256"]), "synthetic.filter_fixture_512() no longer produces the golden's inputs"
    h, w = depths.shape[1:]
    cfg = mv.FilterConfig(float(g["condmask_pixel"]), float(g["condmask_depth"]), float(g["photomask"]), int(g["geomask"]))
    want_m = _unpack(g["pair_mask_bits"], w)
    total = same = 0
    for i, row in enumerate(pairs):
        r = int(row[0])
        for j, s in enumerate(row[1:]):
            s = int(s)
            m, d, _, _ = mv.check_geometric_consistency(depths[r], ks[r], es[r], depths[s], ks[s], es[s], cfg)
            total += m.size
            same += int((m == want_m[i, j]).sum())
            both = (m & want_m[i, j])[::4, ::4]
            assert np.abs(d[::4, ::4][both] - g["pair_depth_reprojected_s4"][i, j][both]).max() < 2e-3
    assert same / total >= 0.9999, same / total
    photo, geo, final, avg, _ = mv.filter_scene(depths, conf, ks, es, pairs, cfg, want_geo_sum=True)
    assert np.array_equal(photo.cpu().numpy(), _unpack(g["photo_bits"], w))
    assert (geo.cpu().numpy() == _unpack(g["geo_bits"], w)).mean() >= 0.9999
    assert (final.cpu().numpy() == _unpack(g["final_bits"], w)).mean() >= 0.9999
    a4, w4 = avg.cpu().numpy()[:, ::4, ::4], g["depth_avg_s4"]
    ok = np.isfinite(w4) & (np.abs(a4 - w4) < 1.0)          # a vote that flipped changes the average by whole millimetres
    assert ok.mean() > 0.999
    assert np.abs(a4[ok] - w4[ok]).max() < 2e-3


@pytest.mark.parametrize("n,c,g,d", [(8, 8, 4, 4), (16, 8, 4, 4), (7, 16, 4, 4), (9, 8, 2, 8)])
def test_k1_many_source_views_recycle_the_staging_buffers(n, c, g, d):
    """More source views than staging buffers (3): every further view is requested after a CTA barrier into a recycled
    buffer, with the mbarrier phase flipping - up to the ABI's maximum of 15 source views."""
    h, w = 40, 72
    feats = [syn.smooth_features(1, c, h, w, 300 + v) for v in range(n)]
    proj = syn.proj_matrices(1, n, h * 8, w * 8, 3, step_rad=0.01)
    inv = 1.0 / syn.smooth_depth_map(h // 2, w // 2, 4, 600, 760)[None]
    half = np.float32(2e-6)
    hypo = O.schedule_inverse_range_np((inv + half).astype(np.float32), (inv - half).astype(np.float32), d, h, w)
    vol, wts = mv.epipolar_weights([f.to(DEV) for f in feats], torch.from_numpy(proj).to(DEV),
                                   torch.from_numpy(hypo).to(DEV), g, 2.0)
    ref64, w64, _ = O.epipolar_aggregate_np(feats[0].numpy(), [f.numpy() for f in feats[1:]], proj, hypo, g, 2.0)
    assert np.abs(vol.cpu().numpy() - ref64).max() < 1e-4
    assert np.abs(wts.cpu().numpy() - w64).max() < 1e-4
