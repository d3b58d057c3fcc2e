"""GPU parity tests: the CUDA path (through the C ABI of libmvster_b200.so) against the golden vectors produced by
the unmodified reference and against the CPU oracle on seeded inputs.  Run on a B200 with ``-m gpu``.

Tolerances (BASELINE.json north_star): attention weights within 1e-4 and volume within 1e-4 for fp32 features
(measured: ~1e-5, the reference's own fp32 noise floor); depth within 1e-3 of the depth interval wherever the
arg-max agrees; filter masks identical on >= 99.99 % of pixels.  bf16 features: 2e-2 against the fp32 reference
(stated looser bound: bf16 keeps 8 mantissa bits) and 1e-4 against exact math on the same bf16-rounded features.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import deep_reconstruction_with_epipolar_lines_mvster_b200 as mv
from deep_reconstruction_with_epipolar_lines_mvster_b200 import ops, synthetic as syn
from oracle import mvster_oracle as O

DEV = "cuda"
K1_CASES = ["k1_stage1", "k1_stage2", "k1_stage3", "k1_stage4", "k1_oob"]


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _features(g, channels_last=False, requires_grad=False):
    feats = [_cuda(g["ref"])] + [_cuda(g["srcs"][:, v]) for v in range(g["srcs"].shape[1])]
    if channels_last:
        feats = [f.contiguous(memory_format=torch.channels_last) for f in feats]
    if requires_grad:
        feats = [f.requires_grad_(True) for f in feats]
    return feats


# --------------------------------------------------------------------------------------------------------------
# K1 forward / backward against the reference's golden vectors
# --------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("channels_last", [False, True])
@pytest.mark.parametrize("name", K1_CASES)
def test_k1_forward_matches_reference(golden, name, channels_last):
    g = golden(name)
    feats = _features(g, channels_last)
    vol, wts = mv.epipolar_weights(feats, _cuda(g["proj"]), _cuda(g["hypo"]), int(g["groups"]), float(g["attn_temp"]))
    assert vol.shape == g["volume"].shape and vol.dtype == torch.float32
    assert np.abs(vol.cpu().numpy() - g["volume"]).max() < 1e-4
    assert np.abs(wts.cpu().numpy() - g["weights"]).max() < 1e-4
    vol2 = mv.epipolar_aggregate(feats, _cuda(g["proj"]), _cuda(g["hypo"]), int(g["groups"]), float(g["attn_temp"]))
    assert torch.equal(vol, vol2)


@pytest.mark.parametrize("name", K1_CASES)
def test_k1_backward_matches_reference_autograd(golden, name):
    g = golden(name)
    feats = _features(g, requires_grad=True)
    vol = mv.epipolar_aggregate(feats, _cuda(g["proj"]), _cuda(g["hypo"]), int(g["groups"]), float(g["attn_temp"]))
    (vol * _cuda(g["gout"])).sum().backward()
    gr = feats[0].grad.cpu().numpy()
    scale = max(1.0, np.abs(g["grad_ref"]).max())
    assert np.abs(gr - g["grad_ref"]).max() < 2e-4 * scale
    for v in range(g["srcs"].shape[1]):
        gs = feats[1 + v].grad.cpu().numpy()
        scale = max(1.0, np.abs(g["grad_srcs"][:, v]).max())
        assert np.abs(gs - g["grad_srcs"][:, v]).max() < 2e-4 * scale, (name, v)


def test_k1_backward_is_reproducible_within_atomic_tolerance(golden):
    g = golden("k1_stage4")
    grads = []
    for _ in range(2):
        feats = _features(g, requires_grad=True)
        vol = mv.epipolar_aggregate(feats, _cuda(g["proj"]), _cuda(g["hypo"]), int(g["groups"]), float(g["attn_temp"]))
        (vol * _cuda(g["gout"])).sum().backward()
        grads.append([f.grad.clone() for f in feats])
    assert torch.equal(grads[0][0], grads[1][0])          # grad_ref: no atomics, bit-exact
    for a, b in zip(grads[0][1:], grads[1][1:]):            # grad_src: fp32 atomics, order-dependent rounding only
        assert (a - b).abs().max().item() < 1e-5


@pytest.mark.parametrize("name", ["k1_stage1", "k1_stage2", "k1_stage3", "k1_stage4", "k1_oob"])
def test_k1_bf16_features(golden, name):
    g = golden(name)
    feats = _features(g)
    proj, hypo = _cuda(g["proj"]), _cuda(g["hypo"])
    vol = mv.epipolar_aggregate(feats, proj, hypo, int(g["groups"]), float(g["attn_temp"]),
                                feature_dtype=torch.bfloat16).cpu().numpy()
    # looser stated bound against the fp32 reference
    assert np.abs(vol - g["volume"]).max() < 2e-2
    # tight bound against exact math on the same bf16-rounded features
    rnd = lambda a: torch.from_numpy(a).bfloat16().float().numpy()
    srcs = [rnd(g["srcs"][:, v]) for v in range(g["srcs"].shape[1])]
    ref64, _, _ = O.epipolar_aggregate_np(rnd(g["ref"]), srcs, g["proj"], g["hypo"], int(g["groups"]),
                                          float(g["attn_temp"]))
    assert np.abs(vol - ref64).max() < 1e-4
    # bf16 tensors passed directly take the same path
    vol_b = mv.epipolar_aggregate([f.bfloat16() for f in feats], proj, hypo, int(g["groups"]),
                                  float(g["attn_temp"])).cpu().numpy()
    assert np.array_equal(vol, vol_b)


def test_k1_all_samples_out_of_frame_gives_zero_volume():
    b, c, g, d, h, w = 1, 8, 4, 4, 9, 11
    feats = [syn.smooth_features(b, c, h, w, s, device=DEV) for s in range(3)]
    proj = syn.proj_matrices(b, 3, h, w, 3)
    proj[:, 1:, 0, 0, 3] += 1e5      # push the source cameras far away: every sample leaves the image
    hypo = mv.init_inverse_range(_cuda(syn.depth_values(b)), d, None, None, h, w)
    vol = mv.epipolar_aggregate(feats, _cuda(proj), hypo, g, 2.0)
    assert torch.count_nonzero(vol).item() == 0


def test_k1_identity_homography_known_answer():
    """src camera == ref camera and src features == ref features: warped == ref exactly, every hypothesis scores the
    same, softmax is uniform, and the volume is mean_c(ref^2) per group (times S/(S+1e-8))."""
    b, c, g, d, h, w = 2, 16, 4, 4, 13, 37
    ref = syn.smooth_features(b, c, h, w, 3, device=DEV)
    proj = syn.proj_matrices(b, 3, h, w, 3)
    proj[:, 1:] = proj[:, :1]
    hypo = mv.init_inverse_range(_cuda(syn.depth_values(b)), d, None, None, h, w)
    vol = mv.epipolar_aggregate([ref, ref.clone(), ref.clone()], _cuda(proj), hypo, g, 2.0)
    expect = (ref.double() ** 2).reshape(b, g, c // g, h, w).mean(2)[:, :, None].expand(-1, -1, d, -1, -1)
    # the identity is exact only up to the fp32 rounding of R = P P^-1 (sub-1e-3 px), hence a small tolerance
    assert (vol.double() - expect).abs().max().item() < 2e-3
    assert (vol[:, :, 0] - vol[:, :, 1]).abs().max().item() < 2e-3


def test_k1_rejects_unsupported_configs(golden):
    g = golden("k1_stage4")
    feats = _features(g)
    proj, hypo = _cuda(g["proj"]), _cuda(g["hypo"])
    # the variance cost has one output channel per feature channel: group_cor=False with G != C is an error
    with pytest.raises(RuntimeError, match="G == C|must be"):
        ops.epi_bwd_mode(ops.to_nhwc(feats[0], torch.float32), [ops.to_nhwc(f, torch.float32) for f in feats[1:]],
                         ops.compose_homographies(proj), hypo, torch.zeros(1, 4, 4, 24, 33, device=DEV),
                         torch.ones(1, 4, 24, 33, device=DEV), torch.zeros(1, 4, 4, 24, 33, device=DEV), 4, 2.0, False, True)
    with pytest.raises(RuntimeError, match="not in"):
        bad = [torch.zeros(1, 24, 8, 8, device=DEV) for _ in range(2)]
        mv.epipolar_aggregate(bad, proj[:, :2], torch.ones(1, 4, 8, 8, device=DEV), 4, 2.0)
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        mv.epipolar_aggregate([f.cpu() for f in feats], proj.cpu(), hypo.cpu(), 4, 2.0)


@pytest.mark.parametrize("stage,h0,w0,n", [(0, 256, 320, 5), (1, 128, 160, 4), (2, 128, 160, 5), (3, 96, 120, 5)])
def test_k1_vs_oracle_seeded(stage, h0, w0, n):
    """Moderate sizes, every shipped (C, G, D) stage shape, against the float64 oracle and the fp32 port."""
    c, g, d = syn.STAGE_CHANNELS[stage], syn.STAGE_GROUPS[stage], syn.STAGE_NDEPTHS[stage]
    h, w = syn.stage_shape(h0, w0, stage)
    b = 2
    feats = [syn.smooth_features(b, c, h, w, 100 * stage + v) for v in range(n)]
    proj = syn.proj_matrices(b, n, h0, w0, stage, per_batch_jitter=0.2, tilt_rad=0.01)
    if stage == 0:
        hypo = O.init_inverse_range_np(syn.depth_values(b), d, h, w)
    else:
        inv = 1.0 / np.stack([syn.smooth_depth_map(h // 2, w // 2, s) for s in range(b)])
        half = np.float32(0.5 * (1 / 425.0 - 1 / 935.0) / 2 ** stage / 4)
        hypo = O.schedule_inverse_range_np((inv + half).astype(np.float32), (inv - half).astype(np.float32), d, h, w)
    vol, wts = mv.epipolar_weights([f.to(DEV) for f in feats], _cuda(proj), _cuda(hypo), g, 2.0)
    ref64, w64, _ = O.epipolar_aggregate_np(feats[0].numpy(), [f.numpy() for f in feats[1:]], proj, hypo, g, 2.0)
    assert np.abs(vol.cpu().numpy() - ref64).max() < 1e-4
    assert np.abs(wts.cpu().numpy() - w64).max() < 1e-4
    port = O.epipolar_aggregate_port(feats, torch.from_numpy(proj), torch.from_numpy(hypo), g, 2.0)
    assert np.abs(vol.cpu().numpy() - port.numpy()).max() < 1e-4


# --------------------------------------------------------------------------------------------------------------
# compatibility seams
# --------------------------------------------------------------------------------------------------------------
def test_homo_warping_matches_reference(golden):
    g = golden("warp")
    out = mv.homo_warping(_cuda(g["src"]), _cuda(g["src_proj"]), _cuda(g["ref_proj"]), _cuda(g["hypo"]))
    assert out.shape == g["warped"].shape
    assert np.abs(out.cpu().numpy() - g["warped"]).max() < 1e-4


def test_schedule_matches_reference(golden):
    g = golden("schedule")
    init = mv.init_inverse_range(_cuda(g["depth_values"]), 8, DEV, torch.float32, 6, 7).cpu().numpy()
    assert np.abs(init - g["init"]).max() / np.abs(g["init"]).max() < 3e-7
    for d in (4, 8):
        s = mv.schedule_inverse_range(_cuda(g["inv_min"]), _cuda(g["inv_max"]), d, 14, 18).cpu().numpy()
        assert s.shape == g["sched%d" % d].shape
        assert np.abs(s - g["sched%d" % d]).max() / np.abs(s).max() < 1e-6


class _Replay(torch.nn.Module):
    def __init__(self, logits):
        super().__init__()
        self.logits = logits
        self.seen = None

    def forward(self, x):
        self.seen = x
        return self.logits


def test_tail_matches_reference(golden):
    g = golden("tail")
    b, d, h, w = g["logits"].shape
    feats = [syn.smooth_features(b, 8, h, w, s, device=DEV) for s in range(2)]
    proj = _cuda(syn.proj_matrices(b, 2, h, w, 3))
    for mode in ("eval", "train"):
        net = mv.stagenet(inverse_depth=True, mono=True, attn_temp=2.0)
        net = net.eval() if mode == "eval" else net.train()
        ret = net(feats, proj, _cuda(g["hypo"]), _Replay(_cuda(g["logits"])), 1, group_cor=True, group_cor_dim=4,
                  split_itv=float(g["split_itv"]))
        assert list(ret.keys()) == ["depth", "photometric_confidence", "hypo_depth", "attn_weight",
                                    "inverse_min_depth", "inverse_max_depth", "mono_feat"]
        assert np.array_equal(ret["depth"].cpu().numpy(), g[mode + "_depth"])
        assert np.abs(ret["attn_weight"].cpu().numpy() - g[mode + "_attn_weight"]).max() < 1e-6
        for k in ("inverse_min_depth", "inverse_max_depth"):
            assert np.abs(ret[k].cpu().numpy() - g[mode + "_" + k]).max() < 1e-9
        conf = ret["photometric_confidence"].cpu().numpy()
        if mode == "train":
            assert conf.shape == () and conf == 0.0
        else:
            ok = np.isfinite(g["eval_photometric_confidence"])
            assert np.allclose(conf[ok], g["eval_photometric_confidence"][ok], rtol=2e-5, atol=1e-6)
        assert ret["mono_feat"] is feats[0]


def test_tail_generic_depth_count_and_regression():
    rng = np.random.RandomState(0)
    for d in (3, 5, 16, 48):
        logits = rng.normal(size=(2, d, 7, 9)).astype(np.float32) * 2
        hypo = np.sort(rng.uniform(400, 900, size=(2, d, 7, 9)).astype(np.float32), 1)[:, ::-1].copy()
        for regress in (False, True):
            want = O.tail_np(logits, hypo, 0.5, regress=regress)
            attn, depth, conf, lo, hi = ops.tail(_cuda(logits), _cuda(hypo), 0.5, True, True,
                                                 ops.DEPTH_REGRESS if regress else ops.DEPTH_ARGMAX)
            assert np.abs(attn.cpu().numpy() - want["attn_weight"]).max() < 1e-6
            if regress:
                assert np.abs(depth.cpu().numpy() - want["depth"]).max() < 1e-3
            else:
                assert np.array_equal(depth.cpu().numpy(), want["depth"])
            assert np.allclose(lo.cpu().numpy(), want["inverse_min_depth"], rtol=1e-5, atol=1e-9)


def test_tail_backward_matches_torch_softmax():
    torch.manual_seed(0)
    logits = torch.randn(2, 8, 6, 5, device=DEV, requires_grad=True)
    hypo = torch.rand(2, 8, 6, 5, device=DEV) * 500 + 400
    gw = torch.randn(2, 8, 6, 5, device=DEV)
    from deep_reconstruction_with_epipolar_lines_mvster_b200.stagenet import _Tail
    attn, depth, _, _, _ = _Tail.apply(logits, hypo, 0.5, False, True, ops.DEPTH_ARGMAX)
    (attn * gw).sum().backward()
    ref = logits.detach().clone().requires_grad_(True)
    (torch.softmax(ref, 1) * gw).sum().backward()
    assert (logits.grad - ref.grad).abs().max().item() < 1e-6
    # regression mode: depth = sum attn * hypo is differentiable too
    l2 = logits.detach().clone().requires_grad_(True)
    attn, depth, _, _, _ = _Tail.apply(l2, hypo, 0.5, False, True, ops.DEPTH_REGRESS)
    ((attn * gw).sum() + depth.sum() * 1e-2).backward()
    r2 = logits.detach().clone().requires_grad_(True)
    a = torch.softmax(r2, 1)
    ((a * gw).sum() + (a * hypo).sum() * 1e-2).backward()
    assert (l2.grad - r2.grad).abs().max().item() < 1e-4


def test_depth_regression(golden):
    g = golden("tail")
    p = torch.softmax(_cuda(g["logits"]), 1)
    got = mv.depth_regression(p, _cuda(g["hypo"]))
    want = (p * _cuda(g["hypo"])).sum(1)
    assert (got - want).abs().max().item() < 1e-2 * 1e-1   # ~1e-6 relative on depths of ~600


# --------------------------------------------------------------------------------------------------------------
# teacher-forced cascade through the drop-in stagenet
# --------------------------------------------------------------------------------------------------------------
def test_cascade_teacher_forced_matches_reference(golden):
    g = golden("cascade")
    net = mv.stagenet(inverse_depth=True, mono=False, attn_fuse_d=True, attn_temp=2.0).eval()
    for s in range(1, 5):
        k = "s%d_" % s
        feats = [_cuda(g[k + "features"][:, v]) for v in range(g[k + "features"].shape[1])]
        replay = _Replay(_cuda(g[k + "logits"]))
        with torch.no_grad():
            ret = net(feats, _cuda(g[k + "proj"]), _cuda(g[k + "hypo"]), replay, s - 1, group_cor=True,
                      group_cor_dim=int(g[k + "groups"]), split_itv=float(g[k + "split_itv"]))
        assert np.abs(replay.seen.cpu().numpy() - g[k + "volume"]).max() < 1e-4, "stage %d volume" % s
        assert np.abs(ret["attn_weight"].cpu().numpy() - g[k + "out_attn_weight"]).max() < 1e-4
        # teacher-forced logits -> identical arg-max -> identical depth, except where the reference's own top-2
        # softmax values are within fp32 rounding of each other (random-init regnet: nearly flat logits)
        srt = np.sort(g[k + "out_attn_weight"], 1)
        decided = (srt[:, -1] - srt[:, -2]) > 1e-6
        same = ret["depth"].cpu().numpy() == g[k + "out_depth"]
        assert same[decided].all(), "stage %d: arg-max differs on a decided pixel" % s
        assert same.mean() > 0.97, "stage %d flip fraction %.4f" % (s, 1 - same.mean())
        same_nb = same
        conf_ref = g[k + "out_photometric_confidence"]
        ok = np.isfinite(conf_ref) & (np.abs(conf_ref) < 1e3)
        assert np.allclose(ret["photometric_confidence"].cpu().numpy()[ok], conf_ref[ok], rtol=1e-4, atol=1e-5)
        for name in ("inverse_min_depth", "inverse_max_depth"):
            assert np.allclose(ret[name].cpu().numpy()[same_nb], g[k + "out_" + name][same_nb], rtol=1e-6, atol=1e-10)
        if s < 4:   # the schedule feeding the next stage (teacher-forced with the reference's inverse range)
            kn = "s%d_" % (s + 1)
            h, w = g[kn + "hypo"].shape[2:]
            nxt = mv.schedule_inverse_range(_cuda(g[k + "out_inverse_min_depth"]), _cuda(g[k + "out_inverse_max_depth"]),
                                            g[kn + "hypo"].shape[1], h, w)
            assert np.abs(nxt.cpu().numpy() - g[kn + "hypo"]).max() / np.abs(g[kn + "hypo"]).max() < 1e-6


# --------------------------------------------------------------------------------------------------------------
# K2b filter
# --------------------------------------------------------------------------------------------------------------
def test_filter_pairs_match_reference(golden):
    g = golden("filter")
    cfg = mv.FilterConfig(float(g["condmask_pixel"]), float(g["condmask_depth"]), float(g["photomask"]), int(g["geomask"]))
    total = same = 0
    for i, row in enumerate(g["pairs"]):
        r = int(row[0])
        for j, s in enumerate(row[1:]):
            s = int(s)
            m, d, x2, y2 = mv.check_geometric_consistency(g["depths"][r], g["ks"][r], g["es"][r], g["depths"][s],
                                                          g["ks"][s], g["es"][s], cfg)
            assert m.dtype == np.bool_ and d.dtype == np.float32
            total += m.size
            same += int((m == g["pair_mask"][i, j]).sum())
            both = m & g["pair_mask"][i, j]
            assert np.abs(d[both] - g["pair_depth_reprojected"][i, j][both]).max() < 2e-3
            assert np.all(d[~m] == 0)
            fin = np.isfinite(g["pair_x2d_src"][i, j])
            assert np.abs(x2[fin] - g["pair_x2d_src"][i, j][fin]).max() < 1e-3
            assert np.abs(y2[fin] - g["pair_y2d_src"][i, j][fin]).max() < 1e-3
    assert same / total >= 0.9999, same / total


def test_filter_fusion_matches_reference(golden):
    g = golden("filter")
    cfg = mv.FilterConfig(float(g["condmask_pixel"]), float(g["condmask_depth"]), float(g["photomask"]), int(g["geomask"]))
    photo, geo, final, avg, gsum = mv.filter_scene(g["depths"], g["conf"], g["ks"], g["es"], g["pairs"], cfg,
                                                   want_geo_sum=True)
    assert np.array_equal(photo.cpu().numpy(), g["photo"])
    assert (geo.cpu().numpy() == g["geo"]).mean() >= 0.9999
    assert (final.cpu().numpy() == g["final"]).mean() >= 0.9999
    ok = (gsum.cpu().numpy() == g["geo_sum"]) & np.isfinite(g["depth_avg"])
    assert ok.mean() > 0.999
    assert np.abs(avg.cpu().numpy()[ok] - g["depth_avg"][ok]).max() < 2e-3
    # the reference's read_pair_file structure is accepted as well, ragged source lists included
    pairs = [(int(r[0]), [int(x) for x in r[1:]]) for r in g["pairs"]]
    pairs[0] = (pairs[0][0], pairs[0][1][:2])
    p2, g2, f2, a2, _ = mv.filter_scene(g["depths"], g["conf"], g["ks"], g["es"], pairs, cfg)
    assert torch.equal(p2[1:], photo[1:]) and torch.equal(g2[1:], geo[1:]) and torch.equal(a2[1:], avg[1:])


@pytest.mark.parametrize("world", [2, 3, 8])
def test_filter_sharded_over_reference_views_equals_the_whole_scene(golden, world):
    """SURVEY 8e for K2b: reference views dealt round-robin (``shard_pairs``), depth stack replicated - the rows every
    rank computes are bit-identical to the same rows of the unsharded launch (ranks without a row stay idle)."""
    g = golden("filter")
    cfg = mv.FilterConfig(float(g["condmask_pixel"]), float(g["condmask_depth"]), float(g["photomask"]), int(g["geomask"]))
    whole = mv.filter_scene(g["depths"], g["conf"], g["ks"], g["es"], g["pairs"], cfg, want_geo_sum=True)
    seen = []
    for rank in range(world):
        rows, idx = mv.shard_pairs(list(g["pairs"]), rank, world)
        seen += idx
        if not rows:
            continue
        part = mv.filter_scene(g["depths"], g["conf"], g["ks"], g["es"], np.stack(rows), cfg, want_geo_sum=True)
        for a, b in zip(part, whole):
            assert torch.equal(a, b[idx])
    assert sorted(seen) == list(range(len(g["pairs"])))


def test_filter_same_camera_is_identity():
    h, w = 40, 56
    k = syn.intrinsics(h, w, 3)
    e = syn.extrinsics(1)
    depth = syn.smooth_depth_map(h, w, 2)
    m, d, x2, y2 = mv.check_geometric_consistency(depth, k, e, depth, k, e, mv.FilterConfig())
    assert m.all()
    assert np.abs(d - depth).max() < 1e-3
    xs, ys = np.meshgrid(np.arange(w), np.arange(h))
    assert np.abs(x2 - xs).max() < 1e-3 and np.abs(y2 - ys).max() < 1e-3


# --------------------------------------------------------------------------------------------------------------
# re-entrancy: the reference's eval path wraps the model in nn.DataParallel (one host thread per GPU,
# test_mvs4.py:393); the library keeps no global mutable state, so concurrent callers must not disturb each other
# --------------------------------------------------------------------------------------------------------------
def test_concurrent_host_threads_and_streams(golden):
    import threading
    cases = [golden(n) for n in ("k1_stage3", "k1_stage4", "k1_stage2", "k1_oob")]
    results, errors = {}, []

    def worker(i, g):
        try:
            stream = torch.cuda.Stream()
            with torch.cuda.stream(stream):
                feats = _features(g)
                for _ in range(20):
                    vol = mv.epipolar_aggregate(feats, _cuda(g["proj"]), _cuda(g["hypo"]), int(g["groups"]),
                                                float(g["attn_temp"]))
                stream.synchronize()
            results[i] = np.abs(vol.cpu().numpy() - g["volume"]).max()
        except Exception as exc:  # pragma: no cover
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(i, g)) for i, g in enumerate(cases)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert all(v < 1e-4 for v in results.values()), results


@pytest.mark.parametrize("group_cor,attn_fuse_d", [(False, True), (True, False), (False, False)])
@pytest.mark.parametrize("name", ["k1_stage2", "k1_stage4", "k1_oob"])
def test_k1_reference_option_variants(golden, name, group_cor, attn_fuse_d):
    """group_cor=False (variance cost) and attn_fuse_d=False (per-pixel weight) against the float64 oracle, and through
    the drop-in stagenet."""
    g = golden(name)
    feats = _features(g)
    proj, hypo = _cuda(g["proj"]), _cuda(g["hypo"])
    groups, temp = int(g["groups"]), float(g["attn_temp"])
    srcs = [g["srcs"][:, v] for v in range(g["srcs"].shape[1])]
    want, _, _ = O.epipolar_aggregate_np(g["ref"], srcs, g["proj"], g["hypo"], groups, temp, group_cor=group_cor,
                                         attn_fuse_d=attn_fuse_d)
    got = mv.epipolar_aggregate_variant(feats, proj, hypo, group_cor, groups, attn_fuse_d, temp)
    assert got.shape == want.shape
    assert np.abs(got.cpu().numpy() - want).max() < 1e-4
    net = mv.stagenet(inverse_depth=True, attn_fuse_d=attn_fuse_d, attn_temp=temp).eval()
    seen = {}

    def regnet(x):
        seen["x"] = x
        return x.sum(1)

    with torch.no_grad():
        net(feats, proj, hypo, regnet, 1, group_cor=group_cor, group_cor_dim=groups, split_itv=1.0)
    assert torch.equal(seen["x"], got)


@pytest.mark.parametrize("group_cor,attn_fuse_d", [(False, True), (True, False), (False, False)])
@pytest.mark.parametrize("name", ["k1_stage2", "k1_stage4", "k1_oob"])
def test_k1_option_variants_backward_matches_reference_autograd(golden, name, group_cor, attn_fuse_d):
    """Volume and feature gradients of group_cor=False / attn_fuse_d=False against the reference's own autograd
    (tests/golden/k1_modes.npz from make_golden_modes.py), through the functional form and the drop-in stagenet."""
    g, m = golden(name), golden("k1_modes")
    key = "%s/%d%d/" % (name, int(group_cor), int(attn_fuse_d))
    groups, temp = int(g["groups"]), float(g["attn_temp"])
    proj, hypo, gout = _cuda(g["proj"]), _cuda(g["hypo"]), _cuda(m[key + "gout"])
    for through_stagenet in (False, True):
        feats = _features(g, requires_grad=True)
        if through_stagenet:
            seen = {}

            def regnet(x):
                seen["x"] = x
                return x.sum(1)

            net = mv.stagenet(inverse_depth=True, attn_fuse_d=attn_fuse_d, attn_temp=temp).eval()
            net(feats, proj, hypo, regnet, 1, group_cor=group_cor, group_cor_dim=groups, split_itv=1.0)
            vol = seen["x"]
        else:
            vol = mv.epipolar_aggregate_variant(feats, proj, hypo, group_cor, groups, attn_fuse_d, temp)
        assert np.abs(vol.detach().cpu().numpy() - m[key + "volume"]).max() < 1e-4
        (vol * gout).sum().backward()
        want = [m[key + "grad_ref"]] + [m[key + "grad_srcs"][:, v] for v in range(g["srcs"].shape[1])]
        for i, (f, wg) in enumerate(zip(feats, want)):
            scale = max(1.0, np.abs(wg).max())
            assert np.abs(f.grad.cpu().numpy() - wg).max() < 2e-4 * scale, (key, i, through_stagenet)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs in one process (nn.DataParallel pattern)")
def test_second_device_in_the_same_process(golden):
    """One host process driving two GPUs (reference test_mvs4.py:393, nn.DataParallel): the device comes from the
    pointers, and per-function attributes (dynamic shared memory of the TMA kernels) are set on every device."""
    for name in ("k1_stage4", "k1_stage2"):
        g = golden(name)
        outs = []
        for dev in ("cuda:0", "cuda:1"):
            feats = [torch.from_numpy(g["ref"]).to(dev)] + \
                    [torch.from_numpy(g["srcs"][:, v]).to(dev) for v in range(g["srcs"].shape[1])]
            outs.append(mv.epipolar_aggregate(feats, torch.from_numpy(g["proj"]).to(dev),
                                              torch.from_numpy(g["hypo"]).to(dev), int(g["groups"]),
                                              float(g["attn_temp"])).cpu())
        assert torch.equal(outs[0], outs[1])
        assert np.abs(outs[1].numpy() - g["volume"]).max() < 1e-4


# --------------------------------------------------------------------------------------------------------------
# fusion tail: depth2pts + masked point selection
# --------------------------------------------------------------------------------------------------------------
def test_depth2pts_matches_reference(golden):
    g, f = golden("fusion"), golden("filter")
    for i in range(len(f["pairs"])):
        r = int(f["pairs"][i, 0])
        d32 = np.nan_to_num(f["depth_avg"][i]).astype(np.float32)   # the reference's average is float64; ours is fp32
        pts = mv.depth2pts(_cuda(d32), f["ks"][r], f["es"][r]).cpu().numpy()
        ref = g["points"][i]
        ok = np.isfinite(ref).all(1)
        assert pts.shape == ref.shape and pts.dtype == np.float64
        assert np.abs(pts - O.depth2pts_np(d32.astype(np.float64), f["ks"][r], f["es"][r])).max() < 1e-9
        assert np.abs(pts[ok] - ref[ok]).max() < 2e-4                # fp32 rounding of a ~700 mm depth
    # numpy in -> numpy out, like the reference's depth2pts_np
    out = mv.depth2pts(np.nan_to_num(f["depth_avg"][0]), f["ks"][0], f["es"][0])
    assert isinstance(out, np.ndarray) and out.shape == (f["depth_avg"][0].size, 3)


def test_fuse_scene_matches_reference_point_cloud(golden):
    """filter -> averaged depth -> world points -> masked selection, all on the GPU, against the reference's vertices
    and colours (same order: reference views in pair-file order, pixels row-major)."""
    g, f = golden("fusion"), golden("filter")
    cfg = mv.FilterConfig(float(f["condmask_pixel"]), float(f["condmask_depth"]), float(f["photomask"]), int(f["geomask"]))
    verts, cols, info = mv.fuse_scene(f["depths"], f["conf"], f["ks"], f["es"], f["pairs"], images=g["images"], config=cfg)
    assert (info["final"].cpu().numpy().astype(bool) == f["final"]).mean() >= 0.9999
    if np.array_equal(info["final"].cpu().numpy().astype(bool), f["final"]):
        assert verts.shape == g["vertices"].shape
        assert np.abs(verts.cpu().numpy() - g["vertices"]).max() < 2e-3       # averaged depth is fp32: ~1e-4 mm
        assert np.array_equal(cols.cpu().numpy(), g["colors"])


# --------------------------------------------------------------------------------------------------------------
# argument contracts added after the round-1 review
# --------------------------------------------------------------------------------------------------------------
def test_k1_entry_points_reject_mismatched_sources():
    """Every K1 entry point validates the source maps (a smaller or foreign-device map would be read out of bounds)."""
    b, h, w, c, d = 1, 16, 24, 8, 4
    ref = torch.zeros((b, h, w, c), device=DEV)
    good = torch.zeros((b, h, w, c), device=DEV)
    small = torch.zeros((b, h - 4, w, c), device=DEV)
    rt = torch.zeros((b, 2, 12), device=DEV)
    hypo = torch.ones((b, d, h, w), device=DEV)
    for bad in ([good, small], [good, good.double()], [good, good[:, :, ::2]]):
        with pytest.raises(RuntimeError, match="source features"):
            ops.epi_fwd_mode(ref, bad, rt, hypo, 4, 2.0, False, True)
        with pytest.raises(RuntimeError, match="source features"):
            ops.epi_fwd(ref, bad, rt, hypo, 4, 2.0)
    with pytest.raises(RuntimeError, match="rt must be"):
        ops.epi_fwd_mode(ref, [good, good], rt[:, :1].contiguous(), hypo, 4, 2.0, False, True)
    with pytest.raises(RuntimeError, match="depth_values|rt must be"):
        ops.homo_warp(good, torch.zeros((b + 1, 12), device=DEV), hypo)


def test_cascade_plan_replay_guards():
    from deep_reconstruction_with_epipolar_lines_mvster_b200.pipeline import CascadePlan
    plan = CascadePlan(1, 3, 64, 96, device=DEV)
    with pytest.raises(RuntimeError, match="capture"):
        plan.replay()
    for s in range(4):
        for f in plan.features[s]:
            f.normal_()
        plan.proj[s].copy_(torch.from_numpy(syn.proj_matrices(1, 3, 64, 96, s)))
    plan.depth_values.copy_(torch.from_numpy(syn.depth_values(1)))
    plan.capture()
    d0, _ = plan.replay()
    d0 = d0.clone()
    d1, _ = plan.run()
    assert torch.equal(d0, d1)
    plan.regnet = lambda s, vol: plan.logits[s]
    with pytest.raises(RuntimeError, match="regnet"):
        plan.replay()
