"""Pin the CPU oracle (oracle/mvster_oracle.py) against outputs of the unmodified reference (tests/golden/*.npz)."""
import numpy as np
import pytest
import torch

from oracle import mvster_oracle as O

K1_CASES = ["k1_stage1", "k1_stage2", "k1_stage3", "k1_stage4", "k1_oob"]


@pytest.mark.parametrize("name", K1_CASES)
def test_k1_float64_restatement_matches_reference(golden, name):
    g = golden(name)
    srcs = [g["srcs"][:, v] for v in range(g["srcs"].shape[1])]
    vol, wts, _ = O.epipolar_aggregate_np(g["ref"], srcs, g["proj"], g["hypo"], int(g["groups"]),
                                          float(g["attn_temp"]))
    # the reference itself is fp32: its distance to exact math is ~1e-5
    assert np.abs(vol - g["volume"]).max() < 3e-5
    assert np.abs(wts - g["weights"]).max() < 3e-5


@pytest.mark.parametrize("name", K1_CASES)
def test_k1_port_matches_reference_and_grads(golden, name):
    g = golden(name)
    feats = [torch.from_numpy(g["ref"]).requires_grad_(True)]
    feats += [torch.from_numpy(g["srcs"][:, v].copy()).requires_grad_(True) for v in range(g["srcs"].shape[1])]
    vol, wts = O.epipolar_aggregate_port(feats, torch.from_numpy(g["proj"]), torch.from_numpy(g["hypo"]),
                                         int(g["groups"]), float(g["attn_temp"]), return_weights=True)
    assert torch.allclose(vol.detach(), torch.from_numpy(g["volume"]), atol=2e-6, rtol=0)
    assert np.abs(torch.stack(wts, 1).detach().numpy() - g["weights"]).max() < 2e-6
    (vol * torch.from_numpy(g["gout"])).sum().backward()
    assert np.abs(feats[0].grad.numpy() - g["grad_ref"]).max() < 1e-5
    for v in range(g["srcs"].shape[1]):
        assert np.abs(feats[1 + v].grad.numpy() - g["grad_srcs"][:, v]).max() < 1e-5


def test_homo_warping(golden):
    g = golden("warp")
    out = O.homo_warping_np(g["src"], g["src_proj"], g["ref_proj"], g["hypo"])
    assert np.abs(out - g["warped"]).max() < 5e-5
    port = O.homo_warping_port(torch.from_numpy(g["src"]), torch.from_numpy(g["src_proj"]),
                               torch.from_numpy(g["ref_proj"]), torch.from_numpy(g["hypo"]))
    assert np.abs(port.numpy() - g["warped"]).max() < 2e-6


def test_schedule(golden):
    g = golden("schedule")
    init = O.init_inverse_range_np(g["depth_values"], 8, 6, 7)
    assert np.abs(init - g["init"]).max() / np.abs(g["init"]).max() < 2e-7
    assert init[0, 0, 0, 0] > init[0, -1, 0, 0]          # index 0 = farthest hypothesis
    for d in (4, 8):
        s = O.schedule_inverse_range_np(g["inv_min"], g["inv_max"], d, 14, 18)
        assert np.abs(s - g["sched%d" % d]).max() / np.abs(s).max() < 5e-7


def test_tail(golden):
    g = golden("tail")
    for mode in ("eval", "train"):
        out = O.tail_np(g["logits"], g["hypo"], float(g["split_itv"]), training=(mode == "train"))
        assert np.array_equal(out["depth"], g[mode + "_depth"])                  # arg-max gather: exact
        assert np.abs(out["attn_weight"] - g[mode + "_attn_weight"]).max() < 1e-6
        for k in ("inverse_min_depth", "inverse_max_depth"):
            assert np.abs(out[k] - g[mode + "_" + k]).max() < 1e-9
        conf = g[mode + "_photometric_confidence"]
        if mode == "train":
            assert conf.shape == () and conf == 0.0
        else:
            ref = conf
            got = out["photometric_confidence"]
            ok = np.isfinite(ref)
            assert np.allclose(got[ok], ref[ok], rtol=2e-5, atol=1e-6)


def test_filter_pairs_and_fusion(golden):
    g = golden("filter")
    pairs = g["pairs"]
    total = 0
    same = 0
    for i, row in enumerate(pairs):
        r = int(row[0])
        for j, s in enumerate(row[1:]):
            s = int(s)
            for use_cv2 in (False, True):
                m, d, x2, y2 = O.check_geometric_consistency_np(
                    g["depths"][r], g["ks"][r], g["es"][r], g["depths"][s], g["ks"][s], g["es"][s],
                    float(g["condmask_pixel"]), float(g["condmask_depth"]), use_cv2=use_cv2)
                if use_cv2:
                    assert np.array_equal(m, g["pair_mask"][i, j])
                    assert np.array_equal(d, g["pair_depth_reprojected"][i, j])
                else:
                    total += m.size
                    same += int((m == g["pair_mask"][i, j]).sum())
                    both = m & g["pair_mask"][i, j]
                    assert np.abs(d[both] - g["pair_depth_reprojected"][i, j][both]).max() < 2e-3
                assert np.array_equal(x2, g["pair_x2d_src"][i, j], equal_nan=True)
                assert np.array_equal(y2, g["pair_y2d_src"][i, j], equal_nan=True)
    assert same / total >= 0.9999, same / total
    photo, geo, final, avg, gsum = O.filter_fuse_np(g["depths"], g["conf"], g["ks"], g["es"], pairs,
                                                    float(g["condmask_pixel"]), float(g["condmask_depth"]),
                                                    float(g["photomask"]), int(g["geomask"]))
    assert np.array_equal(photo, g["photo"])
    assert (geo == g["geo"]).mean() >= 0.9999
    assert (final == g["final"]).mean() >= 0.9999
    ok = (gsum == g["geo_sum"]) & np.isfinite(g["depth_avg"])
    assert ok.mean() > 0.99
    assert np.abs(avg[ok] - g["depth_avg"][ok]).max() < 2e-3


def test_remap_emulation_matches_cv2():
    import cv2
    rng = np.random.RandomState(0)
    img = rng.uniform(0, 100, size=(37, 53)).astype(np.float32)
    mx = rng.uniform(-3, 56, size=(64, 64)).astype(np.float32)
    my = rng.uniform(-3, 40, size=(64, 64)).astype(np.float32)
    mx[0, :8] = np.arange(8)                  # exact integers
    my[0, :8] = 5.0
    mx[1, :4] = [52.0, 52.5, 53.0, -1.0]      # borders
    my[1, :4] = [36.0, 36.5, 37.0, -0.5]
    ref = cv2.remap(img, mx, my, interpolation=cv2.INTER_LINEAR)
    got = O.remap_linear_np(img, mx, my)
    assert np.abs(ref - got).max() < 2e-4


def test_filter_512x640_oracle_matches_the_lifted_reference():
    """Full-size pin of the filter oracle: three 512x640 views, all pairs (tests/golden/make_golden_filter512.py)."""
    import hashlib, os
    from deep_reconstruction_with_epipolar_lines_mvster_b200 import synthetic as syn
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "filter512.npz"))
    depths, conf, ks, es, pairs = syn.filter_fixture_512()
    sha = hashlib.sha256()
    for a in (depths, conf, ks, es, pairs):
        sha.update(np.ascontiguousarray(a).tobytes())
    assert sha.hexdigest() == str(g["inputs_sha256"])
    w = depths.shape[2]
    want = np.unpackbits(g["pair_mask_bits"], axis=-1)[..., :w].astype(bool)
    total = same = 0
    for i, row in enumerate(pairs[:2]):            # two reference views = four pairs keep the CPU suite short
        r = int(row[0])
        for j, s in enumerate(row[1:]):
            s = int(s)
            m, d, _, _ = O.check_geometric_consistency_np(depths[r], ks[r], es[r], depths[s], ks[s], es[s],
                                                          float(g["condmask_pixel"]), float(g["condmask_depth"]))
            total += m.size
            same += int((m == want[i, j]).sum())
            both = (m & want[i, j])[::4, ::4]
            assert np.abs(d[::4, ::4][both] - g["pair_depth_reprojected_s4"][i, j][both]).max() < 2e-3
    assert same / total >= 0.9999, same / total
