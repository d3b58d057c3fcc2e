"""Pin the K3 oracle (``sinkhorn_np`` / ``sinkhorn_port`` in oracle/mvster_oracle.py) against the unmodified reference
(``tests/golden/sinkhorn.npz``, written by ``tests/golden/make_golden_sinkhorn.py`` from models/mvs4net_utils.py:1164-1210
and models/MVS4Net.py:195-240), and check the host-side mirror of the loss API."""
import inspect

import numpy as np
import pytest
import torch

from oracle import mvster_oracle as O

CASES = ["d4_it3_e1", "d8_it3_e1", "d8_it10_e01", "d4_it10_e1_cont", "d8_it3_e1_cont", "d4_it0_e1"]


def _case(g, name):
    d, iters, eps, cont = g[name + "/cfg"]
    return {k: g["%s/%s" % (name, k)] for k in ("gt", "hypo", "attn", "mask", "tmap", "loss", "grad")}, int(d), int(iters), float(eps), bool(cont)


def test_fixture_lists_every_case(golden):
    assert list(golden("sinkhorn")["cases"]) == CASES


@pytest.mark.parametrize("name", CASES)
def test_float64_restatement_matches_reference(golden, name):
    c, d, iters, eps, cont = _case(golden("sinkhorn"), name)
    r = O.sinkhorn_np(c["gt"], c["hypo"], c["attn"], c["mask"], iters, eps, cont)
    assert r["T_map"].shape == c["tmap"].shape
    assert np.abs(r["T_map"] - c["tmap"]).max() < 2e-5 * max(1.0, np.abs(c["tmap"]).max())
    assert abs(r["loss"] - float(c["loss"])) < 2e-5 * abs(float(c["loss"]))
    # analytic reverse sweep == the reference's autograd
    assert np.abs(r["grad_attn"] - c["grad"]).max() < 1e-4 * max(np.abs(c["grad"]).max(), 1e-6)
    assert r["count"] == int(c["mask"].sum())


@pytest.mark.parametrize("name", CASES)
def test_port_matches_reference(golden, name):
    c, d, iters, eps, cont = _case(golden("sinkhorn"), name)
    a = torch.from_numpy(c["attn"]).requires_grad_(True)
    tmap, loss = O.sinkhorn_port(torch.from_numpy(c["gt"]), torch.from_numpy(c["hypo"]), a,
                                 torch.from_numpy(c["mask"]), iters, eps, cont)
    assert np.abs(tmap.detach().numpy() - c["tmap"]).max() < 1e-6 * max(1.0, np.abs(c["tmap"]).max())
    assert abs(float(loss) - float(c["loss"])) < 1e-6 * abs(float(c["loss"]))
    if iters > 0:
        grad, = torch.autograd.grad(loss, a)
        assert np.abs(grad.numpy() - c["grad"]).max() < 1e-6


def test_range_err_ratio_matches_reference_loss(golden):
    g = golden("sinkhorn")
    total = 0.0
    for s in range(1, 5):
        k = "loss4/stage%d/" % s
        r = O.sinkhorn_np(g[k + "gt"], g[k + "hypo"], g[k + "attn"], g[k + "mask"] > 0.5, 10, 1.0, False,
                          inverse_depth=True)
        assert abs(r["range_err_ratio"] - float(g[k + "ratio"])) < 1e-6
        assert abs(r["loss"] - float(g[k + "ot"])) < 2e-5 * float(g[k + "ot"])
        assert np.abs(r["grad_attn"] - g[k + "grad"]).max() < 1e-4 * np.abs(g[k + "grad"]).max()
        total += r["loss"]
    assert abs(total - float(g["loss4/total"])) < 1e-4


def test_empty_mask_is_nan_like_the_reference(golden):
    c, d, iters, eps, cont = _case(golden("sinkhorn"), "d4_it3_e1")
    r = O.sinkhorn_np(c["gt"], c["hypo"], c["attn"], np.zeros_like(c["mask"]), iters, eps, cont)
    assert np.isnan(r["loss"]) and r["count"] == 0 and not r["grad_attn"].any()


def test_loss_api_mirrors_the_reference_signatures():
    from deep_reconstruction_with_epipolar_lines_mvster_b200 import loss
    assert list(inspect.signature(loss.sinkhorn).parameters) == ["gt_depth", "hypo_depth", "attn_weight", "mask",
                                                                 "iters", "eps", "continuous"]
    assert list(inspect.signature(loss.MVS4net_loss).parameters) == ["inputs", "depth_gt_ms", "mask_ms", "kwargs"]


def test_loss_has_no_cpu_fallback(golden):
    from deep_reconstruction_with_epipolar_lines_mvster_b200 import loss
    c, d, iters, eps, cont = _case(golden("sinkhorn"), "d4_it3_e1")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        loss.sinkhorn(torch.from_numpy(c["gt"]), torch.from_numpy(c["hypo"]), torch.from_numpy(c["attn"]),
                      torch.from_numpy(c["mask"]), iters, eps, cont)
