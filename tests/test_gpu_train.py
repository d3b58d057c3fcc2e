"""One TRAINING step on the B200 path against the unmodified reference (tests/golden/train.npz, written by
tests/golden/make_golden_train.py): ``MVS4net.train()`` forward, fused-Sinkhorn ``MVS4net_loss``, backward.

The chain under test: FPN4 / reg2d (cuDNN, training-mode BatchNorm) -> K1 forward + hand-written backward -> tail
forward + backward -> K3 fused OT loss + gradient.  Stages 2-4 are free-running (each consumes the arg-max of the stage
before), so a rare arg-max flip between fp32 implementations moves their hypotheses; stage 1 is compared tightly, the
total loss and the parameter gradients with the looser bounds stated below."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import deep_reconstruction_with_epipolar_lines_mvster_b200 as mv
from deep_reconstruction_with_epipolar_lines_mvster_b200 import loss as L, synthetic as syn

from test_network_host import CFG

DEV = "cuda"
LOSS_KW = dict(stage_lw=[1, 1, 1, 1], l1ot_lw=[0, 1], inverse_depth=True, ot_iter=10, ot_eps=1, ot_continous=False,
               mono=False)


@pytest.fixture(autouse=True)
def _fp32_convs():
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


def _step(g):
    model = mv.MVS4net(**CFG).train()
    model.load_state_dict(syn.fill_state_dict(model.state_dict(), seed=7))
    model = model.to(DEV)
    b, n = g["imgs"].shape[1], g["imgs"].shape[0]
    imgs = [torch.from_numpy(g["imgs"][v]).to(DEV) for v in range(n)]
    proj = {k: torch.from_numpy(v).to(DEV) for k, v in syn.proj_matrices_all_stages(b, n, 64, 128).items()}
    gts = {"stage%d" % s: torch.from_numpy(g["gt_stage%d" % s]).to(DEV) for s in range(1, 5)}
    masks = {"stage%d" % s: torch.from_numpy(g["mask_stage%d" % s]).to(DEV) for s in range(1, 5)}
    out = model(imgs, proj, torch.from_numpy(g["depth_values"]).to(DEV))
    total, l1s, ots, ratios = L.MVS4net_loss(out, gts, masks, **LOSS_KW)
    total.backward()
    return model, out, total, ots, ratios


def test_training_step_matches_reference(golden):
    g = golden("train")
    model, out, total, ots, ratios = _step(g)
    # stage 1 does not depend on any arg-max: tight
    assert np.abs(out["stage1"]["attn_weight"].detach().cpu().numpy() - g["attn_stage1"]).max() < 1e-4
    assert abs(float(ots[0].detach()) - g["ots"][0]) < 1e-4 * g["ots"][0]
    assert abs(float(ratios[0]) - g["ratios"][0]) < 1e-6
    for s in range(1, 4):
        assert abs(float(ots[s].detach()) - g["ots"][s]) < 5e-3 * g["ots"][s], (s, float(ots[s].detach()), g["ots"][s])
        assert abs(float(ratios[s]) - g["ratios"][s]) < 5e-3
    assert abs(float(total.detach()) - float(g["total"])) < 2e-3 * float(g["total"])
    params = dict(model.named_parameters())
    for name in g["watch"]:
        ref = g["grad/" + str(name)]
        got = params[str(name)].grad.cpu().numpy()
        assert got.shape == ref.shape and np.isfinite(got).all()
        rel = np.linalg.norm(got - ref) / max(np.linalg.norm(ref), 1e-12)
        assert rel < 2e-2, (str(name), rel)
    gn = torch.sqrt(sum((p.grad ** 2).sum() for p in model.parameters() if p.grad is not None))
    assert abs(float(gn) - float(g["grad_norm"])) < 2e-2 * float(g["grad_norm"])


def test_training_steps_reduce_the_loss(golden):
    """A few Adam steps on one batch lower the fused OT loss (the gradient points downhill)."""
    g = golden("train")
    model = mv.MVS4net(**CFG).train()
    model.load_state_dict(syn.fill_state_dict(model.state_dict(), seed=7))
    model = model.to(DEV)
    b, n = g["imgs"].shape[1], g["imgs"].shape[0]
    imgs = [torch.from_numpy(g["imgs"][v]).to(DEV) for v in range(n)]
    proj = {k: torch.from_numpy(v).to(DEV) for k, v in syn.proj_matrices_all_stages(b, n, 64, 128).items()}
    gts = {"stage%d" % s: torch.from_numpy(g["gt_stage%d" % s]).to(DEV) for s in range(1, 5)}
    masks = {"stage%d" % s: torch.from_numpy(g["mask_stage%d" % s]).to(DEV) for s in range(1, 5)}
    dv = torch.from_numpy(g["depth_values"]).to(DEV)
    opt = torch.optim.Adam(model.parameters(), lr=2e-3)
    first = None
    for it in range(12):
        opt.zero_grad(set_to_none=True)
        # stage 1 only: later stages move their hypotheses between steps, which makes their loss non-monotone
        total = L.MVS4net_loss(model(imgs, proj, dv), gts, masks, **dict(LOSS_KW, stage_lw=[1, 0, 0, 0]))[0]
        total.backward()
        opt.step()
        first = float(total.detach()) if first is None else first
    assert float(total.detach()) < first - 1e-3, (first, float(total.detach()))
