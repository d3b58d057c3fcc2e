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
from deep_reconstruction_with_epipolar_lines_mvster_b200 import loss as L, ops, synthetic as syn

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


@pytest.mark.parametrize("shape,relu", [((2, 8, 4, 16, 20), True), ((2, 16, 8, 6, 10), True), ((3, 64, 4, 5, 4), False),
                                        ((2, 8, 4, 64, 80), True), ((1, 32, 136, 100), True)])
def test_fused_training_batchnorm_matches_torch(shape, relu):
    """``mvster_bn_train_fwd`` / ``mvster_bn_train_bwd`` against nn.BatchNorm + ReLU in float64: output, running
    statistics, num_batches_tracked, gradients of x / weight / bias (tolerances: 2e-5 of the range)."""
    from deep_reconstruction_with_epipolar_lines_mvster_b200 import network as N
    torch.manual_seed(sum(shape))
    cls = torch.nn.BatchNorm3d if len(shape) == 5 else torch.nn.BatchNorm2d
    x = (torch.randn(shape) * 2.0 + 0.7)
    gout = torch.randn(shape)
    ref_bn = cls(shape[1]).double().train()
    with torch.no_grad():
        ref_bn.weight.copy_(torch.rand(shape[1]) + 0.5)
        ref_bn.bias.copy_(torch.randn(shape[1]) * 0.3)
    bn = cls(shape[1]).train()
    bn.load_state_dict({k: v.float() if v.is_floating_point() else v for k, v in ref_bn.state_dict().items()})
    bn = bn.to(DEV)
    xr = x.double().requires_grad_(True)
    yr = ref_bn(xr)
    yr = torch.relu(yr) if relu else yr
    (yr * gout.double()).sum().backward()
    xg = x.to(DEV).requires_grad_(True)
    launches = mv.launch_count()
    yg = N.bn_act(bn, xg, relu)
    assert mv.launch_count() - launches == 3                       # the fused kernels ran, not cuDNN
    (yg * gout.to(DEV)).sum().backward()
    assert mv.launch_count() - launches == 6
    tol = lambda t: 2e-5 * max(1.0, float(t.abs().max()))
    assert (yg.detach().cpu().double() - yr.detach()).abs().max().item() < tol(yr)
    assert (xg.grad.cpu().double() - xr.grad).abs().max().item() < tol(xr.grad)
    assert (bn.weight.grad.cpu().double() - ref_bn.weight.grad).abs().max().item() < 1e-4 * max(1.0, float(ref_bn.weight.grad.abs().max()))
    assert (bn.bias.grad.cpu().double() - ref_bn.bias.grad).abs().max().item() < 1e-4 * max(1.0, float(ref_bn.bias.grad.abs().max()))
    assert (bn.running_mean.cpu().double() - ref_bn.running_mean).abs().max().item() < 1e-6
    assert (bn.running_var.cpu().double() - ref_bn.running_var).abs().max().item() < 1e-5
    assert int(bn.num_batches_tracked) == int(ref_bn.num_batches_tracked) == 1
    # bit-reproducible: fixed-order partial sums, no atomics
    xg2 = x.to(DEV).requires_grad_(True)
    yg2 = N.bn_act(bn, xg2, relu)
    (yg2 * gout.to(DEV)).sum().backward()
    assert torch.equal(yg2, yg) and torch.equal(xg2.grad, xg.grad)


def test_fused_training_batchnorm_falls_back_where_it_does_not_apply():
    from deep_reconstruction_with_epipolar_lines_mvster_b200 import network as N
    bn = torch.nn.BatchNorm2d(8).to(DEV).train()
    x = torch.randn(2, 8, 6, 10, device=DEV).contiguous(memory_format=torch.channels_last)
    launches = mv.launch_count()
    y = N.bn_act(bn, x, True)                                      # channels_last: cuDNN's NHWC kernels
    assert mv.launch_count() == launches and y.shape == x.shape
    bn.eval()
    y = N.bn_act(bn, torch.randn(2, 8, 6, 10, device=DEV), True)   # eval: running statistics
    assert mv.launch_count() == launches
    with pytest.raises(RuntimeError):
        ops.bn_train_fwd(torch.randn(2, 8, 3, 3, device=DEV), None, None, None, None, 0.1, 1e-5, True)   # plane of 9


def test_training_step_with_and_without_fused_batchnorm_agree(golden):
    from deep_reconstruction_with_epipolar_lines_mvster_b200 import network as N
    g = golden("train")
    totals = []
    for fused in (True, False):
        N.FUSED_TRAIN_BATCHNORM = fused
        try:
            totals.append(float(_step(g)[2].detach()))
        finally:
            N.FUSED_TRAIN_BATCHNORM = True
    assert abs(totals[0] - totals[1]) < 2e-3 * abs(totals[1]), totals


@pytest.mark.parametrize("cin,cout,kd,stride,transposed,d,h,w", [
    (4, 8, 1, 1, False, 4, 16, 40), (8, 16, 1, 2, False, 3, 18, 70), (8, 16, 1, 2, False, 2, 17, 33),
    (16, 16, 3, 1, False, 4, 9, 35), (32, 64, 1, 2, False, 2, 8, 12), (64, 64, 3, 1, False, 3, 5, 6),
    (64, 32, 1, 2, True, 2, 6, 10), (16, 8, 1, 2, True, 4, 70, 36)])
def test_conv3d_hand_wgrad_matches_torch_autograd(cin, cout, kd, stride, transposed, d, h, w):
    """``mvster_conv3d_wgrad`` through ``network.conv3d_train`` against float64 autograd of the same layer: weight
    gradient (2e-5 of its range), data gradient and output (cuDNN, fp32)."""
    from deep_reconstruction_with_epipolar_lines_mvster_b200 import network as N
    torch.manual_seed(cin * 7 + cout + kd + h)
    if transposed:
        conv = torch.nn.ConvTranspose3d(cin, cout, (1, 3, 3), padding=(0, 1, 1), output_padding=(0, 1, 1), stride=(1, 2, 2), bias=False)
    else:
        conv = torch.nn.Conv3d(cin, cout, (kd, 3, 3), stride=(1, stride, stride), padding=(kd // 2, 1, 1), bias=False)
    x = torch.randn(2, cin, d, h, w)
    import copy
    conv64 = copy.deepcopy(conv).double()
    xr = x.double().requires_grad_(True)
    yr = conv64(xr)
    gout = torch.randn(yr.shape)
    (yr * gout.double()).sum().backward()
    conv = conv.to(DEV).train()
    xg = x.to(DEV).requires_grad_(True)
    launches = mv.launch_count()
    yg = N.conv3d_train(conv, xg)
    (yg * gout.to(DEV)).sum().backward()
    assert mv.launch_count() - launches == 2                       # wgrad + reduce ran
    assert (yg.detach().cpu().double() - yr.detach()).abs().max().item() < 1e-4 * max(1.0, float(yr.abs().max()))
    gw, gwr = conv.weight.grad.cpu().double(), conv64.weight.grad
    assert gw.shape == gwr.shape
    assert (gw - gwr).abs().max().item() < 2e-5 * max(1.0, float(gwr.abs().max())), (gw - gwr).abs().max().item()
    assert (xg.grad.cpu().double() - xr.grad).abs().max().item() < 1e-4 * max(1.0, float(xr.grad.abs().max()))
    # reproducible bit for bit
    conv.weight.grad = None
    xg2 = x.to(DEV).requires_grad_(True)
    (N.conv3d_train(conv, xg2) * gout.to(DEV)).sum().backward()
    assert torch.equal(conv.weight.grad.cpu().double(), gw)


def test_training_step_with_and_without_hand_wgrad_agree(golden):
    from deep_reconstruction_with_epipolar_lines_mvster_b200 import network as N
    g = golden("train")
    grads = []
    for hand in (True, False):
        N.HAND_WGRAD3D = N.HAND_WGRAD2D = hand
        try:
            model = _step(g)[0]
            grads.append({k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None})
        finally:
            N.HAND_WGRAD3D = N.HAND_WGRAD2D = True
    assert grads[0].keys() == grads[1].keys() and len(grads[0]) > 20
    for k in grads[0]:
        a, b = grads[0][k], grads[1][k]
        # prob.bias has a mathematically zero gradient (softmax ignores a constant shift): absolute floor for it
        assert (a - b).norm().item() <= 1e-3 * b.norm().item() + 1e-6, k


def test_prob_hand_wgrad_matches_torch_autograd():
    from deep_reconstruction_with_epipolar_lines_mvster_b200 import network as N
    import copy
    torch.manual_seed(3)
    conv = torch.nn.Conv3d(8, 1, 1)
    x, gout = torch.randn(2, 8, 4, 18, 22), torch.randn(2, 1, 4, 18, 22)
    c64 = copy.deepcopy(conv).double()
    xr = x.double().requires_grad_(True)
    (c64(xr) * gout.double()).sum().backward()
    conv = conv.to(DEV).train()
    xg = x.to(DEV).requires_grad_(True)
    launches = mv.launch_count()
    (N.prob_train(conv, xg) * gout.to(DEV)).sum().backward()
    assert mv.launch_count() - launches == 2
    assert (conv.weight.grad.cpu().double() - c64.weight.grad).abs().max().item() < 2e-5 * max(1.0, float(c64.weight.grad.abs().max()))
    assert (conv.bias.grad.cpu().double() - c64.bias.grad).abs().max().item() < 2e-5 * max(1.0, float(c64.bias.grad.abs().max()))
    assert (xg.grad.cpu().double() - xr.grad).abs().max().item() < 1e-6


@pytest.mark.parametrize("cin,cout,h,w,cl", [(3, 8, 20, 36, True), (16, 16, 9, 33, True), (64, 32, 6, 10, False), (64, 8, 12, 40, True)])
def test_conv2d_hand_wgrad_matches_torch_autograd(cin, cout, h, w, cl):
    """FPN4's 3x3 layers in training: weight gradient on ``mvster_conv3d_wgrad`` (D = 1) over planar copies, forward and
    data gradient on cuDNN in the activations' own memory format (channels_last in ``MVS4net.train()``)."""
    from deep_reconstruction_with_epipolar_lines_mvster_b200 import network as N
    import copy
    torch.manual_seed(cin + cout + h)
    conv = torch.nn.Conv2d(cin, cout, 3, padding=1, bias=False)
    x, c64 = torch.randn(2, cin, h, w), copy.deepcopy(conv).double()
    xr = x.double().requires_grad_(True)
    yr = c64(xr)
    gout = torch.randn(yr.shape)
    (yr * gout.double()).sum().backward()
    conv = conv.to(DEV).train()
    xg = x.to(DEV)
    if cl:
        conv = conv.to(memory_format=torch.channels_last)
        xg = xg.contiguous(memory_format=torch.channels_last)
    xg.requires_grad_(True)
    launches = mv.launch_count()
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False                        # the hand-written 2-D path is the fp32-exact mode's
    try:
        (N.conv2d_train(conv, xg) * gout.to(DEV)).sum().backward()
    finally:
        torch.backends.cudnn.allow_tf32 = tf32
    assert mv.launch_count() - launches == 2
    gw = conv.weight.grad
    assert gw.shape == conv.weight.shape and gw.stride() == conv.weight.stride()
    assert (gw.cpu().double() - c64.weight.grad).abs().max().item() < 2e-5 * max(1.0, float(c64.weight.grad.abs().max()))
    assert (xg.grad.cpu().double() - xr.grad).abs().max().item() < 1e-4 * max(1.0, float(xr.grad.abs().max()))
