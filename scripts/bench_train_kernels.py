#!/usr/bin/env python
"""Training-mode kernels of the regulariser at its stage-4 training shapes (512x640, B=2, D=4): fused BatchNorm + ReLU
forward / backward against nn.BatchNorm3d + ReLU (cuDNN), hand-written Conv3d / ConvTranspose3d weight gradients against
cuDNN's (aten.convolution_backward).  One JSON line per case; `--once` runs every hand-written kernel once (for ncu).

    python scripts/bench_train_kernels.py [--iters 20] [--once]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deep_reconstruction_with_epipolar_lines_mvster_b200 import ops  # noqa: E402
from scripts.bench_extra import timed  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--once", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    b, d, h, w = 2, 4, 512, 640
    # ---- BatchNorm + ReLU: conv0's output [2,8,4,512,640] and conv2's [2,16,4,256,320] --------------------------------
    for c, hh, ww in ((8, h, w), (16, h // 2, w // 2), (64, h // 8, w // 8)):
        x = torch.randn(b, c, d, hh, ww, device=dev)
        gy = torch.randn_like(x)
        bn = torch.nn.BatchNorm3d(c).to(dev).train()
        y, mean, invstd = ops.bn_train_fwd(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, 0.1, 1e-5, True)
        ops.bn_train_bwd(x, y, gy, bn.weight, mean, invstd, True)
        if args.once:
            continue
        nbytes = x.numel() * 4
        f = timed(lambda: ops.bn_train_fwd(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, 0.1, 1e-5, True), args.iters)
        bw = timed(lambda: ops.bn_train_bwd(x, y, gy, bn.weight, mean, invstd, True), args.iters)
        xr = x.clone().requires_grad_(True)

        def ref_fwd():
            return F.relu(bn(xr))
        rf = timed(ref_fwd, args.iters)
        yr = ref_fwd()
        rb = timed(lambda: torch.autograd.grad(yr, [xr, bn.weight, bn.bias], gy, retain_graph=True), args.iters)
        print(json.dumps({"bench": "bn_relu_train", "shape": [b, c, d, hh, ww], "fwd_ms": f, "bwd_ms": bw,
                          "fwd_GBps": 3 * nbytes / f / 1e6, "bwd_GBps": 7 * nbytes / bw / 1e6,
                          "cudnn_fwd_ms": rf, "cudnn_bwd_ms": rb,
                          "bytes_model": "fwd: 2 reads + 1 write of the activation; bwd: 6 reads + 1 write"}))
    # ---- weight gradients ----------------------------------------------------------------------------------------------
    cases = [("conv0", 4, 8, 1, 1, False, h, w), ("conv1", 8, 16, 1, 2, False, h, w), ("conv2", 16, 16, 3, 1, False, h // 2, w // 2),
             ("conv4", 32, 32, 3, 1, False, h // 4, w // 4), ("conv6", 64, 64, 3, 1, False, h // 8, w // 8),
             ("conv11", 16, 8, 1, 2, True, h // 2, w // 2)]
    for name, cin, cout, kd, st, tr, hh, ww in cases:
        x = torch.randn(b, cin, d, hh, ww, device=dev)
        if tr:
            wt = torch.randn(cin, cout, 1, 3, 3, device=dev)
            y = F.conv_transpose3d(x, wt, None, (1, 2, 2), (0, 1, 1), (0, 1, 1))
        else:
            wt = torch.randn(cout, cin, kd, 3, 3, device=dev)
            y = F.conv3d(x, wt, None, (1, st, st), (kd // 2, 1, 1))
        gy = torch.randn_like(y)
        a, bb = (x, gy) if tr else (gy, x)
        ops.conv3d_wgrad(a, bb, kd, st)
        if args.once:
            continue
        ms = timed(lambda: ops.conv3d_wgrad(a, bb, kd, st), args.iters)
        ref = timed(lambda: torch.ops.aten.convolution_backward(gy, x, wt, None, [1, st, st], [kd // 2, 1, 1], [1, 1, 1], tr,
                                                                [0, 1, 1] if tr else [0, 0, 0], 1, [False, True, False]),
                    args.iters)
        fma = a.shape[0] * a.shape[2] * a.shape[3] * a.shape[4] * cin * cout * kd * 9
        print(json.dumps({"bench": "conv3d_wgrad", "layer": name, "cin": cin, "cout": cout, "kd": kd, "stride": st,
                          "transposed": tr, "input": [b, cin, d, hh, ww], "ms": ms, "TFMA_per_s": fma / ms / 1e9,
                          "cudnn_ms": ref, "speedup": ref / ms}))


if __name__ == "__main__":
    main()
