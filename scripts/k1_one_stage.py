#!/usr/bin/env python
"""Run K1 forward of one stage shape a few times on smooth synthetic hypotheses (for ncu captures).

    python scripts/k1_one_stage.py [--stage 3] [--iters 3] [--batch 8]
"""
import argparse, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "scripts"))
import bench_extra as be
from deep_reconstruction_with_epipolar_lines_mvster_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--stage", type=int, default=3)
ap.add_argument("--iters", type=int, default=3)
ap.add_argument("--batch", type=int, default=8)
a = ap.parse_args()
dev = torch.device("cuda", 0)
feats, proj, hypo, g, d, c, h, w = be.stage_inputs(a.batch, 5, 864, 1152, a.stage, dev)
nhwc = [ops.to_nhwc(f) for f in feats]
rt = ops.compose_homographies(proj)
for _ in range(a.iters):
    out = ops.epi_fwd(nhwc[0], nhwc[1:], rt, hypo, g, 2.0)
torch.cuda.synchronize()
print("ok", tuple(out[0].shape) if isinstance(out, tuple) else tuple(out.shape))
