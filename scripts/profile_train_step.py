#!/usr/bin/env python
"""torch-profiler kernel table of one whole-network training step on the B200 path (512x640, B=2, N=5, fp32).

    python scripts/profile_train_step.py [--tf32] > profiles/<tag>_train_step_torch_profiler.txt
"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deep_reconstruction_with_epipolar_lines_mvster_b200 as mv  # noqa: E402
from deep_reconstruction_with_epipolar_lines_mvster_b200 import loss as L, synthetic as syn  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "scripts"))
from bench_extra import NET_CFG  # noqa: E402

tf32 = "--tf32" in sys.argv
torch.backends.cudnn.allow_tf32 = tf32
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda", 0)
h0, w0, n, b = 512, 640, 5, 2
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


def step():
    opt.zero_grad(set_to_none=True)
    total = L.MVS4net_loss(model(imgs, proj, dv), gts, masks, **kw)[0]
    total.backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
step()
ev[1].record()
torch.cuda.synchronize()
print("one training step %.2f ms (tf32=%s, cudnn.benchmark=True)" % (ev[0].elapsed_time(ev[1]), tf32))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90))
