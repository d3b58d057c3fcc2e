#!/usr/bin/env python
"""Per-kernel CUDA time of one MVS4net.forward (832x1152, N=5, one scene) on the B200 path - where the whole-network
time goes (FPN4 / reg2d are cuDNN; K1 / tails are this library).  Prints a table sorted by total CUDA time."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deep_reconstruction_with_epipolar_lines_mvster_b200 as mv
from deep_reconstruction_with_epipolar_lines_mvster_b200 import synthetic as syn
sys.path.insert(0, os.path.join(ROOT, "scripts"))
from bench_extra import NET_CFG

dev = torch.device("cuda", 0)
tf32 = "--tf32" in sys.argv
torch.backends.cudnn.allow_tf32 = tf32
h0, w0, n, b = 832, 1152, 5, 1
model = mv.MVS4net(**NET_CFG).eval()
model.load_state_dict(syn.fill_state_dict(model.state_dict(), seed=7))
model = model.to(dev)
imgs = [torch.rand((b, 3, h0, w0), device=dev) for _ in range(n)]
proj = {k: torch.from_numpy(v).to(dev) for k, v in syn.proj_matrices_all_stages(b, n, h0, w0).items()}
dv = torch.from_numpy(syn.depth_values(b)).to(dev)
if "--once" in sys.argv:   # for ncu: a single forward, nothing else
    with torch.no_grad():
        model(imgs, proj, dv)
    torch.cuda.synchronize()
    print("one forward done")
    sys.exit(0)
with torch.no_grad():
    for _ in range(3):
        model(imgs, proj, dv)
    torch.cuda.synchronize()
    # coarse split with CUDA events
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ev[0].record(); feats = model.extract_features(imgs); ev[1].record(); model(imgs, proj, dv); ev[2].record()
    torch.cuda.synchronize()
    print("extract_features %.2f ms, whole forward %.2f ms (tf32=%s)" % (ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), tf32))
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
        model(imgs, proj, dv)
        torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=48,
                                                        max_shapes_column_width=60))
