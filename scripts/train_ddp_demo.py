#!/usr/bin/env python
"""Config 4 of BASELINE.json: the fused op forward+backward at the DTU-train shape (512x640, B=2 per GPU, N=5) inside
torch DistributedDataParallel (NCCL).  The trainable part is a small per-stage 1x1 feature head in front of the
fused op (the op itself has no parameters), so DDP has gradients to all-reduce.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/train_ddp_demo.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deep_reconstruction_with_epipolar_lines_mvster_b200 as mv  # noqa: E402
from deep_reconstruction_with_epipolar_lines_mvster_b200 import synthetic as syn  # noqa: E402


class Heads(nn.Module):
    def __init__(self):
        super().__init__()
        self.heads = nn.ModuleList([nn.Conv2d(c, c, 1) for c in syn.STAGE_CHANNELS])

    def forward(self, feats, projs, hypos):
        loss = 0.0
        for s in range(4):
            f = [self.heads[s](x) for x in feats[s]]
            vol = mv.epipolar_aggregate(f, projs[s], hypos[s], syn.STAGE_GROUPS[s], 2.0)
            loss = loss + vol.square().mean()
        return loss


def main():
    rank, local, world = mv.rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", init_method="env://", device_id=dev)
    b, n, h0, w0 = 2, 5, 512, 640
    feats, projs, hypos = [], [], []
    for s in range(4):
        h, w = syn.stage_shape(h0, w0, s)
        feats.append([syn.smooth_features(b, syn.STAGE_CHANNELS[s], h, w, 100 * rank + 10 * s + v, device=dev)
                      .contiguous(memory_format=torch.channels_last) for v in range(n)])
        projs.append(torch.from_numpy(syn.proj_matrices(b, n, h0, w0, s)).to(dev))
        dv = torch.from_numpy(syn.depth_values(b)).to(dev)
        hypos.append(mv.init_inverse_range(dv, syn.STAGE_NDEPTHS[s], None, None, h, w))
    model = Heads().to(dev).to(memory_format=torch.channels_last)
    if world > 1:
        model = nn.parallel.DistributedDataParallel(model, device_ids=[local])
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = model(feats, projs, hypos)
        loss.backward()
        opt.step()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 20
    a.record()
    for _ in range(iters):
        loss = step()
    e.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(e) / iters], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"bench": "train_ddp_fused_op", "n_gpus": world, "config": "512x640 B=2/GPU N=5 fp32, 4 stages fwd+bwd + Adam",
                          "ms_per_step": float(ms), "samples_per_s": b * world / float(ms) * 1e3, "loss": float(loss)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
