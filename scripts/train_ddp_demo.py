#!/usr/bin/env python
"""Config 4 of BASELINE.json: the fused op forward+backward at the DTU-train shape (512x640, B=2 per GPU, N=5) inside
torch DistributedDataParallel (NCCL).  The trainable part is a small per-stage 1x1 feature head in front of the
fused op (the op itself has no parameters), so DDP has gradients to all-reduce.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/train_ddp_demo.py
    ... scripts/train_ddp_demo.py --full    # the whole network instead: MVS4net.train() + MVS4net_loss (fused Sinkhorn, K3)
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


def full_network(rank, local, world, dev):
    """One process per GPU, B=2 scenes per GPU at 512x640, N=5: ``MVS4net.train()`` forward (cuDNN convolutions around
    the fused K1 / tail), ``MVS4net_loss`` (one fused Sinkhorn launch per stage), backward, DDP all-reduce, Adam."""
    from deep_reconstruction_with_epipolar_lines_mvster_b200 import loss as L
    torch.backends.cudnn.benchmark = True
    b, n, h0, w0 = 2, 5, 512, 640
    model = mv.MVS4net(group_cor=True, group_cor_dim=[8, 8, 4, 4], inverse_depth=True, attn_temp=2.0).train()
    model.load_state_dict(syn.fill_state_dict(model.state_dict(), seed=7))   # same initial weights on every rank
    model = model.to(dev)
    if world > 1:
        model = nn.parallel.DistributedDataParallel(model, device_ids=[local])
    gen = torch.Generator(device=dev).manual_seed(100 + rank)                # different scenes on every rank
    imgs = [torch.rand((b, 3, h0, w0), device=dev, generator=gen) for _ in range(n)]
    proj = {k: torch.from_numpy(v).to(dev) for k, v in syn.proj_matrices_all_stages(b, n, h0, w0).items()}
    dv = torch.from_numpy(syn.depth_values(b)).to(dev)
    gts, masks = {}, {}
    for s in range(4):
        h, w = h0 >> (3 - s), w0 >> (3 - s)
        gts["stage%d" % (s + 1)] = 560 + 300 * torch.rand((b, h, w), device=dev, generator=gen)
        masks["stage%d" % (s + 1)] = (torch.rand((b, h, w), device=dev, generator=gen) > 0.2).float()
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = L.MVS4net_loss(model(imgs, proj, dv), gts, masks, inverse_depth=True, ot_iter=10, ot_eps=1)[0]
        loss.backward()
        opt.step()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 5
    a.record()
    for _ in range(iters):
        loss = step()
    e.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(e) / iters], device=dev)
    w0_ = next(model.parameters()).detach().flatten()[:64].clone()
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ref = w0_.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(ref, w0_), "ranks diverged: DDP did not synchronise the gradients"
    if rank == 0:
        print(json.dumps({"bench": "train_ddp_full_network", "n_gpus": world,
                          "config": "512x640 B=2/GPU N=5, MVS4net.train + MVS4net_loss(ot_iter=10) + bwd + Adam; fused K1/tail/K3 in fp32, "
                                    "cuDNN convolutions with torch defaults (TF32 allowed), cudnn.benchmark",
                          "ms_per_step": float(ms), "samples_per_s": b * world / float(ms) * 1e3, "loss": float(loss.detach())}))


def main():
    rank, local, world = mv.rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", init_method="env://", device_id=dev)
    if "--full" in sys.argv:
        full_network(rank, local, world, dev)
        if world > 1:
            dist.destroy_process_group()
        return
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
