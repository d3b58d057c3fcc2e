#!/usr/bin/env python
"""K2b over the GPUs of one box: the depth / confidence stack of a scene is replicated on every GPU,
the REFERENCE views of the pair list are dealt round-robin (`shard_pairs`), every rank runs `filter_scene` on its rows
(test_mvs4.py:694-749 for those reference views) and the per-view masks / averaged depths are gathered on rank 0
(an all_gather of outputs only: 13 bytes per pixel and reference view; the path itself has no exchange step).

    python scripts/filter_sharded.py                                     # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/filter_sharded.py

Prints one JSON line: time per scene (device-timed, max over ranks), and whether the gathered result is bit-identical
to the one-GPU result computed on rank 0.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import deep_reconstruction_with_epipolar_lines_mvster_b200 as mv  # noqa: E402
from deep_reconstruction_with_epipolar_lines_mvster_b200 import synthetic as syn  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=49)
    ap.add_argument("--srcs", type=int, default=9)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--scenes", type=int, default=1, help="scenes filtered back to back inside the timed region")
    args = ap.parse_args()
    rank, local, world = mv.rank_world()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    h, w, v, s = 512, 640, args.views, args.srcs
    k = syn.intrinsics(h, w, 3)
    es = np.stack([syn.grid_extrinsics(i, 7, 0.04) for i in range(v)])
    ks = np.stack([k] * v)
    cache = "/tmp/mvster_filter_depths_%d.npy" % v                                    # 15 s of CPU rendering: once per box
    if os.path.exists(cache):
        depths = np.load(cache)
    else:
        depths = syn.render_surface_depths(k, list(es), h, w, noise_mm=0.3, seed=0)   # same scene on every rank
        if rank == 0:
            np.save(cache + ".tmp.npy", depths)
            os.replace(cache + ".tmp.npy", cache)
    conf = np.random.RandomState(0).uniform(0, 1, size=(v, h, w)).astype(np.float32)
    pairs = np.concatenate([np.arange(v)[:, None], syn.pair_list(v, s)], 1).astype(np.int32)
    dz, cf = torch.from_numpy(depths).to(dev), torch.from_numpy(conf).to(dev)         # replicated stack
    cfg = mv.FilterConfig()
    rows, idx = mv.shard_pairs(list(pairs), rank, world)
    mine = np.stack(rows) if rows else np.zeros((0, 1 + s), np.int32)

    def run():
        out = None
        for _ in range(args.scenes):
            out = mv.filter_scene(dz, cf, ks, es, mine, cfg) if len(mine) else None
        return out

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.iters):
        out = run()
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / args.iters / args.scenes], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)

    # gather (outside the timed region: it is the output hand-over, not the path) and compare with one GPU
    per = (v + world - 1) // world
    final = torch.zeros((per, h, w), dtype=torch.uint8, device=dev)
    avg = torch.zeros((per, h, w), dtype=torch.float32, device=dev)
    if out is not None:
        final[:len(mine)] = out[2].to(torch.uint8)
        avg[:len(mine)] = out[3]
    if world > 1:
        fl = [torch.empty_like(final) for _ in range(world)]
        al = [torch.empty_like(avg) for _ in range(world)]
        dist.all_gather(fl, final)
        dist.all_gather(al, avg)
    else:
        fl, al = [final], [avg]
    if rank == 0:
        full = mv.filter_scene(dz, cf, ks, es, pairs, cfg)
        f1, a1 = full[2].to(torch.uint8), full[3]
        same_mask = same_avg = True
        for r in range(world):
            ids = list(range(r, v, world))
            same_mask &= bool(torch.equal(fl[r][:len(ids)], f1[ids]))
            same_avg &= bool(torch.equal(al[r][:len(ids)].view(torch.int32), a1[ids].view(torch.int32)))
        t = float(ms.item())
        print(json.dumps({"bench": "filter_%dx%d_512x640_sharded" % (v, s), "n_gpus": world, "ms_per_scene": t,
                          "scenes_per_s": 1e3 / t, "pair_checks_per_s": v * s * 1e3 / t,
                          "ref_views_per_rank": [len(range(r, v, world)) for r in range(world)],
                          "sharding": "reference views round-robin (shard_pairs), depth stack replicated, no collective inside the timed region",
                          "masks_bit_identical_to_one_gpu": same_mask, "averaged_depth_bit_identical_to_one_gpu": same_avg,
                          "final_mask_mean": float(f1.float().mean())}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
