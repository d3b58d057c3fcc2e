#!/usr/bin/env python
"""Benchmark of the MVSTER cost-volume hot path on B200 (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path (default)
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference algorithm on the host CPU cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W                  # N > 1: scenes sharded, no data-path collective

Metric (BASELINE.json): depth maps/s on the DTU 864x1152 N=5 shape.  One depth map = the 4-stage cascade of
``MVS4net.forward`` restricted to the hot path: per stage the hypothesis schedule, the homography composition, the
fused warp/correlation/attention/aggregation kernel (K1) and the fused depth/confidence tail (K2a).  The regnet
(cuDNN 3-D U-Net, out of scope) is replaced by pre-computed synthetic logits resident on the device.

One step = one batch of ``--scenes`` (default 8) synthetic scenes per GPU (weak scaling: every GPU processes its own
batch; the ranks exchange nothing but the timing reduction).  ``value`` = scenes * ranks * steps / max-over-ranks
device time, inputs resident in HBM.  ``e2e`` = the same through ``CascadePlan.run_from_host``: every step copies
the features / cameras from pinned host memory and the depth + confidence maps back.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from deep_reconstruction_with_epipolar_lines_mvster_b200 import synthetic as syn  # noqa: E402

METRIC = "depth_maps_per_s_dtu_864x1152_n5"
UNIT = "depth maps/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--scenes", type=int, default=8, help="scenes (depth maps) per GPU per step")
    ap.add_argument("--height", type=int, default=864)
    ap.add_argument("--width", type=int, default=1152)
    ap.add_argument("--views", type=int, default=5)
    ap.add_argument("--dtype", choices=["fp32", "bf16"], default="fp32", help="feature storage type")
    ap.add_argument("--e2e-steps", type=int, default=None, help="steps of the host-buffer leg (default min(steps,10))")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="issue the launches of a step eagerly instead of replaying them as one CUDA graph")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-network", action="store_true",
                    help="skip the secondary whole-network (FPN4 + reg2d + hot path) measurement on rank 0 at N=1")
    ap.add_argument("--sustain-s", type=float, default=2.0, help="length of the sustained leg in seconds")
    ap.add_argument("--net-scenes", type=int, default=4, help="scenes per call of the images-in leg (e2e_network)")
    ap.add_argument("--net-steps", type=int, default=8, help="timed steps of the images-in leg")
    ap.add_argument("--no-numa-bind", action="store_true",
                    help="do not pin each rank (N > 1) to the CPUs local to its GPU before allocating pinned buffers")
    ap.add_argument("--cpu-scenes", type=int, default=8, help="timed scenes of the CPU baseline sample")
    return ap.parse_args()


def profiled_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture (profiles/)."""
    path = os.path.join(ROOT, "profiles", "k1_stage4_traffic.json")
    if not os.path.exists(path):
        return None, None
    with open(path) as f:
        t = json.load(f)
    return t["dram_bytes_read"] + t["dram_bytes_write"], t["source"]


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ---------------------------------------------------------------------------------------------------------------------
# clocks: sampled with NVML while the timed region runs
# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, device):
        self.samples, self.reasons = [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            try:  # CUDA_VISIBLE_DEVICES renumbers torch's devices; the UUID does not change
                uuid = "GPU-" + str(torch.cuda.get_device_properties(device).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(device.index or 0)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _once(self):
        try:
            self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
            try:
                r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for bit, name in self.REASONS.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _loop(self):
        while not self._stop.is_set():
            self._once()
            self._stop.wait(0.01)

    def start(self):
        if self.nv is None:
            return
        self._stop.clear()
        self._thread = threading.Thread(target=self._loop, daemon=True)
        self._thread.start()

    def stop(self):
        if self.nv is None:
            return
        self._once()
        self._stop.set()
        self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------------------
# synthetic workload
# ---------------------------------------------------------------------------------------------------------------------
def scene_seed(rank: int, index: int) -> int:
    """Seed of scene ``index`` of rank ``rank``'s batch - the ONE place both arms take their scenes from."""
    return 1234 + 1000 * rank + index


def scene_gt_inverse_depth(seed, h, w):
    """Smooth synthetic ground-truth surface (inverse depth) of one scene: what a trained regnet would converge to."""
    return torch.from_numpy((1.0 / syn.smooth_depth_map(h, w, seed, lo=560.0, hi=800.0)).astype(np.float32))[None]


def tracking_logits(hypo, inv_gt, noise):
    """Stand-in for regnet: logits peaked at the hypothesis nearest the smooth ground truth, plus a little noise, so
    that the cascade's depth maps are piecewise smooth as on real data (white-noise logits would make neighbouring
    pixels sample texels hundreds of pixels apart, which no MVS network produces)."""
    inv = 1.0 / hypo
    itv = (inv[:, 1:2] - inv[:, 0:1]).abs().clamp_min(1e-12)
    return -4.0 * (inv - inv_gt[:, None]).abs() / itv + noise


def make_scene(seed, nviews, h0, w0, jitter_index=0):
    """One synthetic scene (B=1, CPU tensors), used by BOTH arms: FPN-like feature maps ``features[s][v]`` [1,C,H,W]
    (N(0, 0.5^2), 3x3 box-smoothed - SURVEY 8d), cameras ``projs[s]`` [1,N,2,4,4] (Appendix B rig; ``jitter_index``
    varies the baseline by 2 % per scene of a batch), the depth range, the ground-truth inverse depth and the logit
    noise per stage."""
    feats, projs, gts, noises = [], [], [], []
    g = torch.Generator().manual_seed(seed)
    for s in range(4):
        h, w = syn.stage_shape(h0, w0, s)
        feats.append([syn.smooth_features(1, syn.STAGE_CHANNELS[s], h, w, seed * 100 + 10 * s + v) for v in range(nviews)])
        projs.append(torch.from_numpy(syn.proj_matrices(jitter_index + 1, nviews, h0, w0, s,
                                                         per_batch_jitter=0.02)[jitter_index:jitter_index + 1]))
        gts.append(scene_gt_inverse_depth(seed, h, w))
        noises.append(torch.randn((1, syn.STAGE_NDEPTHS[s], h, w), generator=g) * 0.5)
    return {"features": feats, "projs": projs, "depth_values": torch.from_numpy(syn.depth_values(1)),
            "gt": gts, "noise": noises}


def fill_plan(plan, rank: int):
    """Upload rank ``rank``'s batch of scenes (``make_scene(scene_seed(rank, i))``) and pre-compute the stand-in regnet
    outputs for the hypotheses the cascade really visits (one untimed set-up pass)."""
    gts = [torch.empty((plan.B,) + tuple(sh), device=plan.device) for sh in plan.shapes]
    noises = [torch.empty_like(l) for l in plan.logits]
    for i in range(plan.B):
        sc = make_scene(scene_seed(rank, i), plan.N, plan.h0, plan.w0, jitter_index=i)
        for s in range(plan.nstage):
            for v in range(plan.N):
                plan.features[s][v][i].copy_(sc["features"][s][v][0].permute(1, 2, 0))
            plan.proj[s][i].copy_(sc["projs"][s][0])
            gts[s][i].copy_(sc["gt"][s][0])
            noises[s][i].copy_(sc["noise"][s][0])
        plan.depth_values[i].copy_(sc["depth_values"][0])
        del sc

    def regnet(s, _volume):
        plan.logits[s].copy_(tracking_logits(plan.hypo[s], gts[s], noises[s]))
        return plan.logits[s]

    plan.regnet = regnet
    plan.run()
    plan.regnet = None
    torch.cuda.synchronize()


def run_reference_arm(args, steps, warmup, device="cpu"):
    """The reference's own implementation of the path on ``device``: the UNMODIFIED reference (``oracle/_ref``) through
    ``MVS4net.forward`` when it is present (kind "reference"), else the oracle's op-for-op torch port (kind "port").
    One step = one scene of rank 0's batch (``make_scene(scene_seed(0, i))``, the scenes the GPU arm uploads)."""
    from oracle import ref_arm
    cores = os.cpu_count() or 1
    if device == "cpu":
        torch.set_num_threads(cores)
    nscene = max(1, min(args.scenes, warmup + steps))
    scenes = [make_scene(scene_seed(0, i), args.views, args.height, args.width, jitter_index=i) for i in range(nscene)]
    runners = []
    for sc in scenes:
        mk = [lambda hypo, gt=sc["gt"][s].to(device), nz=sc["noise"][s].to(device): tracking_logits(hypo, gt, nz) for s in range(4)]
        if ref_arm.available():
            runners.append(ref_arm.ReferenceCascade(torch, sc["features"], sc["projs"], sc["depth_values"], mk,
                                                    syn.STAGE_GROUPS, syn.STAGE_NDEPTHS, syn.STAGE_SPLIT_ITV, 2.0, device).run)
        else:
            from oracle import mvster_oracle as O
            runners.append(lambda sc=sc, mk=mk: O.cascade_port(sc["features"], sc["projs"], sc["depth_values"], mk,
                                                               syn.STAGE_GROUPS, syn.STAGE_NDEPTHS, syn.STAGE_SPLIT_ITV, 2.0))
    kind = "reference" if ref_arm.available() else "port"
    sync = (lambda: torch.cuda.synchronize()) if device != "cpu" else (lambda: None)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            sync()
            t0 = time.perf_counter()
            runners[i % nscene]()
            sync()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    total = sum(times)
    what = ("the unmodified reference (oracle/_ref: MVS4net.forward -> stage loop, schedule_inverse_range, stagenet / "
            "homo_warping / attention / tail; FPN4 and reg2d replaced by the same stand-ins as in the GPU arm)"
            if kind == "reference" else "oracle torch port of the reference op sequence (oracle/_ref absent)")
    return {"value": steps / total, "unit": UNIT, "cores": cores if device == "cpu" else 0, "kind": kind,
            "sample": "%d scene(s) (B=1 per step, %d warm-up) of rank 0's batch of the same %dx%d N=%d 4-stage workload: %s, "
                      "%s" % (steps, warmup, args.height, args.width, args.views, what,
                              ("%d host threads" % cores) if device == "cpu" else "eager PyTorch on the same GPU"),
            "ms_per_step": 1e3 * total / steps}


def run_network(args, dev):
    """Images in, 4-stage depth out: checkpoint-compatible MVS4net (random-init weights from the deterministic recipe)
    with FPN4 / reg2d on this library's direct-convolution and fused kernels + cuDNN for the wide layers, one scene of
    N views per call.  The reference loader snaps the height to a multiple of 64 (864 -> 832, SURVEY finding 5)."""
    import deep_reconstruction_with_epipolar_lines_mvster_b200 as mv
    from deep_reconstruction_with_epipolar_lines_mvster_b200 import synthetic as syn
    h0, w0, n = (args.height // 64) * 64, (args.width // 64) * 64, args.views
    if h0 < 64 or w0 < 64:
        return None
    old_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        model = mv.MVS4net(group_cor=True, group_cor_dim=[8, 8, 4, 4], inverse_depth=True, attn_temp=2.0).eval()
        model.load_state_dict(syn.fill_state_dict(model.state_dict(), seed=7))
        model = model.to(dev)
        gen = torch.Generator(device=dev).manual_seed(0)
        imgs = [torch.rand((1, 3, h0, w0), device=dev, generator=gen) for _ in range(n)]
        proj = {k: torch.from_numpy(v).to(dev) for k, v in syn.proj_matrices_all_stages(1, n, h0, w0).items()}
        dv = torch.from_numpy(syn.depth_values(1)).to(dev)
        with torch.no_grad():
            for _ in range(3):
                model(imgs, proj, dv)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 10
            a.record()
            for _ in range(reps):
                model(imgs, proj, dv)
            b.record()
            torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        del model, imgs
        torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32 = old_tf32
    return {"value": 1e3 / ms, "unit": UNIT, "ms_per_depth_map": ms,
            "workload": "MVS4net.forward, images in -> 4-stage depth out, %dx%d N=%d, one scene per call, fp32 "
                        "(cuDNN TF32 off)" % (h0, w0, n),
            "what": "FPN4 + reg2d (hand-written direct / fused convolutions, tcgen05 implicit GEMMs for the "
                    "32/64-channel layers, linearised FPN top-down behind one cuBLAS projection GEMM; cuDNN only for "
                    "the 1x1 convolutions at 1/8 resolution) around the hot path; secondary number, not part of `value`"}


def run_network_e2e(args, dev, barrier, world):
    """Images in, depth out, from HOST memory, on every rank: ``net_scenes`` scenes per call through ``GraphedMVS4net``
    (one CUDA-graph replay of the whole ``MVS4net.forward``: FPN4 + 4 x (schedule, K1, reg2d, tail)); every step copies
    the pinned images / cameras in and the final depth + confidence maps out.  This is the path a user of
    ``test_mvs4.py:406-416`` has; it is compute-bound (57 MB of images per scene), so it scales with the GPU count where
    the feature-upload leg saturates the host.  Returns per-rank milliseconds for ``steps`` steps."""
    import deep_reconstruction_with_epipolar_lines_mvster_b200 as mv
    h0, w0, n, bsz = (args.height // 64) * 64, (args.width // 64) * 64, args.views, args.net_scenes
    old_tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    try:
        model = mv.MVS4net(group_cor=True, group_cor_dim=[8, 8, 4, 4], inverse_depth=True, attn_temp=2.0).eval()
        model.load_state_dict(syn.fill_state_dict(model.state_dict(), seed=7))
        model = model.to(dev)
        gen = torch.Generator().manual_seed(11)
        pin = lambda t: t.pin_memory()
        imgs = [pin(torch.rand((bsz, 3, h0, w0), generator=gen)) for _ in range(n)]
        proj = {k: pin(torch.from_numpy(v)) for k, v in syn.proj_matrices_all_stages(bsz, n, h0, w0, per_batch_jitter=0.02).items()}
        dv = pin(torch.from_numpy(syn.depth_values(bsz)))
        h_depth = torch.empty((bsz, h0, w0), dtype=torch.float32).pin_memory()
        h_conf = torch.empty((bsz, h0, w0), dtype=torch.float32).pin_memory()
        g = mv.GraphedMVS4net(model, bsz, n, h0, w0, dev)

        g.capture(imgs, proj, dv)
        g.prefetch(imgs, proj, dv)

        def step():
            # steady state of a serving loop: this step's inputs were handed over by the previous step's prefetch and
            # crossed the link while that step computed; every step uploads one full set of inputs (the next request's),
            # runs one forward and reads its depth + confidence maps back to the host
            out = g.run_prefetched()["stage4"]
            h_depth.copy_(out["depth"], non_blocking=True)
            h_conf.copy_(out["photometric_confidence"], non_blocking=True)
            g.prefetch(imgs, proj, dv)
            torch.cuda.current_stream(dev).synchronize()

        for _ in range(3):
            step()
        steps = args.net_steps
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            step()
        b.record()
        barrier()
        ms = a.elapsed_time(b)
        h2d = sum(t.numel() * 4 for t in imgs) + sum(t.numel() * 4 for t in proj.values()) + dv.numel() * 4
        d2h = (h_depth.numel() + h_conf.numel()) * 4
        assert torch.isfinite(h_depth).all()
        del g, model
        torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32 = old_tf32
    return {"ms": ms, "steps": steps, "scenes": bsz, "h2d": int(h2d), "d2h": int(d2h), "h0": h0, "w0": w0}


def run_hotpath_e2e_bf16(args, dev, rank, barrier):
    """The host-buffer leg with bf16 feature maps (half the upload); K1 accumulates in fp32.  Returns (ms, plan)."""
    from deep_reconstruction_with_epipolar_lines_mvster_b200.pipeline import CascadePlan
    plan = CascadePlan(args.scenes, args.views, args.height, args.width, device=dev, feature_dtype=torch.bfloat16)
    fill_plan(plan, rank)
    plan.make_host_buffers()
    for hf, df in zip(plan.h_features, plan.features):
        for a, b in zip(hf, df):
            a.copy_(b)
    for a, b in zip(plan.h_proj, plan.proj):
        a.copy_(b)
    plan.h_depth_values.copy_(plan.depth_values)
    for _ in range(2):
        plan.run_from_host()
    steps = args.e2e_steps or min(args.steps, 10)
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        plan.run_from_host()
    b.record()
    barrier()
    out = {"ms": a.elapsed_time(b), "steps": steps, "h2d": plan.h2d_bytes(), "d2h": plan.d2h_bytes()}
    del plan
    torch.cuda.empty_cache()
    return out


def workload_config(args, world):
    return {"workload": "DTU eval shape N=%d views %dx%d, batch of %d synthetic scenes per GPU (configs[2])"
                        % (args.views, args.height, args.width, args.scenes),
            "stages": "C=64/32/16/8 G=8/8/4/4 D=8/8/4/4 at 1/8,1/4,1/2,1/1 resolution",
            "scenes_per_gpu_per_step": args.scenes, "global_scenes_per_step": args.scenes * world,
            "parallelism": "scenes sharded over %d GPU(s), no data-path collective" % world,
            "regnet": "out of scope (cuDNN); stand-in logits resident on device, peaked at a smooth synthetic ground-truth "
                      "surface + noise so the cascade's depth maps are piecewise smooth as on real data",
            "l2_policy": "inputs larger than L2: %.0f MB of features per step vs 126 MB L2, no flush"
                         % (args.scenes * feature_mb(args)),
            "note_864_vs_832": "the reference loader snaps 864 to 832 (SURVEY.md finding 5); the fused op is benchmarked "
                               "at the nominal 864x1152 named by the metric"}


def feature_mb(args):
    s = 2 if args.dtype == "bf16" else 4
    tot = 0
    for st in range(4):
        h, w = syn.stage_shape(args.height, args.width, st)
        tot += args.views * syn.STAGE_CHANNELS[st] * h * w * s
    return tot / 1e6


# ---------------------------------------------------------------------------------------------------------------------
def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        # the reference's own CPU implementation of the path (oracle/_ref, else the oracle port): rank 0 only, the
        # other ranks exit quietly.  Same config dictionary as the GPU arm (the workload is the same; what one step of
        # this arm samples from it is said in cpu_baseline.sample).
        if rank != 0:
            return 0
        res = run_reference_arm(args, max(1, args.steps), max(0, args.warmup))
        line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args, max(world, args.gpus)),
                "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return 0

    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (the MVSTER B200 path has no CPU fallback); "
                           "use --impl reference for the CPU arm")
    import torch.distributed as dist
    from deep_reconstruction_with_epipolar_lines_mvster_b200 import _lib
    from deep_reconstruction_with_epipolar_lines_mvster_b200.pipeline import CascadePlan

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from deep_reconstruction_with_epipolar_lines_mvster_b200.sharding import bind_to_gpu_numa_node
    # only with several ranks per box: at N=1 there is no contention, and the CPU baseline wants every host core
    numa_cpus = bind_to_gpu_numa_node(local_rank) if (world > 1 and not args.no_numa_bind) else None
    if world > 1:
        dist.init_process_group("nccl", init_method="env://", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fdt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    plan = CascadePlan(args.scenes, args.views, args.height, args.width, device=dev, feature_dtype=fdt)
    fill_plan(plan, rank)
    dom = plan.nstage - 1  # dominant kernel: stage-4 K1 forward
    sampler = ClockSampler(dev)

    # ---- value leg: inputs resident in HBM ------------------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        plan.run()
    pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    use_graph = args.graph
    if use_graph:  # the launches of a step replayed as one CUDA graph
        try:
            plan.capture()
        except Exception as exc:  # capture is an optimisation: fall back to eager launches and say so in the line
            print("[bench] CUDA graph capture failed (%s); eager launches" % exc, file=sys.stderr)
            use_graph = False
    barrier()
    launches0 = _lib.launch_count()
    sampler.start()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    if use_graph:
        for i in range(args.steps):
            plan.replay()
    else:
        for i in range(args.steps):
            plan.stage_events = pairs[i]
            plan.run(time_stage=dom)
    t_end.record()
    barrier()
    sampler.stop()
    plan.stage_events = None
    launches = _lib.launch_count() - launches0
    ms_total = t_start.elapsed_time(t_end)
    if use_graph:
        # a graph replay issues LAUNCHES_PER_RUN kernels without passing through the library's launch counter, and
        # events cannot bracket a kernel inside it: the dominant kernel is timed in a separate eager pass
        launches = args.steps * plan.LAUNCHES_PER_RUN
        for p_ in pairs:
            plan.stage_events = p_
            plan.run(time_stage=dom)
        torch.cuda.synchronize()
        plan.stage_events = None
    k1_ms = statistics.mean(a.elapsed_time(b) for a, b in pairs)

    # ---- sustained leg: the same step back to back for ~2 s (the timed region above lasts tens of milliseconds, too
    #      short for the clocks to leave boost or for the sampler to see more than a handful of readings) ---------------
    sus_steps = max(args.steps, int(args.sustain_s * 1e3 / max(ms_total / args.steps, 1e-3)))
    sus_sampler = ClockSampler(dev)
    barrier()
    sus_sampler.start()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(sus_steps):
        if use_graph:
            plan.replay()
        else:
            plan.run()
    s1.record()
    barrier()
    sus_sampler.stop()
    sus_ms = s0.elapsed_time(s1)
    # every stage's K1 launch, bracketed the same way (eager passes after the timed region)
    stage_ms = []
    for st in range(plan.nstage):
        if st == dom:
            stage_ms.append(k1_ms)
            continue
        sp = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(min(args.steps, 20))]
        for p_ in sp:
            plan.stage_events = p_
            plan.run(time_stage=st)
        torch.cuda.synchronize()
        plan.stage_events = None
        stage_ms.append(statistics.mean(a.elapsed_time(b) for a, b in sp))

    # ---- e2e leg: pinned host buffers in, depth + confidence out, every step ---------------------------------------
    e2e_steps = args.e2e_steps or min(args.steps, 10)
    plan.make_host_buffers()
    for hf, df in zip(plan.h_features, plan.features):
        for a, b in zip(hf, df):
            a.copy_(b)
    for a, b in zip(plan.h_proj, plan.proj):
        a.copy_(b)
    plan.h_depth_values.copy_(plan.depth_values)
    for _ in range(3):
        plan.run_from_host()
    barrier()
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record()
    for _ in range(e2e_steps):
        plan.run_from_host()
    e_end.record()
    barrier()
    e2e_ms = e_start.elapsed_time(e_end)

    # ---- secondary host-buffer legs on every rank: bf16 feature upload, and images in -> depth out --------------------
    del plan.h_features
    e2e16 = net = None
    if args.dtype == "fp32" and not args.no_network:
        e2e16 = run_hotpath_e2e_bf16(args, dev, rank, barrier)
        net = run_network_e2e(args, dev, barrier, world)

    # ---- max over ranks -------------------------------------------------------------------------------------------
    extra = [e2e16["ms"] if e2e16 else 0.0, net["ms"] if net else 0.0, sus_ms]
    t = torch.tensor([ms_total, e2e_ms, k1_ms] + extra + stage_ms, device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, k1_ms, e2e16_ms, net_ms, sus_ms = (float(x) for x in t.tolist()[:6])
    stage_ms = [float(x) for x in t.tolist()[6:]]

    # ---- secondary: the whole MVS4net.forward around the hot path (SURVEY 8d row ii), rank 0 at N=1 only ---------------
    network = None
    if rank == 0 and world == 1 and not args.no_network and args.dtype == "fp32":
        network = run_network(args, dev)

    cpu = gpu_ref = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import ref_arm
        if ref_arm.available():
            # SURVEY 8d row iii: the stock reference on this same GPU (eager ATen kernels), same scenes, B=1 per call
            try:
                gpu_ref = run_reference_arm(args, 8, 2, device=str(dev))
                gpu_ref = {"value": gpu_ref["value"], "unit": UNIT, "ms_per_depth_map": gpu_ref["ms_per_step"],
                           "kind": gpu_ref["kind"], "sample": gpu_ref["sample"]}
            except Exception as exc:  # a baseline, never a reason to lose the bench line
                print("[bench] gpu_reference failed: %s" % exc, file=sys.stderr)
            torch.cuda.empty_cache()
        cpu = run_reference_arm(args, args.cpu_scenes, 1)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}

    if rank == 0:
        peak, peak_src = peaks()
        alg_bytes = plan.k1_bytes(dom)
        achieved = alg_bytes / (k1_ms * 1e-3) / 1e9
        fma_ms = plan.k1_fmas(dom) / 37.2e12 * 1e3
        traffic, traffic_src = profiled_traffic()
        if (args.scenes, args.height, args.width, args.views, args.dtype) != (8, 864, 1152, 5, "fp32"):
            traffic, traffic_src = None, None  # the capture was taken at the default workload only
        line = {
            "metric": METRIC, "value": args.scenes * world * args.steps / (ms_total * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.dtype == "fp32" else "bf16 features, f32 accumulate", "data": "synthetic",
            "config": workload_config(args, world),
            "host_affinity": ("each rank bound to its GPU's local NUMA CPUs (%d cpus on rank 0)" % len(numa_cpus))
                             if numa_cpus else "unbound (single rank, topology not exposed, or --no-numa-bind)",
            "clocks": sampler.summary(),
            "e2e": {"value": args.scenes * world * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": plan.h2d_bytes(), "d2h_bytes_per_step": plan.d2h_bytes(),
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "api": "CascadePlan.run_from_host (pinned host features/cameras in, depth+confidence out)"},
            "gpu_launches": int(launches), "cuda_graph": bool(use_graph),
            "sustained": {"value": args.scenes * world * sus_steps / (sus_ms * 1e-3), "unit": UNIT, "steps": sus_steps,
                          "seconds": sus_ms * 1e-3, "ms_per_step": sus_ms / sus_steps, "clocks": sus_sampler.summary()},
            "roofline": {"kernel": "epi_fwd_box_kernel<C=8,CPG=2,D=4> (stage-4 K1 forward)", "bound": "hbm",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                         "kernel_ms": k1_ms, "fp32_fma_bound_ms": fma_ms,
                         "hbm_bound_ms": alg_bytes / (peak * 1e9) * 1e3,
                         "per_stage": [{"stage": st + 1, "kernel_ms": stage_ms[st],
                                        "algorithmic_bytes_per_launch": plan.k1_bytes(st),
                                        "achieved": plan.k1_bytes(st) / (stage_ms[st] * 1e-3) / 1e9,
                                        "frac": plan.k1_bytes(st) / (stage_ms[st] * 1e-3) / 1e9 / peak}
                                       for st in range(plan.nstage)]},
        }
        if e2e16 is not None:
            line["e2e_bf16"] = {"value": args.scenes * world * e2e16["steps"] / (e2e16_ms * 1e-3), "unit": UNIT,
                                "h2d_bytes_per_step": e2e16["h2d"], "d2h_bytes_per_step": e2e16["d2h"],
                                "steps": e2e16["steps"], "ms_per_step": e2e16_ms / e2e16["steps"],
                                "api": "CascadePlan(feature_dtype=bfloat16).run_from_host: bf16 feature maps uploaded "
                                       "(half the bytes), fp32 accumulation in K1; secondary, not the headline"}
        if net is not None:
            line["e2e_network"] = {"value": net["scenes"] * world * net["steps"] / (net_ms * 1e-3), "unit": UNIT,
                                   "h2d_bytes_per_step": net["h2d"], "d2h_bytes_per_step": net["d2h"],
                                   "steps": net["steps"], "ms_per_step": net_ms / net["steps"],
                                   "scenes_per_gpu_per_step": net["scenes"],
                                   "api": "GraphedMVS4net.prefetch / run_prefetched: pinned images + cameras in (every "
                                          "step uploads one full request on a side stream while the previous forward "
                                          "runs) -> whole MVS4net.forward (FPN4, 4 x schedule / K1 / reg2d / tail) as one "
                                          "CUDA-graph replay -> depth + confidence read back to the host every step, "
                                          "%dx%d N=%d, fp32" % (net["h0"], net["w0"], args.views)}
        if network is not None:
            line["whole_network"] = network
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if gpu_ref is not None:
            line["gpu_reference"] = gpu_ref
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
