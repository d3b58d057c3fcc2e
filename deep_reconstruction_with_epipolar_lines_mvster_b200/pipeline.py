"""``CascadePlan``: the 4-stage hot path of ``MVS4net.forward`` (reference models/MVS4Net.py:99-136) with every
buffer pre-allocated and every launch issued straight through the C ABI.

Per stage it runs, on one stream and with no host synchronisation:

    homography composition (all stages in ONE launch, up front)        -> rt        [B,N-1,12]
    hypothesis schedule (init_inverse_range / schedule_inverse_range)  -> depth_hypo [B,D,H,W]
    fused warp + correlation + epipolar attention + aggregation (K1)    -> cor_feats [B,G,D,H,W]
    regnet(cor_feats) -> logits            (caller-supplied callable; OUT OF SCOPE of this library - cuDNN in the
                                            reference - so benchmarks pass pre-computed logits)
    fused tail (K2a)                                                    -> depth, confidence, attn, inverse range

The plan is what ``bench.py`` times: ``run()`` with features resident in HBM, ``run_from_host()`` with pinned host
features copied in (and depth/confidence copied out) inside the call.
"""
from __future__ import annotations

import ctypes
from typing import Callable, Dict, List, Optional, Sequence

import torch

from . import _lib, ops
from .synthetic import STAGE_CHANNELS, STAGE_GROUPS, STAGE_NDEPTHS, STAGE_SPLIT_ITV, stage_shape


class CascadePlan:
    def __init__(self, batch: int, nviews: int, h0: int, w0: int, device="cuda", feature_dtype=torch.float32,
                 channels: Sequence[int] = STAGE_CHANNELS, groups: Sequence[int] = STAGE_GROUPS,
                 ndepths: Sequence[int] = STAGE_NDEPTHS, split_itv: Sequence[float] = STAGE_SPLIT_ITV,
                 attn_temp: float = 2.0):
        if not torch.cuda.is_available():
            raise RuntimeError("CascadePlan needs a CUDA device: the MVSTER B200 path has no CPU fallback")
        self.lib = _lib.load()
        self.B, self.N, self.h0, self.w0 = batch, nviews, h0, w0
        self.device = torch.device(device)
        self.feature_dtype = feature_dtype
        self.dtype_code = ops.BF16 if feature_dtype == torch.bfloat16 else ops.F32
        self.channels, self.groups, self.ndepths = list(channels), list(groups), list(ndepths)
        self.split_itv = list(split_itv)
        self.attn_temp = float(attn_temp)
        self.nstage = len(self.channels)
        dev = self.device
        f32 = torch.float32
        self.shapes = [stage_shape(h0, w0, s + (4 - self.nstage)) for s in range(self.nstage)]
        # device-resident inputs (NHWC features per stage/view, projections, depth range, stand-in logits)
        self.features: List[List[torch.Tensor]] = [
            [torch.empty((batch, h, w, c), device=dev, dtype=feature_dtype) for _ in range(nviews)]
            for (h, w), c in zip(self.shapes, self.channels)]
        self.proj_all = torch.zeros((self.nstage, batch, nviews, 2, 4, 4), device=dev, dtype=f32)  # one compose launch
        self.proj = [self.proj_all[s] for s in range(self.nstage)]
        self.depth_values = torch.empty((batch, 2), device=dev, dtype=f32)
        self.logits = [torch.zeros((batch, d, h, w), device=dev, dtype=f32)
                       for (h, w), d in zip(self.shapes, self.ndepths)]
        # outputs / intermediates
        self.rt_all = torch.empty((self.nstage, batch, nviews - 1, 12), device=dev, dtype=f32)
        self.rt = [self.rt_all[s] for s in range(self.nstage)]
        self.hypo = [torch.empty((batch, d, h, w), device=dev, dtype=f32) for (h, w), d in zip(self.shapes, self.ndepths)]
        self.volume = [torch.empty((batch, g, d, h, w), device=dev, dtype=f32)
                       for (h, w), g, d in zip(self.shapes, self.groups, self.ndepths)]
        self.attn = [torch.empty_like(x) for x in self.logits]
        mk = lambda: [torch.empty((batch, h, w), device=dev, dtype=f32) for (h, w) in self.shapes]
        self.depth, self.conf, self.inv_min, self.inv_max = mk(), mk(), mk(), mk()
        self._src_ptrs = [ops._ptr_array(fs[1:]) for fs in self.features]
        self.regnet: Optional[Callable[[int, torch.Tensor], torch.Tensor]] = None
        self.stage_events = None  # optional [(start, end)] CUDA events around one stage's K1 launch
        self.graph = None         # set by capture()

    # ---- sizes for the roofline (algorithmic bytes of K1 forward) ---------------------------------
    def k1_bytes(self, stage: int) -> int:
        (h, w), c, g, d = self.shapes[stage], self.channels[stage], self.groups[stage], self.ndepths[stage]
        s = 2 if self.feature_dtype == torch.bfloat16 else 4
        return self.B * h * w * (self.N * c * s + d * 4 + g * d * 4)

    def k1_fmas(self, stage: int) -> int:
        (h, w), c, g, d = self.shapes[stage], self.channels[stage], self.groups[stage], self.ndepths[stage]
        return self.B * (self.N - 1) * d * h * w * (5 * c + g)

    def feature_bytes(self) -> int:
        return sum(f.numel() * f.element_size() for fs in self.features for f in fs)

    # ---- the hot path ------------------------------------------------------------------------------------------------
    def run(self, time_stage: Optional[int] = None, stage_ready: Optional[Sequence] = None):
        """Issue all stages on the current stream.  If ``time_stage`` is given and ``self.stage_events`` holds a
        (start, end) event pair, the pair brackets that stage's K1 launch.  ``stage_ready[s]`` (optional CUDA events)
        are waited on before stage ``s`` starts (its inputs arrive on another stream)."""
        lib, check, P = self.lib, _lib.check, ops._ptr
        st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        B, N = self.B, self.N
        for s in range(self.nstage):
            (h, w), c, g, d = self.shapes[s], self.channels[s], self.groups[s], self.ndepths[s]
            if stage_ready is not None:
                torch.cuda.current_stream(self.device).wait_event(stage_ready[s])
            if s == 0:
                # every stage's projections arrive before stage 0's features (run_from_host copies them first)
                check(lib.mvster_compose_homographies(P(self.proj_all), P(self.rt_all), self.nstage * B, N, st))
                check(lib.mvster_init_inverse_range(P(self.depth_values), 2, P(self.hypo[0]), B, d, h, w, st))
            else:
                check(lib.mvster_schedule_inverse_range(P(self.inv_min[s - 1]), P(self.inv_max[s - 1]), P(self.hypo[s]),
                                                        B, d, h, w, st))
            timed = time_stage == s and self.stage_events is not None
            if timed:
                self.stage_events[0].record()
            check(lib.mvster_epi_fwd(P(self.features[s][0]), self._src_ptrs[s], P(self.rt[s]), P(self.hypo[s]),
                                     P(self.volume[s]), None, None, B, N - 1, c, g, d, h, w, h, w, self.attn_temp,
                                     self.dtype_code, st))
            if timed:
                self.stage_events[1].record()
            logits = self.logits[s] if self.regnet is None else self.regnet(s, self.volume[s])
            check(lib.mvster_tail(P(logits), P(self.hypo[s]), float(self.split_itv[s]), ops.DEPTH_ARGMAX,
                                  P(self.attn[s]), P(self.depth[s]), P(self.conf[s]), P(self.inv_min[s]),
                                  P(self.inv_max[s]), B, d, h, w, st))
        return self.depth[-1], self.conf[-1]

    # compose (all stages), then per stage: hypothesis schedule, K1, tail.  (Building the hypotheses inside K1 instead
    # was measured and rejected: the stage-4 K1 kernel is issue-bound, the extra prologue cost 0.10 ms against the
    # 0.065 ms schedule launch it removed.)
    LAUNCHES_PER_RUN = 13

    # ---- CUDA graph: the launches of one cascade as a single graph launch --------------------------------------
    def capture(self):
        """Capture ``run()`` into a CUDA graph (all buffers are pre-allocated and the kernels take raw pointers, so
        the captured launches stay valid).  Only possible without a Python ``regnet`` callback."""
        if self.regnet is not None:
            raise RuntimeError("CascadePlan.capture(): a Python regnet callback cannot be captured")
        self.run()  # warm-up outside capture: function attributes, TMA entry point, lazy module load
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.run()
        return self

    def replay(self):
        if self.graph is None:
            raise RuntimeError("CascadePlan.replay(): call capture() first")
        if self.regnet is not None:
            raise RuntimeError("CascadePlan.replay(): the captured graph uses the resident stand-in logits; a regnet "
                               "callback assigned after capture() would be ignored - use run() instead")
        self.graph.replay()
        return self.depth[-1], self.conf[-1]

    # ---- end to end from host memory -----------------------------------------------------------------------------------
    def make_host_buffers(self):
        """Pinned host mirrors of every per-step input, and pinned outputs for the final depth / confidence."""
        pin = lambda t: torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        self.h_features = [[pin(f) for f in fs] for fs in self.features]
        self.h_proj = [pin(p) for p in self.proj]
        self.h_depth_values = pin(self.depth_values)
        self.h_depth = pin(self.depth[-1])
        self.h_conf = pin(self.conf[-1])
        return self

    def h2d_bytes(self) -> int:
        n = self.feature_bytes() + sum(p.numel() * 4 for p in self.proj) + self.depth_values.numel() * 4
        return int(n)

    def d2h_bytes(self) -> int:
        return int(self.depth[-1].numel() * 4 + self.conf[-1].numel() * 4)

    def run_from_host(self):
        """Copy this step's inputs from pinned host memory, run the cascade, copy depth + confidence back; returns
        after the results are on the host (the call a user with host-side data makes).

        The copies of stage k+1's features are issued on a side stream while stage k computes (the link, not the GPU,
        is the bottleneck: 2.4 GB in per step against ~2 ms of kernels), so only the last stage's kernels are exposed."""
        cur = torch.cuda.current_stream(self.device)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(self.device)
            self._stage_ready = [torch.cuda.Event() for _ in range(self.nstage)]
        cs = self._copy_stream
        cs.wait_stream(cur)  # the previous step's kernels are done reading the device buffers
        with torch.cuda.stream(cs):
            self.depth_values.copy_(self.h_depth_values, non_blocking=True)
            for s in range(self.nstage):   # all projections first: stage 0 composes the homographies of every stage
                self.proj[s].copy_(self.h_proj[s], non_blocking=True)
            for s in range(self.nstage):
                for a, b in zip(self.h_features[s], self.features[s]):
                    b.copy_(a, non_blocking=True)
                self._stage_ready[s].record(cs)
        depth, conf = self.run(stage_ready=self._stage_ready)
        self.h_depth.copy_(depth, non_blocking=True)
        self.h_conf.copy_(conf, non_blocking=True)
        cur.synchronize()
        return self.h_depth, self.h_conf
