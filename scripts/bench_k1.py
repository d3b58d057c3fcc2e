#!/usr/bin/env python
"""Per-stage K1 forward timing on the bench.py workload (noisy cascade depth) and on smooth hypotheses.

    [MVSTER_B200_LIB=variants/libvar_X.so] [MVSTER_FINE_V1=1] python scripts/bench_k1.py [--iters 30] [--tag NAME]
Prints one JSON line: {"tag":..., "cascade_ms":[s1..s4], "smooth_ms":[...], "frac":[...]}.
"""
import argparse, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import bench  # noqa: E402
from deep_reconstruction_with_epipolar_lines_mvster_b200.pipeline import CascadePlan  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--tag", default=os.environ.get("MVSTER_B200_LIB", "default"))
    ap.add_argument("--scenes", type=int, default=8)
    ap.add_argument("--smooth", action="store_true")
    ap.add_argument("--dtype", default="fp32")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    plan = CascadePlan(args.scenes, 5, 864, 1152, device=dev,
                       feature_dtype=torch.bfloat16 if args.dtype == "bf16" else torch.float32)
    bench.fill_plan(plan, 0)
    for _ in range(3):
        plan.run()
    out = {"tag": args.tag, "cascade_ms": [], "frac": []}
    for s in range(plan.nstage):
        pairs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.iters)]
        for pr in pairs:
            plan.stage_events = pr
            plan.run(time_stage=s)
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in pairs) / len(pairs)
        out["cascade_ms"].append(round(ms, 4))
        out["frac"].append(round(plan.k1_bytes(s) / (ms * 1e-3) / 1e9 / bench.peaks()[0], 4))
    plan.stage_events = None
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.iters):
        plan.run()
    b.record()
    torch.cuda.synchronize()
    out["step_ms"] = round(a.elapsed_time(b) / args.iters, 4)
    if args.smooth:
        import bench_extra as be
        from deep_reconstruction_with_epipolar_lines_mvster_b200 import ops
        out["smooth_ms"] = []
        for stage in range(4):
            feats, proj, hypo, g, d, c, h, w = be.stage_inputs(8, 5, 864, 1152, stage, dev)
            nhwc = [ops.to_nhwc(f) for f in feats]
            rt = ops.compose_homographies(proj)
            out["smooth_ms"].append(round(be.timed(lambda: ops.epi_fwd(nhwc[0], nhwc[1:], rt, hypo, g, 2.0), args.iters), 4))
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
