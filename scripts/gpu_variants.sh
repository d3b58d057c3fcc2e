#!/bin/bash
# A/B timing of library variants (variants/libvar_*.so) on the per-stage K1 benchmark, plus the default build
set -u
mkdir -p gpurun_out
echo "== default"; python scripts/bench_extra.py --which stages --iters 30 2>&1 | python scripts/_fmt_stages.py
for lib in variants/libvar_*.so; do
  echo "== $lib"; MVSTER_B200_LIB=$lib python scripts/bench_extra.py --which stages --iters 30 2>&1 | python scripts/_fmt_stages.py
done
echo "== default, MVSTER_NO_LINE=1"; MVSTER_NO_LINE=1 python scripts/bench_extra.py --which stages --iters 30 2>&1 | python scripts/_fmt_stages.py
