#!/bin/bash
# A/B timing of library variants (variants/libvar_*.so) on the training-shape K1 forward+backward benchmark
set -u
mkdir -p gpurun_out
fmt() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('  ' + '  '.join('s%d fwd+bwd %.3f (bwd %.3f)' % (s, d['stage%d_fwd_bwd_ms' % s], d['stage%d_fwd_bwd_ms' % s] - d['stage%d_fwd_ms' % s]) for s in (1, 2, 3, 4)))
"; }
echo "== default"; python scripts/bench_extra.py --which train --iters 30 2>&1 | fmt
for lib in variants/libvar_*.so; do
  echo "== $lib"; MVSTER_B200_LIB=$lib python scripts/bench_extra.py --which train --iters 30 2>&1 | fmt
done
