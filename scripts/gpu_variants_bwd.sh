#!/bin/bash
# A/B timing of library variants (variants/libvar_*.so) on the training-shape K1 forward+backward benchmark:
# the backward C-ABI call alone (incl. the zero-fill of grad_src) and the autograd forward+backward per stage
set -u
mkdir -p gpurun_out
fmt() { python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('  bwd graph ms: ' + '  '.join('s%d %.4f (%.1f%%)' % (s, d['stage%d_bwd_graph_ms' % s], 100 * d['stage%d_bwd_graph_frac_hbm_roofline' % s]) for s in (1, 2, 3, 4)))
        print('  bwd call ms: ' + '  '.join('s%d %.4f (%.1f%%)' % (s, d['stage%d_bwd_call_ms' % s], 100 * d['stage%d_bwd_frac_hbm_roofline' % s]) for s in (1, 2, 3, 4))
              + '   fwd+bwd: ' + ' '.join('%.3f' % d['stage%d_fwd_bwd_ms' % s] for s in (1, 2, 3, 4)))
"; }
python scripts/bench_extra.py --which train --iters 20 > /dev/null 2>&1   # warm the box
echo "== default"; python scripts/bench_extra.py --which train --iters 100 2>&1 | fmt
for lib in variants/libvar_*.so; do
  echo "== $lib"; MVSTER_B200_LIB=$lib python scripts/bench_extra.py --which train --iters 100 2>&1 | fmt
done
echo "== default again"; python scripts/bench_extra.py --which train --iters 100 2>&1 | fmt
