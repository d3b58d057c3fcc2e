#!/bin/bash
# quick GPU iteration: parity tests + bench (no CPU baseline) [+ optional ncu of K1 with NCU=1]
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    l=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print({k:l[k] for k in ('value','ms_per_step','gpu_launches')}, l['e2e']['value'], l['roofline']['kernel_ms'], l['roofline']['frac'], l['clocks'])
except Exception as e: print('bench parse failed', e); print(open('gpurun_out/bench.err').read()[-2000:])
PY
if [ "${COMPARE:-0}" = "1" ]; then
MVSTER_NO_TMA=1 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_notma.log 2> gpurun_out/bench_notma.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_notma.log').read().strip().splitlines()[-1])
print('NO_TMA', {k:l[k] for k in ('value','ms_per_step')}, l['roofline']['kernel_ms'], l['roofline']['frac'])
PY
fi
if [ "${NCU:-0}" = "1" ]; then
BCMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-network --e2e-steps 1"
$BCMD > gpurun_out/bench_small.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:epi_fwd -s 4 -c 4 -f -o gpurun_out/k1_fwd $BCMD > gpurun_out/ncu_k1.log 2>&1
echo "ncu exit $?"
fi
