#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
python -c "
import json
l=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','ms_per_step','gpu_launches')}, l['e2e']['value'], l['roofline']['kernel_ms'], l['roofline']['frac'])"
python scripts/bench_extra.py --which ${WHICH:-stages,binpick,train,filter,ref_gpu} > gpurun_out/extra.log 2> gpurun_out/extra.err; echo "extra exit $?"
cat gpurun_out/extra.log; tail -5 gpurun_out/extra.err
