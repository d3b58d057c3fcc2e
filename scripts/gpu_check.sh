#!/bin/bash
# One gpurun call: GPU parity tests, smoke, a short bench, then (only if the same bench command exited 0) the ncu
# launch list.  Everything the builder wants back goes to gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
python bench.py --steps 50 --warmup 5 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"
tail -2 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
BCMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-network --e2e-steps 1"
$BCMD > gpurun_out/bench_small.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $BCMD > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit $?"
