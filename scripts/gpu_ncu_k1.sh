#!/bin/bash
# ncu --set full on the four K1 forward kernels of one cascade step (plain run first, same command line).
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu.log
BCMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-network --e2e-steps 1"
$BCMD > gpurun_out/bench_small.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:epi_fwd -s 4 -c 4 -f -o gpurun_out/k1_fwd $BCMD > gpurun_out/ncu_k1.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_k1.log
