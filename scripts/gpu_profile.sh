#!/bin/bash
# Profile capture for profiles/: launch list of the bench command + ncu --set full of the K1 forward kernels.
set -u
mkdir -p gpurun_out
BCMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-network --e2e-steps 1"
$BCMD > gpurun_out/bench_small.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $BCMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
$BCMD > gpurun_out/bench_small2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:epi_fwd -s 4 -c 4 -f -o gpurun_out/k1_fwd $BCMD > gpurun_out/ncu_k1.log 2>&1
echo "ncu full exit $?"
