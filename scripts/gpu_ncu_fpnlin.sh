#!/bin/bash
# ncu --set full of the linearised FPN top-down kernels inside one FPN4 pass (832x1152, 5 views)
set -u
mkdir -p gpurun_out
CMD="python scripts/bench_fpn.py --iters 3"
$CMD > gpurun_out/bench_fpn_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"fpn_lin|fpn_proj" -c 3 -f -o gpurun_out/fpn_lin $CMD > gpurun_out/ncu_fpnlin.log 2>&1
echo "ncu exit $?"; tail -1 gpurun_out/bench_fpn_plain.log | cut -c1-300
