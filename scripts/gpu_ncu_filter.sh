#!/bin/bash
set -u
mkdir -p gpurun_out
CMD="python scripts/bench_extra.py --which filter --iters 4 --cpu-filter-pairs 0"
$CMD > gpurun_out/filter_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"geo_filter" -c 2 -f -o gpurun_out/filter $CMD > gpurun_out/ncu_filter.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/filter_plain.log
