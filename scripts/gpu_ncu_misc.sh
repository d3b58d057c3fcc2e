#!/bin/bash
# ncu --set full of the backward and filter kernels (plain run of the same command first)
set -u
mkdir -p gpurun_out
CMD="python scripts/bench_extra.py --which train,filter --iters 4 --cpu-filter-pairs 0"
$CMD > gpurun_out/misc_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"epi_bwd|geo_filter" -c 6 -f -o gpurun_out/misc $CMD > gpurun_out/ncu_misc.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/misc_plain.log
