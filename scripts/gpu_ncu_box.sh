#!/bin/bash
# ncu --set full on the box kernels (stage 3 + stage 4 K1 forward) of one cascade step; plain run first.
#   bash scripts/gpu_ncu_box.sh [OUTNAME [LIB]]      (LIB: a variants/libvar_*.so to profile instead of the default build)
set -u
mkdir -p gpurun_out
OUT=${1:-k1_box}
[ -n "${2:-}" ] && export MVSTER_B200_LIB=$2
CMD="python scripts/bench_k1.py --iters 2 --tag ncu"
$CMD > gpurun_out/ncu_box_plain.log 2>&1 || { echo plain run failed; tail -5 gpurun_out/ncu_box_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"epi_fwd_(box|stream)" -s 6 -c 2 -f -o gpurun_out/$OUT $CMD > gpurun_out/ncu_box.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_box.log
