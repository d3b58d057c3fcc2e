#!/bin/bash
# ncu --set full of the schedule / tail kernels of one cascade step
set -u
mkdir -p gpurun_out
CMD="python scripts/bench_k1.py --iters 1 --tag ncu"
ncu --set full --clock-control none --import-source on -k regex:"schedule_inverse|tail_kernel" -s 18 -c 6 -f -o gpurun_out/sched $CMD > gpurun_out/ncu_sched.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_sched.log
