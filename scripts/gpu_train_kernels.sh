#!/bin/bash
# training-mode kernels: timings against cuDNN, then ncu --set full of one launch of each
set -u
mkdir -p gpurun_out
python scripts/bench_train_kernels.py > gpurun_out/train_kernels.jsonl 2> gpurun_out/train_kernels.err; echo "bench exit $?"; tail -3 gpurun_out/train_kernels.err
ncu --set full --clock-control none --import-source on -k regex:"bn_fwd_stats|bn_fwd_apply|bn_bwd_stats|bn_bwd_apply|wgrad3d_kernel" -c 18 -f -o gpurun_out/train_kernels python scripts/bench_train_kernels.py --once > gpurun_out/ncu_train_kernels.log 2>&1; echo "ncu exit $?"
