#!/bin/bash
# one iteration on the GPU box: training tests, then the training-step A/B (fused BatchNorm on / off)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_network.py -m gpu -x -q > gpurun_out/pytest_train.log 2>&1; echo "train tests exit $?"; tail -5 gpurun_out/pytest_train.log
timeout 600 python scripts/bench_extra.py --which train_step --iters 20 > gpurun_out/train_step.log 2> gpurun_out/train_step.err; echo "train_step exit $?"; tail -1 gpurun_out/train_step.log; tail -3 gpurun_out/train_step.err
if [ "${PROF:-0}" = "1" ]; then python scripts/profile_train_step.py > gpurun_out/train_step_torch_profiler.txt 2>&1; head -30 gpurun_out/train_step_torch_profiler.txt | cut -c1-70,100-190; fi
