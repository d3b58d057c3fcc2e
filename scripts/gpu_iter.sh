#!/bin/bash
# one iteration on the GPU box: network tests, then the bench line (no CPU baseline)
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_network.py -m gpu -x -q > gpurun_out/pytest_network.log 2>&1; echo "network exit $?"; tail -2 gpurun_out/pytest_network.log
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit $?"; tail -2 gpurun_out/bench.err
