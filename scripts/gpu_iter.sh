#!/bin/bash
# one iteration on the GPU box: network tests, then the FPN A/B
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_network.py -m gpu -x -q > gpurun_out/pytest_network.log 2>&1; echo "network exit $?"; tail -5 gpurun_out/pytest_network.log
timeout 600 python scripts/bench_fpn.py > gpurun_out/bench_fpn.log 2> gpurun_out/bench_fpn.err; echo "bench_fpn exit $?"; tail -1 gpurun_out/bench_fpn.log; tail -3 gpurun_out/bench_fpn.err
