#!/bin/bash
# one iteration on the GPU box: K1 parity tests (incl. the option variants), then the K2b A/B
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_parity.log 2>&1; echo "parity exit $?"; tail -5 gpurun_out/pytest_parity.log
bash scripts/gpu_filter_ab.sh
