#!/bin/bash
# K1 forward A/B on the bench workload: per-stage timings of the default build and every variants/libvar_*.so
# (PARITY=1: run the K1 parity tests with the default build first; ENVS="A=1 B=1": also time the default build under
# each of these environment switches; SMOOTH=1: also time the smooth-depth workload)
set -u
mkdir -p gpurun_out
if [ "${PARITY:-0}" = "1" ]; then
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q > gpurun_out/pytest_k1.log 2>&1; echo "pytest exit $?"; tail -8 gpurun_out/pytest_k1.log
fi
EXTRA=""; [ "${SMOOTH:-0}" = "1" ] && EXTRA="--smooth"
: > gpurun_out/k1_ab.log
timeout 300 python scripts/bench_k1.py $EXTRA --tag default >> gpurun_out/k1_ab.log 2>gpurun_out/k1_ab.err
for e in ${ENVS:-}; do
  env $e timeout 300 python scripts/bench_k1.py $EXTRA --tag $e >> gpurun_out/k1_ab.log 2>>gpurun_out/k1_ab.err
done
for lib in variants/libvar_*.so; do
  [ -f "$lib" ] || continue
  MVSTER_B200_LIB=$lib timeout 300 python scripts/bench_k1.py $EXTRA --tag $lib >> gpurun_out/k1_ab.log 2>>gpurun_out/k1_ab.err
done
cat gpurun_out/k1_ab.log; tail -5 gpurun_out/k1_ab.err
