#!/bin/bash
# r02 final evidence in one call: GPU tests + smoke + bench + launch list (gpu_check.sh), the reference arm, the secondary
# benches, and the ncu captures of the kernels changed last (K2b, linearised FPN top-down, network kernel table).
set -u
mkdir -p gpurun_out
bash scripts/gpu_check.sh
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref.log 2>&1; echo "ref arm exit $?"; tail -c 400 gpurun_out/bench_ref.log
python scripts/bench_extra.py --which filter,network,scene,config0 --iters 20 --cpu-filter-pairs 0 > gpurun_out/extra_final.log 2> gpurun_out/extra_final.err; echo "extra exit $?"
python scripts/bench_fpn.py > gpurun_out/bench_fpn.log 2>&1
bash scripts/gpu_ncu_filter.sh
bash scripts/gpu_ncu_fpnlin.sh
bash scripts/gpu_ncu_network.sh
