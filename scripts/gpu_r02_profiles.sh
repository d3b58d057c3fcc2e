#!/bin/bash
# r02 evidence: full bench line (both arms), ncu --set full of every kernel of one cascade step, the filter kernel
set -u
mkdir -p gpurun_out
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_r02.log 2> gpurun_out/bench_r02.err; echo "bench exit $?"
tail -c 3000 gpurun_out/bench_r02.log; tail -3 gpurun_out/bench_r02.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref_r02.log 2>&1; echo "ref arm exit $?"; tail -c 600 gpurun_out/bench_ref_r02.log
CMD="python scripts/bench_k1.py --iters 1 --tag ncu"
ncu --set full --clock-control none --import-source on -k regex:"epi_fwd|tail_kernel|schedule_inverse|init_inverse|compose" -s 39 -c 13 -f -o gpurun_out/step_r02 $CMD > gpurun_out/ncu_step.log 2>&1
echo "ncu step exit $?"; tail -2 gpurun_out/ncu_step.log
bash scripts/gpu_ncu_filter.sh
