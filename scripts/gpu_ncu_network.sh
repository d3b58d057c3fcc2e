#!/bin/bash
# ncu of this library's kernels inside one MVS4net.forward (832x1152, N=5, one scene): the ten metrics of
# profiles/<tag>_network_kernels_ncu.md only (a --set full pass over ~80 launches costs ten GPU-minutes); only the
# raw-metric CSV travels back.
set -u
mkdir -p gpurun_out /tmp/ncu
M=gpu__time_duration.sum,launch__grid_size,launch__registers_per_thread,smsp__issue_active.avg.pct_of_peak_sustained_active
M=$M,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
M=$M,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_uniform.sum
python scripts/profile_network.py --once > gpurun_out/network_once.log 2>&1 &&
ncu --metrics $M --clock-control none -k regex:"fpn_topdown|fpn_lin|fpn_proj|conv5s2|smallconv|midconv|regtail|epi_fwd|tail_kernel|schedule|nchw" -c 160 -f -o /tmp/ncu/network_kernels python scripts/profile_network.py --once > gpurun_out/ncu_network.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_network.log
python scripts/profile_network.py > gpurun_out/network_torch_profiler.txt 2>&1
ncu -i /tmp/ncu/network_kernels.ncu-rep --page raw --csv > gpurun_out/network_kernels_raw.csv 2>/dev/null
ls -la gpurun_out/network_kernels_raw.csv
