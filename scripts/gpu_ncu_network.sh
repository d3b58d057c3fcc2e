#!/bin/bash
# ncu --set full of this library's kernels inside one MVS4net.forward (832x1152, N=5, one scene); only the raw-metric
# CSV travels back (the .ncu-rep of ~60 full captures exceeds the 64 MiB gpurun_out limit)
set -u
mkdir -p gpurun_out /tmp/ncu
python scripts/profile_network.py --once > gpurun_out/network_once.log 2>&1 &&
ncu --set full --clock-control none -k regex:"fpn_topdown|conv5s2|smallconv|regtail" -c 60 -f -o /tmp/ncu/network_kernels python scripts/profile_network.py --once > gpurun_out/ncu_network.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_network.log
ncu -i /tmp/ncu/network_kernels.ncu-rep --page raw --csv > gpurun_out/network_kernels_raw.csv 2>/dev/null
ls -la gpurun_out/network_kernels_raw.csv
