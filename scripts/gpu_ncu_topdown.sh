#!/bin/bash
# ncu --set full of the three fpn_topdown launches of one MVS4net.forward (832x1152, N=5, one scene)
set -u
mkdir -p gpurun_out
python scripts/profile_network.py --once > gpurun_out/network_once.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:"fpn_topdown" -c 3 -f -o gpurun_out/topdown python scripts/profile_network.py --once > gpurun_out/ncu_topdown.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_topdown.log; ls -la gpurun_out/topdown.ncu-rep
