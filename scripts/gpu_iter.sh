#!/bin/bash
# one iteration on the GPU box: network tests, then the FPN A/B
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_network.py -m gpu -x -q > gpurun_out/pytest_network.log 2>&1; echo "network exit $?"; tail -5 gpurun_out/pytest_network.log
timeout 600 python scripts/bench_fpn.py > gpurun_out/bench_fpn.log 2> gpurun_out/bench_fpn.err; echo "bench_fpn exit $?"; tail -1 gpurun_out/bench_fpn.log; tail -3 gpurun_out/bench_fpn.err
if [ "${PROF:-0}" = "1" ]; then
python - > gpurun_out/fpn_prof.txt 2>&1 <<'PY'
import sys, torch
sys.path.insert(0, '.')
import deep_reconstruction_with_epipolar_lines_mvster_b200 as mv
from deep_reconstruction_with_epipolar_lines_mvster_b200 import synthetic as syn
from scripts.bench_extra import NET_CFG
dev = torch.device('cuda:0')
model = mv.MVS4net(**NET_CFG).eval(); model.load_state_dict(syn.fill_state_dict(model.state_dict(), seed=7)); model = model.to(dev)
imgs = [torch.rand((1, 3, 832, 1152), device=dev) for _ in range(5)]
with torch.no_grad():
    for _ in range(3): model.extract_features(imgs)
    torch.cuda.synchronize()
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        for _ in range(5): model.extract_features(imgs)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
PY
tail -32 gpurun_out/fpn_prof.txt | cut -c1-90,150-230
fi
