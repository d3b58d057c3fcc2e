#!/bin/bash
# per-launch durations of one cascade step (ncu, cold caches, serialised) for the default build
set -u
mkdir -p gpurun_out
CMD="python scripts/bench_k1.py --iters 2 --tag ncu"
$CMD > gpurun_out/launchlist_plain.log 2>&1 || { echo plain run failed; tail -5 gpurun_out/launchlist_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,launch__registers_per_thread,launch__grid_size,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none -k regex:epi_ -c 24 --csv --log-file gpurun_out/launchlist.csv $CMD > gpurun_out/launchlist.log 2>&1
echo "ncu exit $?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launchlist.csv')) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
agg={}
for r in rows[1:]:
    agg.setdefault((r[ix['ID']],r[ix['Kernel Name']][:60]),{})[r[ix['Metric Name']]]=r[ix['Metric Value']]
for (i,k),m in list(agg.items())[-12:]:
    print(i,k,{a.split('.')[0][-28:]:b for a,b in m.items()})
PY
