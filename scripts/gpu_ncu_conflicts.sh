#!/bin/bash
# shared-memory wavefronts / bank conflicts of the stage-4 K1 forward kernel on the cascade workload and on smooth hypotheses
set -u
mkdir -p gpurun_out
export MVSTER_NO_STREAM=1
ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,smsp__inst_executed_op_shared_ld.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum,gpu__time_duration.sum --clock-control none -k epi_fwd_box_kernel --csv --log-file gpurun_out/conflicts.csv python scripts/bench_k1.py --iters 1 --smooth --tag ncu > gpurun_out/conflicts.log 2>&1
echo "ncu exit $?"
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/conflicts.csv')) if len(r)>10]
ix={h:i for i,h in enumerate(rows[0])}
d={}
for r in rows[1:]:
    d.setdefault(int(r[ix['ID']]),{})[r[ix['Metric Name']]]=float(r[ix['Metric Value']].replace(',',''))
for k,v in sorted(d.items()):
    ins=v['smsp__inst_executed_op_shared_ld.sum']
    print(k, "us %.1f"%(v['gpu__time_duration.sum']/1e3), "LDS %d"%ins, "wavefronts/LDS %.3f"%(v['l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum']/ins), "conflicts/LDS %.3f"%(v['l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum']/ins))
PY
