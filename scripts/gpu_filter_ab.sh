#!/bin/bash
# K2b A/B: filter parity tests, then the 49x9-view scene timed for the pixels-per-thread variants (and variant builds
# under variants/) with checksums of the masks / averaged depths; kernel time from an ncu duration pass.
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "filter or fusion or geo" > gpurun_out/pytest_filter.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_filter.log
: > gpurun_out/filter_ab.jsonl
run() {  # name, env...
  name=$1; shift
  env "$@" python scripts/bench_extra.py --which filter --iters 40 --cpu-filter-pairs 0 2>/dev/null | grep filter_49 | sed "s/^{/{\"variant\": \"$name\", /" >> gpurun_out/filter_ab.jsonl
  env "$@" ncu --metrics gpu__time_duration.sum --clock-control none -k regex:geo_filter -c 3 --csv python scripts/bench_extra.py --which filter --iters 4 --cpu-filter-pairs 0 2>/dev/null | grep geo_filter | tail -1 | awk -F'","' -v n=$name '{print n, "ncu_us", $NF}' >> gpurun_out/filter_ncu_times.txt
}
: > gpurun_out/filter_ncu_times.txt
for ppt in 4 2 1; do run ppt$ppt MVSTER_FILTER_PPT=$ppt; done
for lib in variants/libvar_*.so; do [ -f "$lib" ] && run $(basename $lib .so) MVSTER_B200_LIB=$lib; done
python - <<'PY'
import json
for l in open('gpurun_out/filter_ab.jsonl'):
    d = json.loads(l); print(d['variant'], round(d['ms_per_scene'], 4), d['sha1_masks'], d['sha1_depth_avg'], d['geo_mask_mean'])
PY
cat gpurun_out/filter_ncu_times.txt
if [ "${NCU:-0}" = "1" ]; then
ncu --set full --clock-control none --import-source on -k regex:geo_filter -c 1 -f -o gpurun_out/filter_r02b python scripts/bench_extra.py --which filter --iters 4 --cpu-filter-pairs 0 > gpurun_out/ncu_filter.log 2>&1; echo "ncu exit $?"
fi
