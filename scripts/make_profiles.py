#!/usr/bin/env python
"""Turn gpurun_out/launches.csv and gpurun_out/k1_fwd.ncu-rep into the committed summaries under profiles/.

    python scripts/make_profiles.py r01
"""
import collections, csv, json, os, subprocess, sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(root, "profiles")
os.makedirs(out, exist_ok=True)

# ---- launch list: per-kernel mean duration and share of one bench step ------------------------------------------
lines = [l for l in open(os.path.join(root, "gpurun_out", "launches.csv")) if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    if row["Metric Unit"] == "us":
        v *= 1e3
    elif row["Metric Unit"] == "ms":
        v *= 1e6
    name = row["Kernel Name"]
    if "mvster::" not in name:
        continue
    key = (name.split("(")[0].replace("void ", ""), row.get("Grid Size", ""))
    agg.setdefault(key, []).append(v)
tot = sum(sum(v) / len(v) for v in agg.values())
with open(os.path.join(out, "%s_launches.md" % tag), "w") as f:
    f.write("# ncu launch list of `python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-network --e2e-steps 1` (%s)\n\n" % tag)
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none`; cold-cache, serialised launches: compare SHARES.\n")
    f.write("One cascade step = 13 launches of this library: compose (all stages), then 4 stages x schedule/K1/tail.\n\n")
    f.write("| kernel | grid | launches seen | mean µs | share of one step |\n|---|---|---|---|---|\n")
    for (name, grid), v in agg.items():
        m = sum(v) / len(v)
        f.write("| `%s` | %s | %d | %.1f | %.1f %% |\n" % (name, grid, len(v), m / 1e3, 100 * m / tot))
    f.write("\nsum of per-kernel means (one step): %.1f µs\n" % (tot / 1e3))

# ---- full capture: key raw metrics per K1 kernel ---------------------------------------------------------------------
rep = os.path.join(root, "gpurun_out", "k1_fwd.ncu-rep")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
ki = hdr.index("Kernel Name")
with open(os.path.join(out, "%s_k1_fwd_ncu_full.md" % tag), "w") as f:
    f.write("# ncu --set full --clock-control none, K1 forward kernels of one cascade step (%s)\n\n" % tag)
    f.write("B=8 scenes, N=5 views, 864x1152, fp32; one launch per stage (stage 1..4 left to right).\n\n")
    f.write("| metric | unit | " + " | ".join("`%s`" % d[ki].split("(")[0].replace("void ", "")[:44] for d in data) + " |\n")
    f.write("|---|---|" + "---|" * len(data) + "\n")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            f.write("| %s | %s | %s |\n" % (w, units[i], " | ".join(d[i] for d in data)))
# traffic of the dominant kernel (stage-4 K1 forward) for bench.py's roofline.traffic
def num(d, name):
    i = hdr.index(name)
    v = float(d[i].replace(",", ""))
    u = units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
dom = data[-1]
json.dump({"kernel": dom[ki], "dram_bytes_read": num(dom, "dram__bytes_read.sum"),
           "dram_bytes_write": num(dom, "dram__bytes_write.sum"),
           "source": "profiles/%s_k1_fwd_ncu_full.md" % tag},
          open(os.path.join(out, "k1_stage4_traffic.json"), "w"), indent=1)
print("wrote", os.listdir(out))
