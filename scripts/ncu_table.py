#!/usr/bin/env python
"""One markdown table (one column per distinct (kernel, grid)) of the key raw metrics in an .ncu-rep.

    python scripts/ncu_table.py gpurun_out/x.ncu-rep profiles/r01_x.md "title" "command line that was profiled"
"""
import csv, subprocess, sys

rep, out, title, cmd = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ki, gi = hdr.index("Kernel Name"), hdr.index("launch__grid_size")
seen, cols = set(), []
for d in data:
    if (d[ki], d[gi]) not in seen:
        seen.add((d[ki], d[gi]))
        cols.append(d)
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def short(n):
    return n.replace("void ", "").replace("mvster::", "").split("(mvster")[0].replace("(int)", "").replace("(bool)", "")


with open(out, "w") as f:
    f.write("# %s\n\n`ncu --set full --clock-control none` of `%s`; one column per distinct (kernel, grid).\n\n" % (title, cmd))
    f.write("| metric | unit | " + " | ".join("`%s`" % short(d[ki]) for d in cols) + " |\n")
    f.write("|---|---|" + "---|" * len(cols) + "\n")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            f.write("| %s | %s | %s |\n" % (w, units[i], " | ".join(d[i][:12] for d in cols)))
print(open(out).read()[:1500])
