#!/usr/bin/env python
"""gpurun_out/network_kernels_raw.csv (ncu --set full --page raw of this library's kernels inside one MVS4net.forward,
scripts/gpu_ncu_network.sh) -> profiles/<tag>_network_kernels_ncu.md"""
import csv, os, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(os.path.join(root, "gpurun_out", "network_kernels_raw.csv"))))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
cols = [("gpu__time_duration.sum", "time"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe %"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "L1/smem pipe %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %")]
cols = [c for c in cols if c[0] in ix]
out = os.path.join(root, "profiles", "%s_network_kernels_ncu.md" % tag)
with open(out, "w") as f:
    f.write("# ncu --set full --clock-control none: this library's kernels inside one `MVS4net.forward` (%s)\n\n" % tag)
    f.write("832x1152, N=5 views, one scene, fp32 (`scripts/gpu_ncu_network.sh`); one row per distinct (kernel, grid), in "
            "launch order: FPN4 encoder, top-down, then reg2d + regulariser tail of stages 1-4.\n\n")
    f.write("| kernel | " + " | ".join(c[1] for c in cols) + " |\n|---|" + "---|" * len(cols) + "\n")
    seen = set()
    for d in data:
        name = d[ix["Kernel Name"]].replace("mvster::", "").replace("void ", "").split("(")[0]
        key = (name, d[ix["launch__grid_size"]])
        if key in seen:
            continue
        seen.add(key)
        vals = []
        for m, _ in cols:
            v, u = d[ix[m]], units[ix[m]]
            try:
                v = "%.3g" % float(v.replace(",", ""))
            except ValueError:
                pass
            vals.append(v + ("" if u in ("", "%", "register/thread") else " " + u))
        f.write("| `%s` | %s |\n" % (name, " | ".join(vals)))
print("wrote", out)
