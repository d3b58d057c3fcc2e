#!/usr/bin/env python
"""Dynamic instruction count and stall samples per SOURCE LINE of one kernel launch in an .ncu-rep.

ncu's source page only carries SASS rows here (the CUDA view prints no metrics without the GUI), so the rows are
joined by instruction offset with `nvdisasm -gi` of the object file that holds the kernel (needs -lineinfo).

    python scripts/ncu_lines.py REPORT.ncu-rep LAUNCH_SKIP OBJECT.o MANGLED_SUBSTRING [--px NPIX] [--inner] [--top N]

--inner groups by the innermost (inlined callee) line instead of the line inside the kernel body.
"""
import argparse, collections, csv, os, re, subprocess, sys, tempfile

ap = argparse.ArgumentParser()
ap.add_argument("rep"); ap.add_argument("skip"); ap.add_argument("obj"); ap.add_argument("sym")
ap.add_argument("--px", type=float, default=None)
ap.add_argument("--inner", action="store_true")
ap.add_argument("--top", type=int, default=60)
a = ap.parse_args()

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(a.obj)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.lstrip().startswith(".section") and ".text." in l and a.sym in l)
line_of = {}   # offset -> (outer file:line, inner file:line)
pend = []
for l in dis[start + 1:]:
    s = l.strip()
    if s.startswith(".section"):
        break
    m = re.match(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', s)
    if m:
        pend.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    m = re.match(r'/\*([0-9a-f]{4,})\*/', s)
    if m:
        off = int(m.group(1), 16)
        if pend:
            cur = (pend[-1], pend[0])
            pend = []
        line_of[off] = cur

raw = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", a.skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
print(rows[0][1][:110])
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data, seen = [], set()
for r in rows[2:]:
    if r and r[0].startswith("0x") and r[0] not in seen:
        seen.add(r[0]); data.append(r)
base = min(int(r[0], 16) for r in data)
st = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
agg = collections.defaultdict(lambda: [0, 0, collections.Counter(), collections.Counter()])
tot_i = tot_s = 0
for r in data:
    off = int(r[0], 16) - base
    outer, inner = line_of.get(off, (("?", 0), ("?", 0)))
    key = inner if a.inner else outer
    n = int(r[ix["Instructions Executed"]]); s = int(r[ix["# Samples"]])
    e = agg[key]; e[0] += n; e[1] += s
    src = r[ix["Source"]].split(); op = src[1] if src[0].startswith("@") else src[0]
    e[2][op.split(".")[0]] += n
    for c in st:
        e[3][c[6:]] += int(r[ix[c]])
    tot_i += n; tot_s += s
scale = 32.0 / a.px if a.px else None
print("total warp instr %d  samples %d%s" % (tot_i, tot_s, ("  thread-instr/px %.0f" % (tot_i * scale)) if scale else ""))
for key, e in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1]))[:10000]:
    if e[0] < tot_i * 0.002 and e[1] < tot_s * 0.004:
        continue
    print("%-18s %4d  instr %5.1f%% %s samples %5.1f%%  %s | %s" % (
        key[0][:18], key[1], 100.0 * e[0] / tot_i, ("(%6.1f/px)" % (e[0] * scale)) if scale else "",
        100.0 * e[1] / max(tot_s, 1), " ".join("%s:%d" % kv for kv in e[2].most_common(5)),
        " ".join("%s=%d%%" % (k, 100 * v / max(e[1], 1)) for k, v in e[3].most_common(3))))
