#!/usr/bin/env python
"""Stall-reason / instruction-mix summary of one kernel launch in an .ncu-rep (source page)."""
import csv, collections, subprocess, sys
rep, skip = sys.argv[1], sys.argv[2]
npx = float(sys.argv[3]) if len(sys.argv) > 3 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
print(rows[0][1][:100])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data, seen = [], set()
for r in rows[2:]:
    if r and r[0].startswith('0x') and r[0] not in seen:
        seen.add(r[0]); data.append(r)
tot = sum(int(r[ix['# Samples']]) for r in data)
st = [c for c in hdr if c.startswith('stall_') and 'Not Issued' not in c]
agg = {c: sum(int(r[ix[c]]) for r in data) for c in st}
print("samples", tot, " ".join("%s=%.1f%%" % (k[6:], 100 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]))
ops = collections.Counter()
for r in data:
    src = r[ix['Source']].split()
    op = src[1] if src[0].startswith('@') else src[0]
    ops[op.split('.')[0]] += int(r[ix['Instructions Executed']])
t = sum(ops.values())
print("dyn warp instr", t, ("thread-instr/px %.0f" % (t * 32 / npx)) if npx else "")
print(" ".join("%s=%.1f%%" % (k, 100 * v / t) for k, v in ops.most_common(30)))
top = sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:int(sys.argv[4]) if len(sys.argv) > 4 else 14]
for r in top:
    print(r[ix['# Samples']].rjust(6), r[ix['Source']][:80].ljust(80), " ".join("%s=%s" % (c[6:], r[ix[c]]) for c in st if int(r[ix[c]]) > 0.2 * int(r[ix['# Samples']])))
