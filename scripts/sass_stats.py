#!/usr/bin/env python
"""Static SASS statistics of one kernel in an object file: instruction count, opcode histogram, spills, branch map.

    python scripts/sass_stats.py OBJECT.o MANGLED_SUBSTRING [--branches]
"""
import collections, re, subprocess, sys
obj, sym = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
blocks = out.split("Function : ")
blk = next(b for b in blocks if sym in b.split("\n")[0])
ins = []
for l in blk.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
ops = collections.Counter()
for a, i in ins:
    t = i.split()
    op = t[1] if t[0].startswith("@") else t[0]
    ops[op.split(".")[0]] += 1
print(blk.split("\n")[0], "instructions:", len(ins))
print(" ".join("%s=%d" % kv for kv in ops.most_common(45)))
print("STL", sum(1 for a, i in ins if "STL" in i), "LDL", sum(1 for a, i in ins if "LDL" in i))
if "--branches" in sys.argv:
    for a, i in ins:
        if re.search(r"\b(BRA|BAR|SYNCS|UTMALDG|EXIT|CALL|RET)\b", i):
            print("%05x %s" % (a, i))
