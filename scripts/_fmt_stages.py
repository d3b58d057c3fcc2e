import json, sys
for line in sys.stdin:
    line = line.strip()
    try:
        j = json.loads(line)
        print("  %-16s %.4f ms  %.3f of HBM roofline" % (j["bench"], j["ms"], j["frac_hbm_roofline"]))
    except Exception:
        if line:
            print("  ! " + line[:200])
