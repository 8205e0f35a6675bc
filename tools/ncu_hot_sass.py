#!/usr/bin/env python
"""Top stall-sample SASS lines of an `ncu --page source --csv` dump (SASS view): where a kernel's warps wait."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
si, src, ex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stall = {h: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
data = []
for k, r in enumerate(rows[2:]):
    try:
        data.append((float(r[si]), k, r))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for smp, k, r in sorted(data, reverse=True)[:top]:
    why = sorted(((float(r[i] or 0), h) for h, i in stall.items()), reverse=True)[:2]
    print("%5.1f%%  #%4d  %-70s %s" % (100 * smp / tot, k, r[src].strip()[:70], " ".join("%s=%d" % (h[6:], v) for v, h in why if v)))
