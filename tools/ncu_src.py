#!/usr/bin/env python
"""Summarise `ncu --page source --csv` of one kernel: stall mix and the hottest SASS lines.
usage: tools/ncu_src.py REPORT.ncu-rep KERNEL_REGEX [launch_skip]"""
import csv, subprocess, sys, io
rep, rx = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + rx, "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
print(rows[0][:2])
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
si, src, ie = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stalls = [(h, i) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[si]) for r in data)
print("total samples", tot, "warp-instr executed", sum(int(r[ie]) for r in data), "SASS lines", len(data))
agg = {h: sum(int(r[i]) for r in data) for h, i in stalls}
print("stall mix:", [(h[6:], v, "%.0f%%" % (100 * v / max(tot, 1))) for h, v in sorted(agg.items(), key=lambda x: -x[1])[:8]])
top = sorted(enumerate(data), key=lambda x: -int(x[1][si]))[:int(sys.argv[4]) if len(sys.argv) > 4 else 30]
for n, r in top:
    print("%5d %6s %8s  %-90s %s" % (n, r[si], r[ie], r[src].strip()[:90], [(h[6:], r[i]) for h, i in stalls if int(r[i]) > int(r[si]) // 4]))
