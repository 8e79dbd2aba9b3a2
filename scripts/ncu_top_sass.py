#!/usr/bin/env python
"""Top stalled SASS instructions of the first kernel in an .ncu-rep (source page)."""
import csv, subprocess, sys
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25; which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# split per kernel
kernels = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; kernels.append(cur); continue
    if cur is not None: cur["rows"].append(r)
k = kernels[which]
hdr = k["rows"][0]; data = k["rows"][1:]
iS = hdr.index("# Samples"); iSrc = hdr.index("Source"); iEx = hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iS]) for r in data)
totex = sum(int(r[iEx]) for r in data)
print(k["name"], "total samples", tot, "warp-instr executed", totex)
idx = sorted(range(len(data)), key=lambda i: -int(data[i][iS]))[:topn]
for i in sorted(idx):
    r = data[i]
    st = sorted(((int(r[c]), hdr[c]) for c in stall_cols), reverse=True)[:2]
    print("%5d %5.1f%% ex=%8s  %-60s %s" % (i, 100 * int(r[iS]) / tot, r[iEx], r[iSrc].strip()[:60], " ".join("%s=%d" % (h[6:], v) for v, h in st if v)))
